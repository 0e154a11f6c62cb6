"""GPU parity tests of the callers either side of the loss path (SURVEY.md
section 8f): disparity head (n2), discriminator input glue (n3), evaluation
post-processing (n4) -- through the drop-in functions, i.e. the C ABI."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from uncertainty_model_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ------------------------------------------------------------------- n2 ----
@pytest.mark.parametrize('scale', [0.3, 1.0])
def test_disparity_head_matches_oracle(dev, scale):
    """scale * sigmoid(logits) and its gradient, all levels in one launch;
    tolerances of north_star (1e-5 values, 1e-4 gradients, fp32)."""
    from oracle import post_port as PP
    from uncertainty_model_b200.train import utils as U
    g = torch.Generator().manual_seed(2)
    shapes = [(3, 4, 64 >> i, 96 >> i) for i in range(4)]
    shapes[3] = (3, 4, 7, 11)                       # unaligned level: scalar path
    logits = [4 * torch.randn(s, generator=g) for s in shapes]
    ups = [torch.randn(s, generator=g) for s in shapes]
    ol = [x.double().requires_grad_(True) for x in logits]
    op = PP.disparity_head(ol, scale)
    sum((p * u.double()).sum() for p, u in zip(op, ups)).backward()
    gl = [x.to(dev).requires_grad_(True) for x in logits]
    gp = U.disparity_head(gl, scale)
    sum((p * u.to(dev)).sum() for p, u in zip(gp, ups)).backward()
    for a, b, x, y in zip(gp, op, gl, ol):
        assert a.shape == b.shape
        assert rel_l2(a.detach().cpu().numpy(), b.detach().numpy()) < 1e-5
        assert rel_l2(x.grad.cpu().numpy(), y.grad.numpy()) < 1e-4


def test_disparity_head_feeds_the_loss(dev):
    """logits -> head -> 4-scale loss -> backward: gradients reach the logits
    (the step immediately before the hot path, decoder.py:246)."""
    from oracle import loss_port as P
    from oracle import post_port as PP
    from oracle.make_golden import loss_config
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    import parity
    g = torch.Generator().manual_seed(4)
    b, h, w, scale = 2, 64, 128, 0.3
    stereo = torch.rand(b, 6, h, w, generator=g)
    logits = [torch.randn(b, 4, h >> i, w >> i, generator=g) for i in range(4)]
    cfg = loss_config('bayesian')
    ol = [x.double().requires_grad_(True) for x in logits]
    op = PP.disparity_head(ol, scale)
    opyr = P.pyramid(stereo.double(), 4)
    odl, oel = P.total_loss(opyr, op, P.recon_pyramid(op, opyr), cfg)
    (odl + oel).backward()
    gl = [x.to(dev).requires_grad_(True) for x in logits]
    gp = U.disparity_head(gl, scale)
    pyr = U.scale_pyramid(stereo.to(dev), 4)
    dl, el = L.TukraUncertaintyLoss(**cfg)(pyr, gp, U.reconstruct_pyramid(gp, pyr))
    (dl + el).backward()
    assert abs(dl.item() - odl.item()) <= 1e-5 * abs(odl.item())
    assert abs(el.item() - oel.item()) <= 1e-5 * abs(oel.item())
    masks = parity.masks_for(stereo, [p.detach().float() for p in op], cfg)
    parity.check_grads([x.grad for x in gl], [x.grad for x in ol], masks)


# ------------------------------------------------------------------- n3 ----
@pytest.mark.parametrize('lazy', [True, False])
def test_discriminator_input_matches_oracle(dev, lazy):
    """[images ; reconstructions] along the batch axis for every level, the
    reconstruction half warped on the fly (lazy pyramid) or copied from the
    materialised one."""
    from oracle import loss_port as P
    from oracle import post_port as PP
    from oracle.make_golden import make_inputs
    from uncertainty_model_b200 import functional as K
    from uncertainty_model_b200.train import utils as U
    left, right, preds = make_inputs(3, 40, 72, 0.3, 13)
    stereo = torch.cat([left, right], 1)
    opyr = P.pyramid(stereo, 4)
    ref = PP.discriminator_input(opyr, P.recon_pyramid(preds, opyr))
    pyr = U.scale_pyramid(stereo.to(dev), 4)
    gp = [p.to(dev).requires_grad_(True) for p in preds]
    rec = U.reconstruct_pyramid(gp, pyr)
    if not lazy:
        rec = list(rec)
    seen = {}

    class Disc(torch.nn.Module):
        def forward(self, pyramid):
            seen['pyramid'] = pyramid
            return torch.stack([p.mean(dim=(1, 2, 3)) for p in pyramid]).sum(0)[:, None]

    out = U.run_discriminator(pyr, rec, Disc(), torch.nn.MSELoss(), 3)
    assert out.dim() == 0
    for a, b in zip(seen['pyramid'], ref):
        assert a.shape == b.shape and not a.requires_grad
        assert np.allclose(a.cpu().numpy(), b.numpy(), atol=2e-6)
    if lazy:
        assert not rec.materialised          # the glue did not materialise it
    # the reference's loss value: labels 1 for the real half (utils.py:268-273)
    pred = torch.stack([p.mean(dim=(1, 2, 3)) for p in ref]).sum(0)[:, None]
    labels = torch.zeros_like(pred)
    labels[:3] = 1
    want = torch.nn.functional.mse_loss(pred, labels) / 2
    assert abs(out.item() - want.item()) <= 1e-5 * abs(want.item())


# ------------------------------------------------------------------- n4 ----
def test_combine_disparity_matches_the_numpy_original(dev):
    from oracle import post_port as PP
    from uncertainty_model_b200.train import utils as U
    g = torch.Generator().manual_seed(6)
    for shape in ((1, 33, 77), (2, 64, 128)):
        left = 0.3 * torch.rand(shape, generator=g)
        right = 0.3 * torch.rand(shape, generator=g)
        for alpha, beta in ((20, 0.05), (7.5, 0.2)):
            ref = PP.combine_disparity(left, right, alpha, beta)
            out = U.combine_disparity(left.to(dev), right.to(dev), 'cpu', alpha, beta)
            assert out.dtype == ref.dtype == torch.float64
            assert out.shape == ref.shape
            assert np.allclose(out.numpy(), ref.numpy(), rtol=0, atol=1e-15)
    gold = np.load(os.path.join(GOLDEN, 'post.npz'))
    out = U.combine_disparity(torch.from_numpy(gold['left']).to(dev),
                              torch.from_numpy(gold['right']).to(dev))
    assert np.allclose(out.numpy(), gold['combined'], rtol=0, atol=1e-15)


def test_to_heatmap_matches_oracle(dev):
    from oracle import post_port as PP
    from uncertainty_model_b200.train import utils as U
    g = torch.Generator().manual_seed(8)
    table = torch.rand(256, 3, generator=g, dtype=torch.float64)
    x = torch.rand(1, 37, 53, generator=g)
    x[0, 0, :6] = torch.tensor([0.0, 1.0, -0.25, 1.5, float('nan'), 255.0 / 256.0])
    for inverse in (False, True):
        ref = PP.to_heatmap(x, table.numpy(), inverse)
        out = U.to_heatmap(x.to(dev), 'cpu', inverse, table)
        assert out.dtype == torch.float64 and out.shape == (3, 37, 53)
        assert np.array_equal(out.numpy(), ref.numpy())
    with pytest.raises(ValueError, match='no CPU fallback'):
        U.to_heatmap(x, 'cpu', False, table)
