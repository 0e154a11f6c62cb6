"""GPU parity tests of the evaluation front half (SURVEY.md section 8f n1,
reference train/evaluate.py:136-160): the Gaussian SSIM metric and the fused
per-batch evaluation against the CPU oracle ports."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from uncertainty_model_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


@pytest.mark.parametrize('shape', [(2, 3, 48, 96), (1, 3, 37, 53), (3, 1, 64, 40)])
def test_gaussian_ssim_matches_oracle(dev, shape):
    """torchmetrics' SSIM as evaluate.py:142-146 calls it (kernel 11, sigma 1.5,
    data_range 1): per-image values to 1e-5 of the fp64 oracle; 'sum'."""
    from oracle import ssim_port as SS
    from uncertainty_model_b200.train import evaluate as E
    g = torch.Generator().manual_seed(3)
    target = torch.rand(shape, generator=g)
    preds = (target + 0.2 * (torch.rand(shape, generator=g) - 0.5)).clamp(0, 1)
    ref = SS.ssim_per_image(preds.double(), target.double())
    ref32 = SS.ssim_per_image(preds, target)
    out = E.ssim(preds.to(dev), target.to(dev), reduction=None)
    assert out.shape == ref.shape
    assert np.allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=0)
    assert np.allclose(ref32.numpy(), ref.numpy(), rtol=1e-5, atol=0)
    s = E.ssim(preds.to(dev), target.to(dev), kernel_size=11, reduction='sum',
               data_range=1.0)
    assert abs(s.item() - float(ref.sum())) <= 1e-5 * abs(float(ref.sum()))
    # identical images -> exactly 1 per image; symmetric in its arguments
    one = E.ssim(target.to(dev), target.to(dev), reduction=None)
    assert np.allclose(one.cpu().numpy(), 1.0, rtol=0, atol=1e-6)
    swapped = E.ssim(target.to(dev), preds.to(dev), reduction=None)
    assert np.allclose(swapped.cpu().numpy(), out.cpu().numpy(), rtol=1e-6)
    for k in (7, 15):
        r = SS.ssim_per_image(preds.double(), target.double(), kernel_size=k)
        o = E.ssim(preds.to(dev), target.to(dev), kernel_size=k, reduction=None)
        assert np.allclose(o.cpu().numpy(), r.numpy(), rtol=1e-5, atol=0)


def test_fused_evaluation_batch_matches_oracle_and_the_drop_in_sequence(dev):
    """evaluate.py:136-160 as one call: reconstructions + error map from one
    launch, SSIM sums, curves, AUSE / AURG -- against the CPU oracle ports of
    the same lines, and against the drop-in functions called one by one."""
    from oracle import loss_port as P
    from oracle import spars_port as SP
    from oracle import ssim_port as SS
    from oracle.make_golden import make_inputs
    from uncertainty_model_b200.train import evaluate as E
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import sparsification as S
    from uncertainty_model_b200.train import utils as U
    left, right, preds = make_inputs(2, 64, 128, 0.3, 41)
    prediction = preds[0]
    images = torch.cat([left, right], 1)
    dl, dr = prediction[:, 0:1], prediction[:, 1:2]
    unc = prediction[:, 2:4]
    recon_ref = torch.cat((P.warp_to_left(dl, right), P.warp_to_right(dr, left)), 1)
    err_ref = P.image_error(images.double(), recon_ref.double(), alpha=1.0)

    g = torch.Generator().manual_seed(9)
    rnd = torch.rand(err_ref.shape, generator=g)
    out = E.evaluate_batch(left.to(dev), right.to(dev), prediction.to(dev),
                           device=dev, random_error=rnd.to(dev))
    assert np.allclose(out['recon'].cpu().numpy(), recon_ref.numpy(), atol=2e-6)
    assert np.allclose(out['error'].cpu().numpy(), err_ref.numpy(), atol=1e-5)
    for key, a, b in (('left_ssim', recon_ref[:, 0:3], left),
                      ('right_ssim', recon_ref[:, 3:6], right)):
        want = float(SS.ssim_sum(a.double(), b.double()))
        # (white-noise pairs: the per-pixel values are O(0.1) of either sign and
        #  the sum over the batch nearly cancels -- absolute, per image)
        assert abs(out[key].item() - want) <= 1e-5 * left.size(0) * 0.1, key
    # curves: bit-exact against the canonical oracle fed the GPU error map
    e_np = out['error'].cpu().numpy()
    assert np.array_equal(out['oracle_curve'].cpu().numpy(),
                          SP.curve_canonical(e_np, e_np))
    assert np.array_equal(out['pred_curve'].cpu().numpy(),
                          SP.curve_canonical(e_np, unc.contiguous().numpy()))
    assert np.float32(out['ause'].item()) == SP.ause_canonical(
        out['oracle_curve'].cpu().numpy(), out['pred_curve'].cpu().numpy())
    rc = SP.curve_canonical(e_np, rnd.numpy())
    assert np.float32(out['aurg'].item()) == SP.ause_canonical(
        out['pred_curve'].cpu().numpy(), rc)

    # the drop-in functions one by one (what evaluate.py itself would call)
    gl, gr, gp = left.to(dev), right.to(dev), prediction.to(dev)
    recon = torch.cat((U.reconstruct_left_image(gp[:, 0:1], gr),
                       U.reconstruct_right_image(gp[:, 1:2], gl)), dim=1)
    error = L.WeightedSSIMLoss(alpha=1).image_error(torch.cat([gl, gr], 1), recon)
    assert np.allclose(recon.cpu().numpy(), out['recon'].cpu().numpy(), atol=2e-6)
    assert np.allclose(error.cpu().numpy(), out['error'].cpu().numpy(), atol=2e-6)
    ause = S.ause(S.curve(error, error, device=dev),
                  S.curve(error, gp[:, 2:4], device=dev))
    assert abs(ause.item() - out['ause'].item()) <= 1e-4 * abs(out['ause'].item()) + 1e-7
