"""GPU parity tests of the loss path: the CUDA kernels (through the drop-in
classes, i.e. through the C ABI) against the golden fixtures generated from the
real reference and against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): loss 1e-5 relative, gradients 1e-4
relative against the fp64 run of the reference/oracle -- element by element,
see tests/parity.py: the elements where the loss is not differentiable within
fp32 rounding are set aside by an explicit mask computed from the oracle
(oracle/kinks.py), never by the size of their error."""
import os

import numpy as np
import pytest
import torch

import parity
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

LOSS_REL = 1e-5
GRAD_REL = 1e-4


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from uncertainty_model_b200 import _lib
    _lib.lib()      # must be the in-tree libusl.so; raises if missing
    return torch.device('cuda:0')


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def grad_rel(mine, ref, mask=None):
    """Relative L2 error of one gradient tensor over the elements outside
    `mask` (an explicit kink mask from oracle/kinks.py; None = every element).
    The full-loss tests use parity.check_grads, which also bounds every single
    element and the size of the mask."""
    mine = mine.detach().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert np.isfinite(mine).all()
    if mask is None:
        return rel_l2(mine, ref)
    keep = ~np.asarray(mask, dtype=bool)
    assert keep.mean() >= 1.0 - 9 * parity.MASK_MAX
    return rel_l2(mine[keep], ref[keep])


def cases():
    from oracle.make_golden import loss_config
    return {
        'l1_default': loss_config('l1'),
        'bayesian_default': loss_config('bayesian'),
        'log_bayesian_scale1': loss_config('log_bayesian'),
        'l1_allterms': loss_config('l1', smoothness_weight=0.6,
                                   consistency_weight=0.8),
        'bayesian_pooling': loss_config('bayesian', smoothness_weight=0.6,
                                        consistency_weight=0.8, pooling=True),
        'bayesian_smooth': loss_config('bayesian'),
    }


def run_ours(dev, stereo, preds, cfg, materialise=False, flags=0):
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    images = stereo.to(dev)
    gp = [p.to(dev).clone().requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(images, len(gp))
    rec = U.reconstruct_pyramid(gp, pyr)
    if materialise:
        rec = list(rec)
    fn = L.TukraUncertaintyLoss(**cfg).to(dev)
    fn.kernel_flags = flags
    dl, el = fn(pyr, gp, rec, 0, None)
    (dl + el).backward()
    return dl, el, gp, pyr, rec, fn


# ------------------------------------------------------------------ pyramid --
def test_pyramid_matches_reference(dev):
    from uncertainty_model_b200.train import utils as U
    g = load('loss_l1_default.npz')
    stereo = torch.cat([torch.from_numpy(g['left']),
                        torch.from_numpy(g['right'])], 1).to(dev)
    pyr = U.scale_pyramid(stereo, 4)
    assert pyr[0].data_ptr() == stereo.data_ptr()      # level 0 is aliased
    for i in range(4):
        assert pyr[i].shape == g[f'pyr{i}_f32'].shape
        assert np.allclose(pyr[i].cpu().numpy(), g[f'pyr{i}_f32'], atol=5e-7)


def test_pyramid_odd_shapes_and_channel_slices(dev):
    from oracle import loss_port as P
    from uncertainty_model_b200.train import utils as U
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 8, 37, 53, generator=g)
    ours = U.scale_pyramid(x.to(dev)[:, 1:7], 3)       # strided batch
    ref = P.pyramid(x[:, 1:7], 3)
    for a, b in zip(ours, ref):
        assert np.allclose(a.cpu().numpy(), b.numpy(), atol=5e-7)


# --------------------------------------------------------------------- warp --
def test_reconstruct_matches_reference(dev):
    from uncertainty_model_b200.train import utils as U
    g = load('components.npz')
    im = torch.from_numpy(g['images']).to(dev)
    pr = torch.from_numpy(g['pred']).to(dev)
    left = U.reconstruct_left_image(pr[:, 0:1], im[:, 3:6])
    right = U.reconstruct_right_image(pr[:, 1:2], im[:, 0:3])
    assert np.allclose(left.cpu().numpy(), g['recon_left_f64'], atol=5e-6)
    assert np.allclose(right.cpu().numpy(), g['recon_right_f64'], atol=5e-6)
    gen = U.reconstruct(-pr[:, 0:1], im[:, 3:6])
    assert torch.equal(gen, left)


@pytest.mark.parametrize('shape', [(2, 3, 32, 48, 0.5), (1, 1, 17, 37, 1.0),
                                   (2, 3, 64, 512, 0.3)])
def test_reconstruct_is_differentiable_in_the_sampled_image(dev, shape):
    """utils.py:96-97: the reference's grid_sample is differentiable w.r.t. the
    image it samples as well.  Against torch autograd on the fp64 port; and
    the transposed warp is deterministic (ATen's uses float atomics)."""
    from oracle import loss_port as P
    from uncertainty_model_b200.train import utils as U
    b, c, h, w, scale = shape
    g = torch.Generator().manual_seed(11)
    disp = scale * torch.rand(b, 1, h, w, generator=g) - 0.25 * scale
    image = torch.rand(b, c, h, w, generator=g)
    upstream = torch.randn(b, c, h, w, generator=g)
    d64 = disp.double().requires_grad_(True)
    i64 = image.double().requires_grad_(True)
    (P.warp(d64, i64) * upstream.double()).sum().backward()
    outs = []
    for _ in range(2):
        dd = disp.to(dev).requires_grad_(True)
        ii = image.to(dev).requires_grad_(True)
        out = U.reconstruct(dd, ii)
        (out * upstream.to(dev)).sum().backward()
        outs.append((dd.grad.clone(), ii.grad.clone()))
    gd, gi = outs[0]
    # (the tap weights come from an fp32 sampling coordinate of magnitude w:
    #  their absolute error grows with the width, 512 * 2^-24 = 3e-5)
    err = rel_l2(gi.cpu().numpy(), i64.grad.numpy())
    worst = np.abs(gi.cpu().numpy() - i64.grad.numpy()).max() / \
        float(i64.grad.abs().max())
    print('image gradient: rel-L2', err, 'worst element', worst)
    assert err <= 1e-5 * max(1.0, w / 48)
    assert worst <= 2e-5 * max(1.0, w / 48)
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[1][0])


def test_reconstruct_pyramid_is_lazy_and_differentiable(dev):
    from oracle import loss_port as P
    from oracle.make_golden import make_inputs
    from uncertainty_model_b200.train import utils as U
    left, right, preds = make_inputs(2, 32, 48, 0.5, 7)
    stereo = torch.cat([left, right], 1)
    gp = [p.to(dev).requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(stereo.to(dev), 4)
    rec = U.reconstruct_pyramid(gp, pyr)
    assert not rec.materialised and len(rec) == 4
    w = [torch.rand_like(r) for r in rec]               # materialises
    assert rec.materialised
    sum((r * k).sum() for r, k in zip(rec, w)).backward()

    op = [p.double().requires_grad_(True) for p in preds]
    orec = P.recon_pyramid(op, P.pyramid(stereo.double(), 4))
    sum((r * k.cpu().double()).sum() for r, k in zip(orec, w)).backward()
    from oracle import kinks as KK
    opyr = P.pyramid(stereo.double(), 4)
    for i in range(4):
        assert np.allclose(rec[i].detach().cpu().numpy(),
                           orec[i].detach().numpy(), atol=2e-5)
        d = op[i].detach()
        mask = torch.stack((KK.warp_kinks(-d[:, 0:1], opyr[i][:, 3:6]),
                            KK.warp_kinks(d[:, 1:2], opyr[i][:, 0:3])), 1)
        assert grad_rel(gp[i].grad[:, 0:2], op[i].grad.numpy()[:, 0:2],
                        mask.numpy()) < GRAD_REL


# --------------------------------------------------------------- fused loss --
@pytest.mark.parametrize('name', sorted(cases()))
@pytest.mark.parametrize('materialise', [False, True])
def test_total_loss_matches_reference_fixture(dev, name, materialise):
    cfg = cases()[name]
    g = load(f'loss_{name}.npz')
    stereo = torch.cat([torch.from_numpy(g['left']),
                        torch.from_numpy(g['right'])], 1)
    preds = [torch.from_numpy(g[f'pred{i}']) for i in range(4)]
    dl, el, gp, pyr, rec, fn = run_ours(dev, stereo, preds, cfg, materialise)
    for mine, key in ((dl, 'disp_loss'), (el, 'error_loss')):
        for tag in ('f32', 'f64'):
            ref = float(g[f'{key}_{tag}'])
            assert abs(mine.item() - ref) <= LOSS_REL * abs(ref), \
                (key, tag, mine.item(), ref)
    ref = parity.oracle_reference(stereo, preds, cfg)
    parity.check_grads([p.grad for p in gp],
                       [g[f'grad{i}_f64'] for i in range(4)], ref['masks'],
                       name, mask_max=ref['mask_max'])
    # loss.py:548 -- the last scale's error map stays readable
    prev = fn.wssim.previous_image_error
    assert np.allclose(prev.cpu().numpy(), g['err3_f32'], atol=1e-5)


@pytest.mark.parametrize('loss_type', ['l1', 'bayesian', 'log_bayesian'])
@pytest.mark.parametrize('shape', [(2, 64, 128, 0.3), (1, 96, 160, 1.0),
                                   (3, 40, 600, 0.3), (1, 33, 77, 0.5)])
def test_total_loss_matches_oracle(dev, loss_type, shape):
    """Seeded inputs, CPU oracle in fp64; (3,40,600) spans several column
    tiles and exercises the seams; (1,33,77): odd sizes, units that do not
    fill their warps, unaligned rows (no bulk copies)."""
    from oracle import loss_port as P
    from oracle.make_golden import loss_config, make_inputs
    b, h, w, scale = shape
    cfg = loss_config(loss_type, smoothness_weight=0.25)
    left, right, preds = make_inputs(b, h, w, scale, 31)
    stereo = torch.cat([left, right], 1)
    ref = parity.oracle_reference(stereo, preds, cfg)
    rdl, rel = ref['disp_loss'], ref['error_loss']
    dl, el, gp, *_ = run_ours(dev, stereo, preds, cfg)
    assert abs(dl.item() - float(rdl)) <= LOSS_REL * abs(float(rdl))
    assert abs(el.item() - float(rel)) <= LOSS_REL * abs(float(rel))
    parity.check_grads([p.grad for p in gp], ref['grads'], ref['masks'])


@pytest.mark.parametrize('kind', ['ramps', 'ramps_with_jumps'])
def test_smooth_disparities_match_oracle(dev, kind):
    """Piecewise-linear disparities: the destination columns of the transposed
    warp are monotone along most chunks, so the scatter groups equal
    destinations as contiguous runs (slopes < 1: runs of two and three; the
    white-noise inputs of the other tests always take the match.any branch).
    `ramps_with_jumps` adds disparity steps inside rows: chunks of both kinds
    side by side."""
    from oracle import loss_port as P
    from oracle.make_golden import loss_config, make_inputs
    b, h, w = 2, 64, 256
    cfg = loss_config('bayesian', smoothness_weight=0.25)
    left, right, preds = make_inputs(b, h, w, 0.3, 35)
    g = torch.Generator().manual_seed(5)
    smooth = []
    for i, p in enumerate(preds):
        hh, ww = p.shape[2:]
        x = torch.linspace(0, 1, ww).view(1, 1, 1, ww)
        y = torch.linspace(0, 1, hh).view(1, 1, hh, 1)
        a = torch.rand(b, 4, 1, 1, generator=g)
        q = 0.02 + 0.25 * (a * x + (1 - a) * 0.5 * y) + 0.002 * p
        if kind == 'ramps_with_jumps':
            q = q + 0.05 * (x > 0.37).float() - 0.04 * (x > 0.71).float()
        smooth.append(q.contiguous())
    stereo = torch.cat([left, right], 1)
    ref = parity.oracle_reference(stereo, smooth, cfg)
    rdl, rel = ref['disp_loss'], ref['error_loss']
    dl, el, gp, *_ = run_ours(dev, stereo, smooth, cfg)
    assert abs(dl.item() - float(rdl)) <= LOSS_REL * abs(float(rdl))
    assert abs(el.item() - float(rel)) <= LOSS_REL * abs(float(rel))
    # (ramps put whole runs of sampling coordinates on integers: the mask is
    #  larger than on random inputs -- 0.8 % at the 8 x 32 level)
    parity.check_grads([p.grad for p in gp], ref['grads'], ref['masks'],
                       kind, mask_max=0.05)


def test_general_strip_kernels_match_oracle(dev):
    """The same training step kept off the column kernels
    (USL_SCALE_GENERAL_KERNELS): forward sums, two-launch backward on the
    general strip kernels -- the path every call the hot kernels do not take
    runs on."""
    from oracle.make_golden import loss_config, make_inputs
    from uncertainty_model_b200._lib import USL_SCALE_GENERAL_KERNELS
    cfg = loss_config('bayesian', smoothness_weight=0.25)
    left, right, preds = make_inputs(2, 64, 128, 0.3, 33)
    stereo = torch.cat([left, right], 1)
    ref = parity.oracle_reference(stereo, preds, cfg)
    rdl, rel = ref['disp_loss'], ref['error_loss']
    dl, el, gp, *_ = run_ours(dev, stereo, preds, cfg,
                              flags=USL_SCALE_GENERAL_KERNELS)
    assert abs(dl.item() - float(rdl)) <= LOSS_REL * abs(float(rdl))
    assert abs(el.item() - float(rel)) <= LOSS_REL * abs(float(rel))
    parity.check_grads([p.grad for p in gp], ref['grads'], ref['masks'])
    # ... and it really is another kernel family: not bit-identical
    dl2, el2, gp2, *_ = run_ours(dev, stereo, preds, cfg)
    assert any(not torch.equal(a.grad, b.grad) for a, b in zip(gp, gp2))


def test_repeated_backward_with_other_upstream_gradients(dev):
    """backward(retain_graph=True) twice with different upstream gradients
    (GradScaler-style): each call returns its own correctly scaled gradients;
    the tensors returned by the first call are not touched by the second."""
    from oracle.make_golden import loss_config, make_inputs
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    cfg = loss_config('bayesian')
    left, right, preds = make_inputs(2, 48, 96, 0.3, 21)
    stereo = torch.cat([left, right], 1).to(dev)

    def grads_of(weights):
        gp = [p.to(dev).requires_grad_(True) for p in preds]
        pyr = U.scale_pyramid(stereo, 4)
        dl, el = L.TukraUncertaintyLoss(**cfg)(
            pyr, gp, U.reconstruct_pyramid(gp, pyr))
        outs = []
        for i, (a, b) in enumerate(weights):
            outs.append(torch.autograd.grad(
                a * dl + b * el, gp, retain_graph=i + 1 < len(weights)))
        return outs

    single = {w: grads_of([w])[0] for w in ((3.0, 0.5), (1.0, 1.0))}
    first, second = grads_of([(3.0, 0.5), (1.0, 1.0)])
    for a, b in zip(first, single[(3.0, 0.5)]):
        assert torch.equal(a, b)
    for a, b in zip(second, single[(1.0, 1.0)]):
        assert torch.equal(a, b)
    first, second = grads_of([(1.0, 1.0), (3.0, 0.5)])
    for a, b in zip(first, single[(1.0, 1.0)]):
        assert torch.equal(a, b)
    for a, b in zip(second, single[(3.0, 0.5)]):
        assert torch.equal(a, b)


def test_separate_upstream_gradients(dev):
    """disp_loss and error_loss back-propagated with different weights."""
    from oracle import loss_port as P
    from oracle.make_golden import loss_config, make_inputs
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    cfg = loss_config('bayesian', smoothness_weight=0.5)
    left, right, preds = make_inputs(2, 48, 80, 0.4, 17)
    stereo = torch.cat([left, right], 1)
    op = [p.double().requires_grad_(True) for p in preds]
    opyr = P.pyramid(stereo.double(), 4)
    odl, oel = P.total_loss(opyr, op, P.recon_pyramid(op, opyr), cfg)
    (0.3 * odl - 2.0 * oel).backward()
    gp = [p.to(dev).requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(stereo.to(dev), 4)
    dl, el = L.TukraUncertaintyLoss(**cfg)(pyr, gp,
                                           U.reconstruct_pyramid(gp, pyr))
    (0.3 * dl - 2.0 * el).backward()
    masks = parity.masks_for(stereo, preds, cfg)   # kinks: same for any weights
    parity.check_grads([p.grad for p in gp], [p.grad for p in op], masks)
    # only one of the two outputs used
    gp2 = [p.to(dev).requires_grad_(True) for p in preds]
    dl2, _ = L.TukraUncertaintyLoss(**cfg)(pyr, gp2,
                                           U.reconstruct_pyramid(gp2, pyr))
    dl2.backward()
    op2 = [p.double().requires_grad_(True) for p in preds]
    odl2, _ = P.total_loss(opyr, op2, P.recon_pyramid(op2, opyr), cfg)
    odl2.backward()
    parity.check_grads([p.grad for p in gp2], [p.grad for p in op2], masks)


@pytest.mark.parametrize('shape', [(4, 64, 256, 1.0), (16, 256, 512, 0.3)])
def test_backward_is_bitwise_deterministic(dev, shape):
    """Small two-view units, and the benchmark shape (BASELINE config 2: the
    one-view w = 512 units, multi-wave strips, every scale concurrently)."""
    from oracle.make_golden import loss_config, make_inputs
    cfg = loss_config('bayesian', smoothness_weight=0.5 if shape[0] == 4 else 0)
    left, right, preds = make_inputs(*shape, 5)
    stereo = torch.cat([left, right], 1)
    runs = []
    for _ in range(3):
        dl, el, gp, *_ = run_ours(dev, stereo, preds, cfg)
        runs.append((dl.item(), el.item(), [p.grad.clone() for p in gp]))
    for r in runs[1:]:
        assert r[0] == runs[0][0] and r[1] == runs[0][1]
        for a, b in zip(r[2], runs[0][2]):
            assert torch.equal(a, b)


BENCH_SHAPES = {
    # BASELINE.json configs as bench.py runs them (c4: the reference batch of 8
    # needs minutes of fp64 CPU time; 2 samples run the same units -- 3 column
    # tiles x 6 strips per view -- on fewer SMs)
    'c1': (2, 256, 512, 'l1'),
    'c2': (16, 256, 512, 'bayesian'),
    'c3_shard': (8, 192, 384, 'l1'),
    'c3_full': (64, 192, 384, 'l1'),
    'c4': (2, 512, 1024, 'l1'),
}


@pytest.mark.parametrize('name', sorted(BENCH_SHAPES))
def test_benchmark_shapes_elementwise(dev, name):
    """The kernel instantiations the benchmark numbers come from (one-view
    w = 512 TMA units with 86-row strips, w = 384 units, column tiles at
    w = 1024), gradients compared ELEMENT BY ELEMENT with the fp64 oracle;
    the reference's own fp32 scalars (anchors.npz) for the losses."""
    from oracle.make_golden import loss_config, make_inputs
    b, h, w, lt = BENCH_SHAPES[name]
    cfg = loss_config(lt)
    left, right, preds = make_inputs(b, h, w, 0.3, 0)
    stereo = torch.cat([left, right], 1)
    dl, el, gp, *_ = run_ours(dev, stereo, preds, cfg)
    g = load('anchors.npz')
    if f'{name}_disp_loss' in g.files:
        assert [int(v) for v in g[f'{name}_shape']] == [b, h, w]
        for mine, key in ((dl, 'disp_loss'), (el, 'error_loss')):
            ref32 = float(g[f'{name}_{key}'])
            assert abs(mine.item() - ref32) <= LOSS_REL * abs(ref32), (name, key)
    ref = parity.oracle_reference(stereo, preds, cfg)
    for mine, key in ((dl, 'disp_loss'), (el, 'error_loss')):
        assert abs(mine.item() - float(ref[key])) <= \
            LOSS_REL * abs(float(ref[key])), (name, key)
    stats = parity.check_grads([p.grad for p in gp], ref['grads'],
                               ref['masks'], name)
    parity.record(name, stats)


@pytest.mark.parametrize('shape', [(3, 40, 72, 'l1'), (16, 256, 512, 'bayesian')])
def test_mirror_and_view_swap_symmetry(dev, shape):
    """Size-independent property (holds for the reference's formulas; checked
    on the fp64 oracle in tests/test_oracle.py): mirror every image left-right
    and swap the two views -- the stereo geometry is the same scene seen in a
    mirror, so both losses are unchanged and the gradients are the mirrored,
    swapped ones.  Run at the benchmark shape too.  fp32 rounds the mirrored
    coordinates differently, so a few taps flip: losses to 1e-5, gradients
    element-wise on all but 1 % of the elements."""
    from oracle.make_golden import loss_config, make_inputs
    b, h, w, lt = shape
    cfg = loss_config(lt, smoothness_weight=0.25)
    left, right, preds = make_inputs(b, h, w, 0.3, 23)

    def mirrored(p):
        return torch.stack([p[:, 1].flip(2), p[:, 0].flip(2),
                            p[:, 3].flip(2), p[:, 2].flip(2)], 1).contiguous()

    dl, el, gp, *_ = run_ours(dev, torch.cat([left, right], 1), preds, cfg)
    dl2, el2, gp2, *_ = run_ours(
        dev, torch.cat([right.flip(3), left.flip(3)], 1).contiguous(),
        [mirrored(p) for p in preds], cfg)
    assert abs(dl.item() - dl2.item()) <= LOSS_REL * abs(dl.item())
    assert abs(el.item() - el2.item()) <= LOSS_REL * abs(el.item())
    for a, b2 in zip(gp, gp2):
        ga = a.grad
        gb = mirrored(b2.grad)
        for c in range(4):
            scale = ga[:, c].abs().max().item()
            off = ((ga[:, c] - gb[:, c]).abs() > parity.ELEM_REL * scale)
            assert off.float().mean().item() < 0.01


def test_batch_shard_additivity(dev):
    """loss(B) equals the mean of the shard losses and the
    raw term sums add up -- the property the multi-GPU path relies on."""
    from oracle.make_golden import loss_config, make_inputs
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    cfg = loss_config('bayesian')
    left, right, preds = make_inputs(4, 32, 64, 0.3, 9)
    stereo = torch.cat([left, right], 1).to(dev)
    preds = [p.to(dev) for p in preds]

    def run(lo, hi, world):
        fn = L.TukraUncertaintyLoss(**cfg)
        fn.world_size = fn.grad_world_size = world
        gp = [p[lo:hi].clone().requires_grad_(True) for p in preds]
        pyr = U.scale_pyramid(stereo[lo:hi].contiguous(), 4)
        dl, el = fn(pyr, gp, U.reconstruct_pyramid(gp, pyr))
        (dl + el).backward()
        return dl, el, gp, fn.last_term_sums

    dl, el, gp, sums = run(0, 4, 1)
    a = run(0, 2, 2)
    b = run(2, 4, 2)
    assert torch.allclose(a[3] + b[3], sums, rtol=1e-6)
    assert abs((a[0] + b[0]).item() - dl.item()) < LOSS_REL * abs(dl.item())
    assert abs((a[1] + b[1]).item() - el.item()) < LOSS_REL * abs(el.item())
    for i in range(4):
        cat = torch.cat([a[2][i].grad, b[2][i].grad], 0)
        assert rel_l2(cat.cpu().numpy(), gp[i].grad.cpu().numpy()) < 1e-5


# --------------------------------------------------------------- components --
def test_component_modules_match_reference_fixture(dev):
    from uncertainty_model_b200.train import loss as L
    g = load('components.npz')
    im = torch.from_numpy(g['images']).to(dev)
    er = torch.from_numpy(g['error']).to(dev)

    def fresh(key):
        return torch.from_numpy(g[key]).to(dev).requires_grad_(True)

    from oracle import kinks as KK
    pred64 = torch.from_numpy(g['pred']).double()
    err64 = torch.from_numpy(g['error']).double()
    none4 = torch.zeros(pred64.shape, dtype=torch.bool)

    def check(val, wrt, key, mask=None):
        grad, = torch.autograd.grad(val, wrt)
        ref = float(g[f'{key}_f64'])
        assert abs(val.item() - ref) <= LOSS_REL * abs(ref), key
        assert grad_rel(grad, g[f'{key}_grad_f64'],
                        None if mask is None else mask.numpy()) < GRAD_REL, key

    for alpha in (0.85, 1.0):
        rc = fresh('recon')
        ws = L.WeightedSSIMLoss(alpha)
        e = ws.image_error(im, rc)
        assert not e.requires_grad
        assert np.allclose(e.cpu().numpy(), g[f'image_error_a{alpha}_f64'],
                           atol=1e-5)
        check(ws(im, rc), rc, f'wssim_a{alpha}')
        assert torch.equal(ws.previous_image_error, e)
    # (the gradient w.r.t. a GIVEN reconstruction has no ambiguous elements:
    #  I - recon is a difference of two inputs; the consistency terms do)
    ma, mb = KK.consistency_kinks(pred64[:, 0:2], pred64[:, 0:2])
    m = none4.clone(); m[:, 0:2] = ma | mb
    pr = fresh('pred')
    check(L.ConsistencyLoss()(pr[:, 0:2]), pr, 'cons', m)
    ma, mb = KK.consistency_kinks(pred64[:, 2:4], pred64[:, 0:2])
    m = none4.clone(); m[:, 2:4] = ma; m[:, 0:2] = mb
    pr = fresh('pred')
    check(L.ConsistencyLoss()(pr[:, 2:4], pr[:, 0:2]), pr, 'cons_ab', m)
    m = none4.clone(); m[:, 0:2] = KK.smoothness_kinks(pred64[:, 0:2])
    pr = fresh('pred')
    check(L.SmoothnessLoss()(pr[:, 0:2], im), pr, 'smooth', m)
    pool = torch.nn.functional.avg_pool2d
    for lt in ('l1', 'bayesian', 'log_bayesian'):
        for pooling in (False, True):
            pr = fresh('pred')
            fn = L.ReprojectionErrorLoss(lt, 0.7, 0.3, pooling)
            m = KK.dilate3(KK.error_term_kinks(pool(pred64, 3, 1),
                                               pool(err64, 3, 1), lt, True,
                                               True)) if pooling else \
                KK.error_term_kinks(pred64, err64, lt, True, True)
            check(fn(pr, im, er), pr,
                  f'reproj_{lt}_{"pool" if pooling else "nopool"}', m)


def test_zero_weight_terms_vanish(dev):
    from oracle.make_golden import make_inputs
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    left, right, preds = make_inputs(1, 32, 64, 0.3, 2)
    stereo = torch.cat([left, right], 1).to(dev)
    gp = [p.to(dev).requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(stereo, 4)
    fn = L.TukraUncertaintyLoss(
        wssim_weight=0.0, consistency_weight=0.0, smoothness_weight=0.0,
        predictive_error_weight=0.0,
        error_loss_config=dict(loss_type='l1', smoothness_weight=0,
                               consistency_weight=0))
    dl, el = fn(pyr, gp, U.reconstruct_pyramid(gp, pyr))
    (dl + el).backward()
    assert dl.item() == 0.0 and el.item() == 0.0
    assert all(float(p.grad.abs().max()) == 0.0 for p in gp)


class TinyDisc(torch.nn.Module):
    """Stand-in for the reference's discriminator (model/discriminator.py):
    `features(pyramid)` -> list of maps, `forward(pyramid)` -> (B,1) verdict."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.conv = torch.nn.ModuleList(
            [torch.nn.Conv2d(6, 4, 3, padding=1) for _ in range(4)])

    def features(self, pyramid):
        return [torch.tanh(c(x)) for c, x in zip(self.conv, pyramid)]

    def forward(self, pyramid):
        f = self.features(pyramid)
        return torch.sigmoid(sum(x.mean(dim=(1, 2, 3)) for x in f))[:, None]


@pytest.mark.parametrize('shape', [(2, 32, 64), (2, 256, 512)])
def test_adversarial_path_with_a_discriminator(dev, shape):
    """loss.py:552-558: generator + perceptual terms consume the materialised
    reconstructions; gradients flow through them into the predictions.
    BASELINE config 4's code path (given reconstruction + external gradient
    w.r.t. it), against the fp64 oracle at north_star's tolerances.  The small
    case runs past `perceptual_start` (generator + perceptual term); the large
    one before it (generator term only): the perceptual term is an L1 distance
    of feature maps, whose own sign kinks (|feature difference| within fp32
    rounding of 0, each touching a 3x3 patch of pixels) are the toy network's,
    not the path's, and are not in the kink mask."""
    from oracle import loss_port as P
    from oracle.make_golden import loss_config, make_inputs
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U

    cfg = loss_config('l1')
    left, right, preds = make_inputs(*shape, 0.3, 8)
    stereo = torch.cat([left, right], 1)
    disc = TinyDisc().double()

    op = [p.double().requires_grad_(True) for p in preds]
    opyr = P.pyramid(stereo.double(), 4)
    orec = P.recon_pyramid(op, opyr)
    odl, oel = P.total_loss(opyr, op, orec, cfg)
    verdict = disc(orec)
    epoch = 7 if shape[1] <= 64 else 0
    odl = odl + 0.85 * torch.nn.functional.mse_loss(
        verdict, torch.ones_like(verdict))
    if epoch >= 5:
        odl = odl + 0.05 * sum((a - b).abs().mean() for a, b in
                               zip(disc.features(opyr), disc.features(orec)))
    (odl + oel).backward()

    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False       # the conv net in true fp32
    try:
        gdisc = TinyDisc().to(dev)
        gp = [p.to(dev).requires_grad_(True) for p in preds]
        pyr = U.scale_pyramid(stereo.to(dev), 4)
        rec = U.reconstruct_pyramid(gp, pyr)
        dl, el = L.TukraUncertaintyLoss(**cfg)(pyr, gp, rec, epoch, gdisc)
        (dl + el).backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert abs(dl.item() - odl.item()) <= LOSS_REL * abs(odl.item())
    assert abs(el.item() - oel.item()) <= LOSS_REL * abs(oel.item())
    masks = parity.masks_for(stereo, preds, cfg)
    parity.check_grads([p.grad for p in gp], [p.grad for p in op], masks)


@pytest.mark.parametrize('shape', [(2, 256, 512), (1, 512, 1024)])
def test_given_reconstruction_with_external_gradient(dev, shape):
    """The adversarial step without a network in the way: the loss of the
    MATERIALISED reconstructions plus a fixed linear functional of them, so
    that the gradient arriving at the reconstruction from outside is known
    exactly (BASELINE config 4: recon materialised + grad_recon)."""
    from oracle import loss_port as P
    from oracle.make_golden import loss_config, make_inputs
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    cfg = loss_config('l1')
    left, right, preds = make_inputs(*shape, 0.3, 12)
    stereo = torch.cat([left, right], 1)
    g = torch.Generator().manual_seed(4)
    n = shape[0] * shape[1] * shape[2]
    ws = [(torch.rand(shape[0], 6, shape[1] >> i, shape[2] >> i, generator=g)
           - 0.5) * (4.0 ** i / n) for i in range(4)]

    op = [p.double().requires_grad_(True) for p in preds]
    opyr = P.pyramid(stereo.double(), 4)
    orec = P.recon_pyramid(op, opyr)
    odl, oel = P.total_loss(opyr, op, orec, cfg)
    oext = sum((r * k.double()).sum() for r, k in zip(orec, ws))
    (odl + oel + oext).backward()

    gp = [p.to(dev).requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(stereo.to(dev), 4)
    rec = list(U.reconstruct_pyramid(gp, pyr))           # materialised
    dl, el = L.TukraUncertaintyLoss(**cfg)(pyr, gp, rec, 0, None)
    ext = sum((r * k.to(dev)).sum() for r, k in zip(rec, ws))
    (dl + el + ext).backward()
    assert abs(dl.item() - odl.item()) <= LOSS_REL * abs(odl.item())
    assert abs(el.item() - oel.item()) <= LOSS_REL * abs(oel.item())
    masks = parity.masks_for(stereo, preds, cfg)
    parity.check_grads([p.grad for p in gp], [p.grad for p in op], masks)


def test_errors(dev):
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U
    with pytest.raises(ValueError, match='Loss must be either'):
        L.TukraUncertaintyLoss(error_loss_config=dict(loss_type='huber'))
    with pytest.raises(ValueError, match='no CPU fallback'):
        U.scale_pyramid(torch.rand(1, 6, 16, 16), 2)
    with pytest.raises(TypeError):
        U.scale_pyramid(torch.rand(1, 6, 16, 16, dtype=torch.float64,
                                   device=dev), 2)


def test_tensors_on_a_device_that_is_not_current(dev):
    """The reference's DDP launcher moves model, loss and data `.to(cuda:i)`
    and never calls torch.cuda.set_device (parallel_main.py:152-160): every
    rank but the first runs with tensors on a device that is not the current
    one.  The library launches where its tensors live (DeviceGuard in the C
    ABI) and leaves the caller's current device alone; results are bit-identical
    to the run on cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from oracle.make_golden import loss_config, make_inputs
    from oracle import spars_port as SP
    from uncertainty_model_b200.train import sparsification as S
    cfg = loss_config('bayesian', smoothness_weight=0.25)
    left, right, preds = make_inputs(2, 64, 128, 0.3, 91)
    stereo = torch.cat([left, right], 1)
    outs = []
    for name in ('cuda:0', 'cuda:1'):
        d = torch.device(name)
        assert torch.cuda.current_device() == 0
        dl, el, gp, pyr, rec, fn = run_ours(d, stereo, preds, cfg)
        e_map, u_map = SP.synthetic_maps(1, 40, 56, seed=2)
        curve = S.curve(e_map.to(d), u_map.to(d), device=d)
        torch.cuda.synchronize(d)
        assert torch.cuda.current_device() == 0
        assert dl.device == d and gp[0].grad.device == d
        # a stand-alone term: no image tensor names the device
        from uncertainty_model_b200.train import loss as L
        pc = preds[0].to(d).requires_grad_(True)
        (L.ConsistencyLoss()(pc[:, 0:2]) +
         L.ConsistencyLoss()(pc[:, 2:4], pc[:, 0:2])).backward()
        torch.cuda.synchronize(d)
        assert torch.cuda.current_device() == 0
        outs.append((dl.item(), el.item(),
                     [g.grad.cpu().numpy() for g in gp] + [pc.grad.cpu().numpy()],
                     [p.cpu().numpy() for p in pyr], curve.cpu().numpy()))
    a, b = outs
    assert a[0] == b[0] and a[1] == b[1]
    for x, y in zip(a[2] + a[3] + [a[4]], b[2] + b[3] + [b[4]]):
        assert np.array_equal(x, y)
