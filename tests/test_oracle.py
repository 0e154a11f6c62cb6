"""CPU: the oracle ports against the fixtures generated from the real
reference (oracle/make_golden.py), and -- when /root/reference is mounted --
against the reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_port as P
from oracle import spars_port as SP
from oracle.make_golden import loss_config, make_inputs
from oracle.ref_shim import reference_available

from conftest import GOLDEN

LOSS_CASES = {
    'l1_default': loss_config('l1'),
    'bayesian_default': loss_config('bayesian'),
    'log_bayesian_scale1': loss_config('log_bayesian'),
    'l1_allterms': loss_config('l1', smoothness_weight=0.6,
                               consistency_weight=0.8),
    'bayesian_pooling': loss_config('bayesian', smoothness_weight=0.6,
                                    consistency_weight=0.8, pooling=True),
    'bayesian_smooth': loss_config('bayesian'),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize('name', sorted(LOSS_CASES))
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_loss_port_matches_reference_fixture(name, tag):
    g = load(f'loss_{name}.npz')
    dt = torch.float32 if tag == 'f32' else torch.float64
    stereo = torch.cat([torch.from_numpy(g['left']),
                        torch.from_numpy(g['right'])], 1).to(dt)
    preds = [torch.from_numpy(g[f'pred{i}']).to(dt) for i in range(4)]
    dl, el, grads = P.step(stereo, preds, LOSS_CASES[name])
    # same ATen ops in the same order -> identical up to reduction order
    tol = 1e-6 if tag == 'f32' else 1e-12
    assert abs(float(dl) - float(g[f'disp_loss_{tag}'])) \
        <= tol * abs(float(g[f'disp_loss_{tag}']))
    assert abs(float(el) - float(g[f'error_loss_{tag}'])) \
        <= tol * abs(float(g[f'error_loss_{tag}']))
    for i in range(4):
        ref = g[f'grad{i}_{tag}']
        assert np.allclose(grads[i].numpy(), ref, rtol=0,
                           atol=tol * 10 * np.abs(ref).max())
    if tag == 'f32':
        pyr = P.pyramid(stereo, 4)
        rec = P.recon_pyramid(preds, pyr)
        for i in range(4):
            assert np.array_equal(pyr[i].numpy(), g[f'pyr{i}_f32'])
            assert np.array_equal(rec[i].numpy(), g[f'rec{i}_f32'])
            err = P.image_error(pyr[i], rec[i])
            assert np.allclose(err.numpy(), g[f'err{i}_f32'], atol=1e-7)


def test_component_ports_match_reference_fixture():
    g = load('components.npz')
    for tag, dt, tol in (('f32', torch.float32, 2e-6),
                         ('f64', torch.float64, 1e-12)):
        im = torch.from_numpy(g['images']).to(dt)
        rc = torch.from_numpy(g['recon']).to(dt).requires_grad_(True)
        pr = torch.from_numpy(g['pred']).to(dt).requires_grad_(True)
        er = torch.from_numpy(g['error']).to(dt)

        def check(val, wrt, key):
            grad, = torch.autograd.grad(val, wrt)
            assert np.allclose(val.detach().numpy(), g[f'{key}_{tag}'],
                               rtol=tol, atol=0), key
            ref = g[f'{key}_grad_{tag}']
            assert np.allclose(grad.numpy(), ref, rtol=0,
                               atol=10 * tol * np.abs(ref).max()), key

        for alpha in (0.85, 1.0):
            e = P.image_error(im, rc, alpha)
            assert np.allclose(e.detach().numpy(),
                               g[f'image_error_a{alpha}_{tag}'], atol=tol)
            check((e[:, 0:1] + e[:, 1:2]).mean(), rc, f'wssim_a{alpha}')
        check(P.consistency(pr[:, 0:2]), pr, 'cons')
        check(P.consistency(pr[:, 2:4], pr[:, 0:2]), pr, 'cons_ab')
        check(P.smoothness(pr[:, 0:2], im), pr, 'smooth')
        for lt in P.LOSS_TYPES:
            for pooling in (False, True):
                key = f'reproj_{lt}_{"pool" if pooling else "nopool"}'
                check(P.uncertainty_loss(pr, im, er, lt, 0.7, 0.3, pooling),
                      pr, key)
        assert np.allclose(P.warp_to_left(pr[:, 0:1].detach(), im[:, 3:6]),
                           g[f'recon_left_{tag}'], atol=tol)
        assert np.allclose(P.warp_to_right(pr[:, 1:2].detach(), im[:, 0:3]),
                           g[f'recon_right_{tag}'], atol=tol)


def test_loss_type_validation():
    with pytest.raises(ValueError, match='Loss must be either'):
        P.uncertainty_loss(torch.rand(1, 4, 8, 8), torch.rand(1, 6, 8, 8),
                           torch.rand(1, 2, 8, 8), loss_type='l2')


def test_anchor_c1_full_size():
    """Appendix C / anchors.npz: config-1 shape, seed 0, reference fp32."""
    g = load('anchors.npz')
    b, h, w = g['c1_shape']
    left, right, preds = make_inputs(int(b), int(h), int(w), 0.3, 0)
    dl, el, grads = P.step(torch.cat([left, right], 1), preds,
                           loss_config('l1'))
    assert abs(float(dl) - float(g['c1_disp_loss'])) < 1e-5
    assert abs(float(el) - float(g['c1_error_loss'])) < 1e-5
    l2 = np.array([float(x.double().norm()) for x in grads])
    assert np.allclose(l2, g['c1_grad_l2'], rtol=1e-5)


# ---------------------------------------------------------------- spars ----
def test_pool_exact_is_bit_identical_to_aten():
    g = load('spars.npz')
    for name in ('small', 'ties', 'ragged'):
        for m in ('err', 'unc'):
            mine = SP.pool_exact(g[f'{name}_{m}'], 11)
            assert np.array_equal(mine, g[f'{name}_pooled_{m}']), (name, m)


def test_stable_order_matches_torch_stable_argsort():
    g = load('spars.npz')
    for name in ('small', 'ties', 'ragged'):
        p = SP.pool_exact(g[f'{name}_unc'], 11)
        rows = p.shape[0] * 2
        order = SP.stable_order(p.reshape(rows, -1))
        assert np.array_equal(order.reshape(p.shape[0], 2, -1),
                              g[f'{name}_stable_order_unc']), name


def test_curves_match_reference_fixture():
    g = load('spars.npz')
    for name in ('small', 'ties', 'ragged'):
        err, unc = g[f'{name}_err'], g[f'{name}_unc']
        oc = SP.curve_canonical(err, err)
        pc = SP.curve_canonical(err, unc)
        # the reference sums in fp32 in ATen's order: 1e-6 relative
        assert np.allclose(oc, g[f'{name}_oracle_curve'], rtol=2e-6, atol=0)
        if name != 'ties':   # unstable argsort in the reference on ties
            assert np.allclose(pc, g[f'{name}_pred_curve'], rtol=2e-6, atol=0)
            assert abs(float(SP.ause_canonical(oc, pc))
                       - float(g[f'{name}_ause'])) < 1e-6
        t_oc = SP.curve_reference_style(torch.from_numpy(err),
                                        torch.from_numpy(err))
        t_pc = SP.curve_reference_style(torch.from_numpy(err),
                                        torch.from_numpy(unc))
        assert np.allclose(t_oc.numpy(), g[f'{name}_oracle_curve'],
                           rtol=1e-6)
        assert np.allclose(t_pc.numpy(), pc, rtol=2e-6)


def test_ause_length_check():
    with pytest.raises(Exception, match='different step sizes'):
        SP.ause_canonical(np.zeros(3, np.float32), np.zeros(4, np.float32))
    with pytest.raises(Exception, match='different step sizes'):
        SP.ause_reference_style(torch.zeros(3), torch.zeros(4))


def test_cut_points_follow_the_float_formula():
    """sparsification.py:26-27 uses Python float arithmetic; it is NOT always
    step*N//100 (e.g. N=1380), so the host computes the cuts the same way."""
    n = (1024 - 10) * (1280 - 10)
    cuts = SP.cut_points(n)
    assert all(int(cuts[s]) == s * n // 100 for s in range(100))
    for h, w in ((256, 512), (192, 384), (40, 56), (23, 31)):
        n = (h - 10) * (w - 10)
        cuts = SP.cut_points(n)
        assert all(int(cuts[s]) == int(s / 100 * n) for s in range(100))
        assert cuts[-1] == n and np.all(np.diff(cuts) >= 0)


@pytest.mark.skipif(not reference_available(),
                    reason='reference tree not mounted')
def test_port_against_live_reference():
    from oracle.ref_shim import import_reference
    L, u, S = import_reference()
    left, right, preds = make_inputs(1, 24, 40, 0.7, 9)
    stereo = torch.cat([left, right], 1)
    cfg = loss_config('bayesian', smoothness_weight=0.3)
    dl, el, grads = P.step(stereo, preds, cfg)
    pr = [p.clone().requires_grad_(True) for p in preds]
    pyr = u.scale_pyramid(stereo, 4)
    rdl, rel = L.TukraUncertaintyLoss(**cfg)(
        pyr, pr, u.reconstruct_pyramid(pr, pyr), 0, None)
    (rdl + rel).backward()
    assert float(dl) == pytest.approx(float(rdl), rel=1e-6)
    assert float(el) == pytest.approx(float(rel), rel=1e-6)
    for a, b in zip(grads, pr):
        assert torch.allclose(a, b.grad, atol=1e-7)


# ------------------------------------------------------------- kink masks --
@pytest.mark.parametrize('name', sorted(LOSS_CASES))
def test_kink_mask_accounts_for_fp32_vs_fp64_of_the_reference(name):
    """oracle/kinks.py, the explicit mask of the parity metric (tests/parity.py):
    the REAL reference's fp32 gradients (fixture) agree with its fp64 gradients
    element by element outside the mask -- every element that differs by more
    than rounding is one where the loss is one-sided."""
    import parity
    g = load(f'loss_{name}.npz')
    cfg = LOSS_CASES[name]
    stereo = torch.cat([torch.from_numpy(g['left']),
                        torch.from_numpy(g['right'])], 1)
    preds = [torch.from_numpy(g[f'pred{i}']) for i in range(4)]
    ref = parity.oracle_reference(stereo, preds, cfg)
    parity.check_grads([g[f'grad{i}_f32'] for i in range(4)],
                       [g[f'grad{i}_f64'] for i in range(4)], ref['masks'],
                       name, mask_max=ref['mask_max'])


def test_kink_mask_margins_and_explicit_warp():
    """The tap-by-tap warp behind the masks equals the oracle's grid_sample,
    and the coordinate margin covers the fp32 evaluation with room to spare."""
    import parity
    from oracle import kinks as K
    cfg = loss_config('l1')
    left, right, preds = make_inputs(2, 128, 512, 0.3, 0)
    stereo = torch.cat([left, right], 1)
    p64 = preds[0].double()
    im64 = stereo.double()
    ew = K.explicit_warp(-p64[:, 0:1], im64[:, 3:6])
    assert torch.allclose(ew['out'], P.warp_to_left(p64[:, 0:1], im64[:, 3:6]),
                          atol=1e-14)
    ix32 = K.explicit_warp(-preds[0][:, 0:1], stereo[:, 3:6])['ix']
    worst = float((ix32.double() - ew['ix']).abs().max())
    assert worst <= 0.2 * K.IX_EPS_PER_W * 512, worst
    # the fp32 oracle passes the metric the CUDA path is held to
    ref = parity.oracle_reference(stereo, preds, cfg)
    _, _, g32 = P.step(stereo, preds, cfg)
    stats = parity.check_grads(g32, ref['grads'], ref['masks'])
    # ... and would not pass untrimmed: the masked elements carry the error
    assert max(s['full'] for s in stats) > 10 * parity.GRAD_REL
    assert max(s['masked'] for s in stats) < 2e-3


# ------------------------------------------------- post-processing ports ----
def test_combine_disparity_port_matches_reference_fixture():
    from oracle import post_port as PP
    g = load('post.npz')
    out = PP.combine_disparity(torch.from_numpy(g['left']),
                               torch.from_numpy(g['right']))
    assert out.dtype == torch.float64
    assert np.array_equal(out.numpy(), g['combined'])


def test_heatmap_port_indexing_rule():
    """matplotlib is absent here: the restated Colormap.__call__ rule on the
    values that exercise every branch of it."""
    from oracle import post_port as PP
    table = np.arange(256 * 3, dtype=np.float64).reshape(256, 3)
    x = torch.tensor([[[0.0, 1.0, -0.1, 2.0, float('nan'), 0.5, 255.0 / 256]]])
    out = PP.to_heatmap(x, table)[0].numpy().ravel()      # R channel = 3 * index
    assert list(out) == [0.0, 765.0, 0.0, 765.0, 0.0, 384.0, 765.0]


@pytest.mark.parametrize('loss_type', ['l1', 'bayesian', 'log_bayesian'])
def test_mirror_and_view_swap_symmetry_of_the_port(loss_type):
    """The property tests/test_gpu_loss.py::test_mirror_and_view_swap_symmetry
    checks on the GPU at the benchmark shape, established here on the fp64
    port: mirroring the images left-right and swapping the views leaves both
    losses unchanged and mirrors + swaps the gradients."""
    import torch
    from oracle import loss_port as P
    from oracle.make_golden import loss_config, make_inputs
    cfg = loss_config(loss_type, smoothness_weight=0.25)
    left, right, preds = make_inputs(2, 32, 64, 0.3, 7)

    def mirrored(p):
        return torch.stack([p[:, 1].flip(2), p[:, 0].flip(2),
                            p[:, 3].flip(2), p[:, 2].flip(2)], 1)

    preds = [p.double() for p in preds]
    dl, el, g = P.step(torch.cat([left, right], 1).double(), preds, cfg)
    dl2, el2, g2 = P.step(torch.cat([right.flip(3), left.flip(3)], 1).double(),
                          [mirrored(p) for p in preds], cfg)
    assert abs(float(dl) - float(dl2)) <= 1e-7 * abs(float(dl))
    assert abs(float(el) - float(el2)) <= 1e-7 * abs(float(el))
    for a, b in zip(g, g2):
        assert float((a - mirrored(b)).norm() / a.norm()) <= 1e-5
