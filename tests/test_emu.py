"""CPU: the fused-loss kernel logic (csrc/loss_core.cuh, csrc/cons_core.cuh),
executed by the emulation harness, against the oracle port -- forward sums,
error maps and gradients, over several tilings so that strip/tile seams and
ring reuse are exercised."""
import numpy as np
import pytest
import torch

from oracle import loss_port as P
from oracle.make_golden import loss_config, make_inputs
from uncertainty_model_b200.functional import LossSettings

from emu_harness import emu_scale


def settings_from(cfg):
    e = cfg['error_loss_config']
    return LossSettings(
        wssim_weight=cfg['wssim_weight'],
        consistency_weight=cfg['consistency_weight'],
        smoothness_weight=cfg['smoothness_weight'],
        predictive_error_weight=cfg['predictive_error_weight'],
        alpha=cfg['wssim_alpha'], loss_type=e['loss_type'],
        err_smoothness_weight=e['smoothness_weight'],
        err_consistency_weight=e['consistency_weight'])


def oracle_scale(images, pred, cfg, i, dtype=torch.float64):
    """One scale of loss.py:541-550 in the oracle, with autograd."""
    im = images.to(dtype)
    pr = pred.to(dtype).clone().requires_grad_(True)
    rec = P.recon_pyramid([pr], [im])[0]
    dl, el, terms = P.total_loss([im], [pr], [rec], cfg, return_terms=True)
    # total_loss treats its single scale as scale 0: rescale the smoothness
    dl = dl - terms['smooth'] * cfg['smoothness_weight'] * (1 - 0.5 ** i)
    return dl, el, pr, terms['errors'][0], rec


CASES = [
    # (b, h, w, scale, loss cfg, tiling (TW, R, consR))
    (1, 24, 40, 0.3, loss_config('l1'), (256, 32, 16)),
    (2, 21, 37, 0.5, loss_config('bayesian', smoothness_weight=0.6,
                                 consistency_weight=0.8), (16, 8, 5)),
    (1, 33, 52, 1.0, loss_config('log_bayesian', smoothness_weight=0.4),
     (20, 7, 4)),
    (1, 16, 70, 0.3, loss_config('l1'), (33, 16, 16)),
]


@pytest.mark.parametrize('case', range(len(CASES)))
def test_emulated_kernels_match_oracle(case):
    b, h, w, scale, cfg, (TW, R, consR) = CASES[case]
    left, right, preds = make_inputs(b, h, w, scale, 20 + case)
    images, pred = torch.cat([left, right], 1), preds[0]
    i = case % 3
    st = settings_from(cfg)
    coefs = st.coefs(i, b * h * w)
    g = (0.7, 1.3)
    out = emu_scale(st, st.terms(), coefs, images, pred, g=g, TW=TW, R=R,
                    consR=consR, want_recon=True)

    dl, el, pr, err, rec = oracle_scale(images, pred, cfg, i)
    (g[0] * dl + g[1] * el).backward()

    sums = out['sums']
    mine_dl = sum(coefs[k] * sums[k] for k in range(3))
    mine_el = sum(coefs[k] * sums[k] for k in range(3, 6))
    assert mine_dl == pytest.approx(float(dl.detach()), rel=2e-6)
    assert mine_el == pytest.approx(float(el.detach()), rel=2e-6)
    assert torch.isfinite(out['recon']).all()
    assert np.allclose(out['recon'].numpy(), rec.detach().numpy(), atol=2e-5)
    assert np.allclose(out['err'].numpy(), err.detach().numpy(), atol=1e-5)

    ref = pr.grad.numpy()
    got = out['grad_pred'].double().numpy()
    assert np.isfinite(got).all()
    for ch in range(4):
        num = np.linalg.norm(got[:, ch] - ref[:, ch])
        den = np.linalg.norm(ref[:, ch])
        assert num <= 1e-4 * den, (ch, num / den)


def test_emulated_given_recon_and_error_modes():
    """recon_in (WeightedSSIMLoss / adversarial path) and err_in
    (ReprojectionErrorLoss stand-alone) against the oracle."""
    from uncertainty_model_b200._lib import (TERM_CONS_U, TERM_REPROJ,
                                             TERM_SMOOTH_U, TERM_UNC)
    b, h, w = 2, 19, 45
    g = torch.Generator().manual_seed(5)
    images = torch.rand(b, 6, h, w, generator=g)
    recon = torch.rand(b, 6, h, w, generator=g)
    pred = 0.5 * torch.sigmoid(torch.randn(b, 4, h, w, generator=g))
    err = torch.rand(b, 2, h, w, generator=g)
    n = b * h * w

    # WeightedSSIMLoss(images, recon): mean(err_L + err_R)
    st = LossSettings()
    coefs = [1.0 / n, 0, 0, 0, 0, 0]
    out = emu_scale(st, TERM_REPROJ, coefs, images, pred, recon=recon,
                    g=(1.0, 0.0), TW=17, R=6)
    rc = recon.double().requires_grad_(True)
    e = P.image_error(images.double(), rc)
    val = (e[:, 0:1] + e[:, 1:2]).mean()
    val.backward()
    assert coefs[0] * out['sums'][0] == pytest.approx(float(val), rel=2e-6)
    assert np.allclose(out['err'].numpy(), e.detach().numpy(), atol=1e-5)
    ref = rc.grad.numpy()
    got = out['grad_recon'].double().numpy()
    assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref)

    # ReprojectionErrorLoss(pred, images, err) for each loss type
    for lt in ('l1', 'bayesian', 'log_bayesian'):
        st = LossSettings(loss_type=lt, err_smoothness_weight=0.7,
                          err_consistency_weight=0.3)
        coefs = st.coefs(0, n)
        coefs[0] = coefs[1] = coefs[2] = 0.0
        terms = TERM_UNC | TERM_SMOOTH_U | TERM_CONS_U
        out = emu_scale(st, terms, coefs, images, pred, err=err,
                        g=(0.0, 1.0), TW=23, R=5, consR=3)
        pr = pred.double().requires_grad_(True)
        val = P.uncertainty_loss(pr, images.double(), err.double(), lt, 0.7,
                                 0.3, False)
        val.backward()
        mine = sum(coefs[k] * out['sums'][k] for k in range(3, 6))
        assert mine == pytest.approx(float(val), rel=2e-6), lt
        ref = pr.grad.numpy()
        got = out['grad_pred'].double().numpy()
        assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref), lt
