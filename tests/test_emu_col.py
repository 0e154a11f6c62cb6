"""CPU: the column-marching kernel logic (csrc/col_core.cuh), executed by the
emulation harness, against the oracle port -- forward sums, error maps,
reconstructions and gradients -- over unit shapes that exercise every kernel
mode: full-row units of both views (plain and masked), one view per unit,
column tiles with halos, several row strips."""
import numpy as np
import pytest
import torch

from oracle.make_golden import loss_config, make_inputs

from emu_harness import emu_col_scale
from test_emu import oracle_scale, settings_from

CASES = [
    # (b, h, w, scale, loss cfg, (maxT, R, consR))
    (1, 24, 32, 0.3, loss_config('l1'), (512, 32, 16)),          # both views, plain
    (2, 20, 32, 0.5, loss_config('bayesian', smoothness_weight=0.6,
                                 consistency_weight=0.8), (32, 8, 5)),  # 1 view
    (1, 33, 52, 1.0, loss_config('log_bayesian', smoothness_weight=0.4),
     (512, 8, 4)),                                               # masked, odd h
    (1, 16, 70, 0.3, loss_config('l1'), (36, 16, 16)),           # column tiles
    (2, 21, 37, 0.5, loss_config('bayesian', smoothness_weight=0.6,
                                 consistency_weight=0.8), (24, 6, 5)),  # tiles
]


@pytest.mark.parametrize('case', range(len(CASES)))
def test_emulated_column_kernels_match_oracle(case):
    b, h, w, scale, cfg, (maxT, R, consR) = CASES[case]
    left, right, preds = make_inputs(b, h, w, scale, 40 + case)
    images, pred = torch.cat([left, right], 1), preds[0]
    i = case % 3
    st = settings_from(cfg)
    coefs = st.coefs(i, b * h * w)
    g = (0.7, 1.3)
    out = emu_col_scale(st, st.terms(), coefs, images, pred, g=g, maxT=maxT,
                        R=R, consR=consR, want_recon=True)

    dl, el, pr, err, rec = oracle_scale(images, pred, cfg, i)
    (g[0] * dl + g[1] * el).backward()

    for key in ('sums', 'sums_grad'):
        sums = out[key]
        mine_dl = sum(coefs[k] * sums[k] for k in range(3))
        mine_el = sum(coefs[k] * sums[k] for k in range(3, 6))
        assert mine_dl == pytest.approx(float(dl.detach()), rel=2e-6), key
        assert mine_el == pytest.approx(float(el.detach()), rel=2e-6), key
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
