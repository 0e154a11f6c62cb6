"""CPU: host-side logic of the drop-in -- constructor contract, coefficient
folding, lazy reconstruction pyramid, sharding helpers and the world_size-2
(gloo) reduction path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

from uncertainty_model_b200 import distributed as D
from uncertainty_model_b200.functional import LossSettings
from uncertainty_model_b200.train import loss as L
from uncertainty_model_b200.train import sparsification as S
from uncertainty_model_b200.train import utils as U


def test_constructor_contract_matches_config_yml():
    import yaml
    cfg = yaml.safe_load('''
loss:
  wssim_weight: 1.0
  consistency_weight: 1.0
  smoothness_weight: 1.0
  adversarial_weight: 0.85
  perceptual_weight: 0.05
  predictive_error_weight: 1.0
  wssim_alpha: 0.85
  perceptual_start: 5
  adversarial_loss_type: mse
  error_loss_config:
    loss_type: bayesian
    smoothness_weight: 0
    consistency_weight: 0.5
    pooling: false
''')
    fn = L.TukraUncertaintyLoss(**cfg['loss']).to('cpu')
    assert isinstance(fn, torch.nn.Module)
    assert len(list(fn.parameters())) == 0 and len(fn.state_dict()) == 0
    assert fn.predictive_error.loss_type == 'bayesian'
    assert fn.wssim.alpha == 0.85 and fn.wssim.k1 == 0.01 ** 2
    st = fn._settings()
    assert st.err_consistency_weight == 0.5 and st.err_smoothness_weight == 0
    # defaults of the reference constructors (loss.py:25-26,357-360,438-447)
    d = L.TukraUncertaintyLoss()
    assert (d.wssim_weight, d.adversarial_weight, d.perceptual_weight,
            d.perceptual_start) == (1.0, 0.85, 0.05, 5)
    e = L.ReprojectionErrorLoss()
    assert (e.loss_type, e.smoothness_weight, e.consistency_weight,
            e.pooling) == ('l1', 1.0, 1.0, False)
    assert isinstance(d.adversarial.adversarial, torch.nn.MSELoss)
    assert isinstance(L.GeneratorLoss('bce').adversarial, torch.nn.BCELoss)


def test_loss_type_error_message():
    with pytest.raises(ValueError) as e:
        L.ReprojectionErrorLoss(loss_type='huber')
    assert str(e.value) == ('Loss must be either "l1", "bayesian" '
                            'or "log_bayesian".')


def test_cpu_tensors_are_rejected_not_silently_computed():
    x = torch.rand(1, 6, 16, 32)
    with pytest.raises(ValueError, match='no CPU fallback'):
        U.scale_pyramid(x, 4)
    with pytest.raises(ValueError, match='no CPU fallback'):
        U.reconstruct_left_image(torch.rand(1, 1, 16, 32), x[:, :3])
    with pytest.raises(ValueError, match='no CPU fallback'):
        L.SmoothnessLoss()(torch.rand(1, 2, 16, 32), x)
    with pytest.raises(ValueError, match='no CPU fallback'):
        S.curve(torch.rand(1, 2, 16, 32), torch.rand(1, 2, 16, 32))


def test_coefficients_fold_weights_and_means():
    st = LossSettings(wssim_weight=2.0, consistency_weight=3.0,
                      smoothness_weight=5.0, predictive_error_weight=7.0,
                      loss_type='log_bayesian', err_smoothness_weight=0.5,
                      err_consistency_weight=0.25)
    n = 4 * 10 * 20
    c = st.coefs(2, n)
    assert c == pytest.approx([2 / n, 3 / n, 5 / (n * 4), 7 * 0.5 / (2 * n),
                               7 * 0.5 / n, 7 * 0.25 / n])
    assert LossSettings(loss_type='l1').coefs(0, n)[3] == \
        pytest.approx(1 / (2 * n))
    assert LossSettings(err_smoothness_weight=0,
                        err_consistency_weight=0).terms() == 1 | 2 | 4 | 8
    assert LossSettings().terms() == 63


def test_recon_pyramid_is_lazy_and_tracks_its_sources():
    preds = [torch.rand(1, 4, 8 >> i, 16 >> i) for i in range(3)]
    pyr = [torch.rand(1, 6, 8 >> i, 16 >> i) for i in range(3)]
    rec = U.reconstruct_pyramid(preds, pyr)
    assert len(rec) == 3 and not rec.materialised
    assert rec.built_from(preds, pyr)
    assert not rec.built_from([p.clone() for p in preds], pyr)
    assert not rec.built_from(preds[:2], pyr[:2])
    with pytest.raises(ValueError, match='no CPU fallback'):
        rec[0]        # materialising needs the GPU


def test_cut_points_and_curve_algebra():
    assert S.cut_points(1380, 100)[:3] == [0, 13, 27]
    assert S.cut_points(1380, 100)[-1] == 1380
    assert S.cut_points(1380, 100)[29] == int(29 / 100 * 1380)
    o = torch.linspace(1, 0.5, 100)
    p = o + 0.01
    assert torch.allclose(S.error(o, p), torch.full((100,), 0.01))
    assert S.ause(o, p).item() == pytest.approx(0.01, rel=1e-5)
    assert S.aurg(p, o).item() == pytest.approx(-0.01, rel=1e-5)
    with pytest.raises(Exception, match='different step sizes'):
        S.ause(o, p[:50])


def test_shard_bounds_cover_the_batch():
    for total, world in ((64, 8), (64, 3), (5, 8), (16, 1)):
        spans = [D.shard_bounds(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        D.shard_bounds(4, 4, 4)


def _gloo_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from emu_harness import emu_scale
    from oracle.make_golden import make_inputs
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        b, h, w = 4, 20, 36
        left, right, preds = make_inputs(b, h, w, 0.4, 77)
        stereo, pred = torch.cat([left, right], 1), preds[0]
        lo, hi = D.shard_bounds(b, rank, world)
        st = LossSettings(loss_type='bayesian', err_smoothness_weight=0.5)
        coefs = st.coefs(0, b * h * w)          # GLOBAL batch normaliser
        out = emu_scale(st, st.terms(), coefs, stereo[lo:hi], pred[lo:hi],
                        g=(1.0, 1.0), TW=16, R=8)
        sums = torch.tensor([out['sums']], dtype=torch.float64)
        D.reduce_term_sums(sums)
        dl, el = D.combine_terms(sums, torch.tensor([coefs]))
        torch.save(dict(dl=float(dl), el=float(el), lo=lo, hi=hi,
                        grad=out['grad_pred']),
                   os.path.join(tmp, f'rank{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_sharded_loss_world_size_2_gloo(tmp_path):
    """Two ranks, each with half the batch: all-reduced term sums give the
    full-batch loss and the concatenated shard gradients are the full-batch
    gradient (per-rank kernels emulated on the CPU, reduction over gloo)."""
    from emu_harness import emu_scale
    from oracle.make_golden import make_inputs
    import socket
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / 'rank0.pt')
    r1 = torch.load(tmp_path / 'rank1.pt')
    b, h, w = 4, 20, 36
    left, right, preds = make_inputs(b, h, w, 0.4, 77)
    stereo, pred = torch.cat([left, right], 1), preds[0]
    st = LossSettings(loss_type='bayesian', err_smoothness_weight=0.5)
    coefs = st.coefs(0, b * h * w)
    full = emu_scale(st, st.terms(), coefs, stereo, pred, g=(1.0, 1.0),
                     TW=16, R=8)
    dl = sum(coefs[k] * full['sums'][k] for k in range(3))
    el = sum(coefs[k] * full['sums'][k] for k in range(3, 6))
    for r in (r0, r1):
        assert r['dl'] == pytest.approx(dl, rel=1e-6)
        assert r['el'] == pytest.approx(el, rel=1e-6)
    assert (r0['lo'], r0['hi'], r1['lo'], r1['hi']) == (0, 2, 2, 4)
    cat = torch.cat([r0['grad'], r1['grad']], 0)
    assert torch.allclose(cat, full['grad_pred'], rtol=1e-5, atol=1e-9)
