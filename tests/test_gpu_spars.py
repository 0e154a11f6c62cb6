"""GPU parity tests of the sparsification / AUSE path: pooled keys, the sort
permutation, the curves and the AUSE are bit-exact against the canonical CPU
oracle (oracle/spars_port.py) and match the reference's own fp32 outputs
(tests/golden/spars.npz) to 1e-6."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    from uncertainty_model_b200 import _lib
    _lib.lib()
    return torch.device('cuda:0')


@pytest.mark.parametrize('name', ['small', 'ties', 'ragged'])
def test_fixture_cases_bit_exact(dev, name):
    from oracle import spars_port as SP
    from uncertainty_model_b200.train import sparsification as S
    g = np.load(os.path.join(GOLDEN, 'spars.npz'))
    err_np, unc_np = g[f'{name}_err'], g[f'{name}_unc']
    err = torch.from_numpy(err_np).to(dev)
    unc = torch.from_numpy(unc_np).to(dev)

    acc, rows, parts = S.curve_sums(err, unc, return_order=True)
    f = err.shape[0]
    # pooled maps == ATen avg_pool2d bit for bit
    assert np.array_equal(parts['pooled_oracle'].cpu().numpy().reshape(f, 2, -1),
                          g[f'{name}_pooled_err'].reshape(f, 2, -1))
    assert np.array_equal(parts['pooled_pred'].cpu().numpy().reshape(f, 2, -1),
                          g[f'{name}_pooled_unc'].reshape(f, 2, -1))
    # permutation == torch argsort(descending=True, stable=True)
    assert np.array_equal(parts['order'].cpu().numpy().reshape(f, 2, -1),
                          g[f'{name}_stable_order_unc'])

    oc = S.curve(err, err, device=dev)
    pc = S.curve(err, unc, device=dev)
    ref_oc, oparts = SP.curve_canonical(err_np, err_np, return_parts=True)
    ref_pc, pparts = SP.curve_canonical(err_np, unc_np, return_parts=True)
    assert np.array_equal(acc.cpu().numpy(), pparts['row_norm_sum'])
    assert np.array_equal(oc.cpu().numpy(), ref_oc)
    assert np.array_equal(pc.cpu().numpy(), ref_pc)
    a = S.ause(oc, pc)
    assert np.float32(a.item()) == SP.ause_canonical(ref_oc, ref_pc)
    assert np.float32(S.ause(oc.cpu(), pc.cpu()).item()) == np.float32(a.item())
    assert np.float32(S.aurg(pc, oc).item()) == \
        SP.ause_canonical(ref_pc, ref_oc)
    # against the reference's own fp32 results
    assert np.allclose(oc.cpu().numpy(), g[f'{name}_oracle_curve'], rtol=2e-6)
    if name != 'ties':     # the reference's argsort is unstable on ties
        assert np.allclose(pc.cpu().numpy(), g[f'{name}_pred_curve'],
                           rtol=2e-6)
        assert abs(a.item() - float(g[f'{name}_ause'])) < 1e-6


def test_c5_shape_against_reference_anchor(dev):
    """4 frames of the SCARED-shape 1024x1280 case."""
    from oracle import spars_port as SP
    from uncertainty_model_b200.train import sparsification as S
    g = np.load(os.path.join(GOLDEN, 'spars.npz'))
    err, unc = SP.synthetic_maps(4, 1024, 1280, seed=0)
    err, unc = err.to(dev), unc.to(dev)
    oc = S.curve(err, err, device=dev)
    pc = S.curve(err, unc, device=dev)
    assert np.allclose(oc.cpu().numpy(), g['c5_4frames_oracle_curve'],
                       rtol=2e-6)
    assert np.allclose(pc.cpu().numpy(), g['c5_4frames_pred_curve'], rtol=2e-6)
    assert abs(S.ause(oc, pc).item() - float(g['c5_4frames_ause'])) < 1e-7
    # size-independent properties at full size
    acc, rows, parts = S.curve_sums(err[:1], unc[:1], return_order=True)
    order = parts['order'].long()
    keys = parts['pooled_pred']
    sorted_keys = torch.gather(keys, 1, order)
    assert bool((sorted_keys[:, 1:] <= sorted_keys[:, :-1]).all())   # sorted
    tie = sorted_keys[:, 1:] == sorted_keys[:, :-1]
    assert bool((order[:, 1:][tie] > order[:, :-1][tie]).all())      # stable
    assert bool((torch.sort(order, dim=1).values ==
                 torch.arange(order.shape[1], device=dev)).all())    # perm
    assert oc[0].item() == 1.0 and pc[0].item() == 1.0
    # the oracle ranking is the best possible one
    assert bool((oc <= pc + 1e-6).all())


def test_chunked_processing_is_identical(dev, monkeypatch):
    from oracle import spars_port as SP
    from uncertainty_model_b200.train import sparsification as S
    err, unc = SP.synthetic_maps(5, 40, 56, seed=4)
    err, unc = err.to(dev), unc.to(dev)
    full = S.curve(err, unc, device=dev)
    monkeypatch.setattr(S, 'WORKSPACE_BUDGET_BYTES', 1)   # one row at a time
    chunked = S.curve(err, unc, device=dev)
    assert torch.equal(full, chunked)


def test_random_curve_and_errors(dev):
    from uncertainty_model_b200.train import sparsification as S
    err = torch.rand(1, 2, 30, 40, device=dev)
    rc = S.random_curve(err, device=dev)
    assert rc.shape == (100,) and rc.dtype == torch.float32
    assert rc[0].item() == 1.0
    assert S.curve(err, err).device.type == 'cpu'      # reference default
    with pytest.raises(Exception, match='different step sizes'):
        S.ause(torch.zeros(3), torch.zeros(4))
    with pytest.raises(ValueError):
        S.curve(err.cpu(), err.cpu())
    with pytest.raises(ValueError):
        S.curve(err[:, :, :8, :8], err[:, :, :8, :8])


def test_evaluation_front_half_matches_oracle(dev):
    """The evaluation loop up to its metrics (reference evaluate.py:136-160,
    SURVEY.md section 8f n1): split the prediction, reconstruct both views,
    `WeightedSSIMLoss(alpha=1).image_error`, the (identity) align-corners
    resize, then oracle / predicted / random curves, AUSE and AURG -- through
    the drop-in modules on the GPU against the CPU oracle ports of the same
    lines.  The error maps agree to float rounding, so the GPU curves are
    compared (a) bit-exactly with the canonical oracle fed the GPU error map
    and (b) to tolerance with the all-CPU flow."""
    from oracle import loss_port as P
    from oracle import spars_port as SP
    from oracle.make_golden import make_inputs
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import sparsification as S
    from uncertainty_model_b200.train import utils as U
    left, right, preds = make_inputs(2, 48, 96, 0.3, 77)
    prediction = preds[0]
    images = torch.cat([left, right], 1)

    # reference flow, CPU oracle
    disparity, uncertainty = torch.split(prediction, [2, 2], dim=1)
    dl, dr = torch.split(disparity, [1, 1], dim=1)
    recon_ref = torch.cat((P.warp_to_left(dl, right), P.warp_to_right(dr, left)), 1)
    err_ref = P.image_error(images, recon_ref, alpha=1.0)
    oc_ref = SP.curve_canonical(err_ref.numpy(), err_ref.numpy())
    pc_ref = SP.curve_canonical(err_ref.numpy(), uncertainty.contiguous().numpy())

    # the same lines through the drop-in modules
    g_images, g_pred = images.to(dev), prediction.to(dev)
    g_disp, g_unc = torch.split(g_pred, [2, 2], dim=1)
    g_dl, g_dr = torch.split(g_disp, [1, 1], dim=1)
    g_left, g_right = left.to(dev), right.to(dev)
    recon = torch.cat((U.reconstruct_left_image(g_dl, g_right),
                       U.reconstruct_right_image(g_dr, g_left)), dim=1)
    assert np.allclose(recon.cpu().numpy(), recon_ref.numpy(), atol=2e-6)
    error = L.WeightedSSIMLoss(alpha=1).image_error(g_images, recon)
    assert error.shape == (2, 2, 48, 96)
    assert np.allclose(error.cpu().numpy(), err_ref.numpy(), atol=1e-5)
    oc = S.curve(error, error, device=dev)
    pc = S.curve(error, g_unc, device=dev)
    rc = S.random_curve(error, device=dev)
    ause, aurg = S.ause(oc, pc), S.aurg(pc, rc)

    e_np = error.cpu().numpy()
    u_np = g_unc.contiguous().cpu().numpy()
    assert np.array_equal(oc.cpu().numpy(), SP.curve_canonical(e_np, e_np))
    assert np.array_equal(pc.cpu().numpy(), SP.curve_canonical(e_np, u_np))
    assert np.allclose(oc.cpu().numpy(), oc_ref, rtol=1e-4)
    assert np.allclose(pc.cpu().numpy(), pc_ref, rtol=1e-3, atol=1e-5)
    assert np.float32(ause.item()) == SP.ause_canonical(
        oc.cpu().numpy(), pc.cpu().numpy())
    assert np.isfinite(aurg.item()) and rc.shape == (100,)
    assert bool((oc <= pc + 1e-6).all())


@pytest.mark.parametrize('k', [3, 7, 15])
def test_other_kernel_sizes_bit_exact(dev, k):
    """Window sizes other than the default 11 run the run-time-size pooling
    kernel; pooled maps and curves stay bit-exact against the canonical oracle
    (ragged shape: neither dimension a multiple of the thread patch)."""
    from oracle import spars_port as SP
    from uncertainty_model_b200.train import sparsification as S
    err, unc = SP.synthetic_maps(2, 37, 61, seed=3 + k)
    acc, rows, parts = S.curve_sums(err.to(dev), unc.to(dev), kernel_size=k,
                                    return_order=True)
    e, u = err.numpy(), unc.numpy()
    pe = SP.pool_exact(e, k).reshape(2, 2, -1)
    pu = SP.pool_exact(u, k).reshape(2, 2, -1)
    assert np.array_equal(parts['pooled_oracle'].cpu().numpy().reshape(2, 2, -1), pe)
    assert np.array_equal(parts['pooled_pred'].cpu().numpy().reshape(2, 2, -1), pu)
    pc = S.curve(err.to(dev), unc.to(dev), kernel_size=k, device=dev)
    assert np.array_equal(pc.cpu().numpy(), SP.curve_canonical(e, u, k))


# ----------------------------------------------------- frames over 2 GPUs ----
def _sharded_worker(rank, world, port, tmp):
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import spars_port as SP
    from uncertainty_model_b200 import distributed as D
    from uncertainty_model_b200.train import sparsification as S
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    d = torch.device('cuda', rank)
    torch.cuda.set_device(d)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=d)
    try:
        err, unc = SP.synthetic_maps(6, 56, 72, seed=4)
        lo, hi = D.shard_bounds(6, rank, world)
        out = {}
        for name, a, c in (('oracle', err, err), ('pred', err, unc)):
            shard = D.sharded_curve(a[lo:hi].to(d), c[lo:hi].to(d), device=d)
            single = S.curve(a.to(d), c.to(d), device=d)
            out[name] = (shard.cpu(), single.cpu())
        torch.save(out, os.path.join(tmp, f'rank{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_sharded_curve_over_two_gpus_is_the_single_gpu_curve(dev, tmp_path):
    """north_star: frames shard over ranks, the 100-entry fp64 curve sums are
    all-reduced over NCCL; the curve every rank ends up with is bit for bit
    the single-GPU curve (distributed.sharded_curve)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import socket
    import torch.multiprocessing as mp
    from oracle import spars_port as SP
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    err, unc = SP.synthetic_maps(6, 56, 72, seed=4)
    want = {'oracle': SP.curve_canonical(err.numpy(), err.numpy()),
            'pred': SP.curve_canonical(err.numpy(), unc.numpy())}
    for rank in range(2):
        got = torch.load(tmp_path / f'rank{rank}.pt')
        for name in ('oracle', 'pred'):
            shard, single = got[name]
            assert torch.equal(shard, single), (rank, name)
            assert np.array_equal(single.numpy(), want[name]), (rank, name)
