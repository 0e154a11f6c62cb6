"""Throughput of the evaluation front half (reference evaluate.py:136-160
without the torchmetrics SSIM: reconstruct both views, DSSIM error map,
oracle / predicted / random sparsification curves, AUSE, AURG) at the
config-5 frame shape on one GPU, stage by stage, with the CPU oracle flow
beside it.

    python tools/eval_bench.py [--frames 8] [--cpu-frames 1]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import sparsification as S  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def flow(left, right, prediction, dev, ssim_loss, stages=None):
    def mark(name):
        if stages is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            stages.append((name, e))
    mark('start')
    images = torch.cat([left, right], dim=1)
    disparity, uncertainty = torch.split(prediction, [2, 2], dim=1)
    dl, dr = torch.split(disparity, [1, 1], dim=1)
    recon = torch.cat((U.reconstruct_left_image(dl, right),
                       U.reconstruct_right_image(dr, left)), dim=1)
    mark('reconstruct')
    error = ssim_loss.image_error(images, recon)
    mark('image_error')
    oc = S.curve(error, error, device=dev)
    mark('oracle_curve')
    pc = S.curve(error, uncertainty, device=dev)
    mark('predicted_curve')
    rc = S.random_curve(error, device=dev)
    mark('random_curve')
    out = S.ause(oc, pc), S.aurg(pc, rc)
    mark('ause_aurg')
    return out


def fused_flow(left, right, prediction, dev):
    """The same lines as one call (uncertainty_model_b200.train.evaluate):
    reconstructions + error map from one launch of the column kernels, plus the
    Gaussian SSIM metric evaluate.py:142-146 asks torchmetrics for."""
    from uncertainty_model_b200.train import evaluate as E
    out = E.evaluate_batch(left, right, prediction, device=dev)
    return out['ause'], out['aurg'], out['left_ssim'], out['right_ssim']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--cpu-frames', type=int, default=1)
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    h, w = 1024, 1280
    g = torch.Generator().manual_seed(0)
    left = torch.rand(args.frames, 3, h, w, generator=g)
    right = torch.rand(args.frames, 3, h, w, generator=g)
    pred = 0.3 * torch.sigmoid(torch.randn(args.frames, 4, h, w, generator=g))
    dev = torch.device('cuda:0')
    gl, gr, gp = left.to(dev), right.to(dev), pred.to(dev)
    ssim_loss = L.WeightedSSIMLoss(alpha=1)
    for _ in range(2):
        a = flow(gl, gr, gp, dev, ssim_loss)
    torch.cuda.synchronize()
    acc = {}
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.reps):
        stages = []
        a = flow(gl, gr, gp, dev, ssim_loss, stages)
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(stages[:-1], stages[1:]):
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
    t1.record()
    t1.synchronize()
    ms = t0.elapsed_time(t1) / args.reps
    out = {'frames': args.frames, 'shape': [h, w], 'ms': round(ms, 3),
           'frames_per_s': round(args.frames / (ms * 1e-3), 1),
           'stage_ms': {k: round(v / args.reps, 3) for k, v in acc.items()},
           'ause': float(a[0]), 'aurg': float(a[1])}
    # ---- fused front half (+ the SSIM metric) --------------------------------
    for _ in range(2):
        fa = fused_flow(gl, gr, gp, dev)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(args.reps):
        fa = fused_flow(gl, gr, gp, dev)
    t1.record()
    t1.synchronize()
    fms = t0.elapsed_time(t1) / args.reps
    # the stages of the fused flow (the same calls as evaluate_batch)
    from uncertainty_model_b200.train import evaluate as E
    facc = {}
    for _ in range(args.reps):
        ev = []

        def fmark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))
        fmark('start')
        images = torch.cat([gl, gr], dim=1)
        recon, error = E.reconstruct_and_error(images, gp)
        fmark('reconstruct_and_error')
        E.ssim(recon[:, 0:3], gl, reduction='sum', data_range=1.0)
        E.ssim(recon[:, 3:6], gr, reduction='sum', data_range=1.0)
        fmark('ssim_x2')
        S.curve(error, error, device=dev)
        S.curve(error, gp[:, 2:4], device=dev)
        S.curve(error, torch.rand_like(error), device=dev)
        fmark('three_curves')
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            facc[n1] = facc.get(n1, 0.0) + e0.elapsed_time(e1)
    out['fused'] = {'ms': round(fms, 3),
                    'stage_ms': {k: round(v / args.reps, 3) for k, v in facc.items()},
                    'frames_per_s': round(args.frames / (fms * 1e-3), 1),
                    'includes': 'SSIM metric of both views',
                    'ause': float(fa[0]), 'left_ssim': float(fa[2]),
                    'right_ssim': float(fa[3])}
    if args.cpu_frames > 0:
        from oracle import loss_port as P
        from oracle import spars_port as SP
        n = args.cpu_frames
        torch.set_num_threads(os.cpu_count() or 1)
        t = time.perf_counter()
        images = torch.cat([left[:n], right[:n]], 1)
        d, u = torch.split(pred[:n], [2, 2], dim=1)
        dl, dr = torch.split(d, [1, 1], dim=1)
        recon = torch.cat((P.warp_to_left(dl, right[:n]),
                           P.warp_to_right(dr, left[:n])), 1)
        err = P.image_error(images, recon, alpha=1.0)
        oc = SP.curve_reference_style(err, err)
        pc = SP.curve_reference_style(err, u)
        rc = SP.curve_reference_style(err, torch.rand_like(err))
        SP.ause_reference_style(oc, pc)
        SP.ause_reference_style(pc, rc)
        out['cpu_frames_per_s'] = round(n / (time.perf_counter() - t), 3)
        out['cpu_cores'] = os.cpu_count()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
