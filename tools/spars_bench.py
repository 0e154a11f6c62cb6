"""Throughput of the sparsification / AUSE path (BASELINE.json config 5 shape)
on one GPU, with the CPU oracle port beside it.

    python tools/spars_bench.py [--frames 16] [--cpu-frames 1]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from uncertainty_model_b200.train import sparsification as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--cpu-frames', type=int, default=1)
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    h, w, k = 1024, 1280, 11
    g = torch.Generator().manual_seed(0)
    err = torch.rand(args.frames, 2, h, w, generator=g)
    unc = (err + 0.2 * torch.rand(args.frames, 2, h, w, generator=g)).clamp(0, 1)
    dev = torch.device('cuda:0')
    e, u = err.to(dev), unc.to(dev)

    def run():
        oc = S.curve(e, e, device=dev)
        pc = S.curve(e, u, device=dev)
        return S.ause(oc, pc)

    for _ in range(2):
        a = run()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.reps):
        a = run()
    t1.record()
    t1.synchronize()
    ms = t0.elapsed_time(t1) / args.reps
    n = (h - k + 1) * (w - k + 1)
    elems = args.frames * 2 * n * 2            # two curves
    out = {'frames': args.frames, 'ms': ms,
           'frames_per_s': args.frames / (ms * 1e-3),
           'alg_GBps': 16.0 * elems / (ms * 1e-3) / 1e9, 'ause': float(a)}
    if args.cpu_frames > 0:
        from oracle import spars_port as P
        ce, cu = err[:args.cpu_frames], unc[:args.cpu_frames]
        torch.set_num_threads(os.cpu_count() or 1)
        t = time.perf_counter()
        oc = P.curve_reference_style(ce, ce)
        pc = P.curve_reference_style(ce, cu)
        P.ause_reference_style(oc, pc)
        dt = time.perf_counter() - t
        out['cpu_frames_per_s'] = args.cpu_frames / dt
        out['cpu_cores'] = os.cpu_count()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
