"""Time line of one training step inside a CUDA-graph replay, from the
library's marker kernels (include/usl.h: usl_debug_timeline).

    python tools/step_timeline.py [--eager]

Prints, per marker, the median time since the step's first marker over the
replays.  The markers are one-thread kernels: they cost a few microseconds of
stream latency each, so the step is a little longer than in bench.py."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from uncertainty_model_b200 import _lib  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402
from uncertainty_model_b200.train.loss import TukraUncertaintyLoss  # noqa: E402

NAMES = {0: 'step start (before pyramid)', 1: 'pyramid done',
         2: 'fused scale 0 done', 3: 'fused scale 1 done',
         4: 'fused scale 2 done', 5: 'fused scale 3 done',
         6: 'transposed warp 0 done', 7: 'transposed warp 1 done',
         8: 'transposed warp 2 done', 9: 'transposed warp 3 done',
         10: 'reduce done', 11: 'combine done', 12: 'rescale done'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--eager', action='store_true')
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--height', type=int, default=256)
    ap.add_argument('--width', type=int, default=512)
    ap.add_argument('--loss', default='bayesian')
    ap.add_argument('--reps', type=int, default=40)
    ap.add_argument('--steps-per-graph', type=int, default=1)
    ap.add_argument('--mode', default='backward', choices=['backward', 'grad', 'forward'])
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    fn = TukraUncertaintyLoss(**bench.loss_config(a.loss)).to(dev)
    B, H, W = a.batch, a.height, a.width
    sets = []
    for _ in range(4):      # rotate inputs: > L2 in total
        st = torch.rand(B, 6, H, W, device=dev)
        preds = [(0.3 * torch.sigmoid(torch.randn(B, 4, H >> i, W >> i, device=dev)))
                 .requires_grad_() for i in range(4)]
        sets.append((st, preds))
    slots = torch.zeros(16, dtype=torch.int64, device=dev)
    keep = []

    def step(k):
        st, preds = sets[k]
        for p in preds:
            p.grad = None
        pyr = U.scale_pyramid(st, 4)
        rec = U.reconstruct_pyramid(preds, pyr)
        dl, el = fn(pyr, preds, rec, 1, None)
        if a.mode == 'backward':
            (dl + el).backward()
        elif a.mode == 'grad':
            keep[:] = torch.autograd.grad(dl + el, preds)

    for k in range(4):
        step(k)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().usl_debug_timeline(slots.data_ptr()))
    graphs = []
    if not a.eager:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            # (fresh leaves on the capture stream: see bench.py, Harness.capture)
            sets[:] = [(st, [p.detach().requires_grad_(True) for p in preds])
                       for st, preds in sets]
            for k in range(4):
                step(k)
            for k in range(4):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for j in range(a.steps_per_graph):
                        step((k + j) % 4)
                graphs.append(g)
        torch.cuda.synchronize()
    rows = []
    for r in range(a.reps):
        slots.zero_()
        torch.cuda.synchronize()
        if graphs:
            graphs[r % 4].replay()
        else:
            step(r % 4)
        torch.cuda.synchronize()
        rows.append(slots.cpu().tolist())
    # back-to-back replays: the period of a step and the gap between the end
    # of one step and the start of the next (slots 13 / 14)
    if graphs:
        for r in range(8):
            graphs[r % 4].replay()
        torch.cuda.synchronize()
        v = slots.cpu().tolist()
        print(f'back to back: period {(v[0] - v[13]) * 1e-3:.1f} us, '
              f'last stamp of the previous step -> this start {(v[0] - v[14]) * 1e-3:.1f} us')
    _lib.check(_lib.lib().usl_debug_timeline(None))
    t = torch.tensor(rows[4:], dtype=torch.float64)
    t0 = t[:, 0:1]
    rel = (t - t0) * 1e-3
    med = rel.median(0).values
    for i in sorted(NAMES, key=lambda i: float(med[i])):
        if (t[:, i] > 0).all():
            print(f'{float(med[i]):8.1f} us  {NAMES[i]}')


if __name__ == '__main__':
    main()
