"""A few eager steps of the hot path on config 2 (or --workload), for ncu.

    python tools/profile_step.py [--workload c2] [--steps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS, loss_config, synth  # noqa: E402
from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--steps', type=int, default=3)
    args = ap.parse_args()
    b, h, w, lt, scale = WORKLOADS[args.workload]
    dev = torch.device('cuda:0')
    fn = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
    sets = []
    for s in range(2):
        st, pr = synth(b, h, w, scale, s)
        sets.append((st.to(dev), [p.to(dev).requires_grad_(True) for p in pr]))
    for i in range(args.steps):
        stereo, preds = sets[i % 2]
        for p in preds:
            p.grad = None
        pyr = U.scale_pyramid(stereo, 4)
        rec = U.reconstruct_pyramid(preds, pyr)
        dl, el = fn(pyr, preds, rec, 0, None)
        (dl + el).backward()
    torch.cuda.synchronize()
    print('ok', float(dl), float(el))


if __name__ == '__main__':
    main()
