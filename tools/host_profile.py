"""cProfile of the host side of an eager training step (Python + ctypes +
launch calls).  python tools/host_profile.py [--steps 300]"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS, loss_config, synth  # noqa: E402
from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=300)
    args = ap.parse_args()
    b, h, w, lt, scale = WORKLOADS['c2']
    dev = torch.device('cuda:0')
    fn = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
    st, pr = synth(b, h, w, scale, 0)
    st = st.to(dev)
    pr = [p.to(dev).requires_grad_(True) for p in pr]

    def step():
        for p in pr:
            p.grad = None
        pyr = U.scale_pyramid(st, 4)
        rec = U.reconstruct_pyramid(pr, pyr)
        dl, el = fn(pyr, pr, rec, 0, None)
        (dl + el).backward()

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'host time per step (launch only): {1e6 * (t1 - t0) / args.steps:.0f} us; '
          f'with final sync: {1e6 * (t2 - t0) / args.steps:.0f} us')
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(args.steps):
        step()
    prof.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(prof, stream=s).sort_stats('tottime').print_stats(30)
    print(s.getvalue())


if __name__ == '__main__':
    main()
