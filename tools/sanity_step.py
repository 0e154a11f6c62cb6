"""Small forward-only and training steps of the hot path, for
compute-sanitizer / repeated-run stress.  python tools/sanity_step.py [--reps N]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import loss_config, synth  # noqa: E402
from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--shape', default='2,64,128')
    args = ap.parse_args()
    b, h, w = (int(v) for v in args.shape.split(','))
    dev = torch.device('cuda:0')
    for lt in ('bayesian', 'l1'):
        fn = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
        st, pr = synth(b, h, w, 0.3, 1)
        st = st.to(dev)
        pr = [p.to(dev).requires_grad_(True) for p in pr]
        for i in range(args.reps):
            with torch.no_grad():
                pyr = U.scale_pyramid(st, 4)
                rec = U.reconstruct_pyramid(pr, pyr)
                a, c = fn(pyr, pr, rec, 0, None)
            for p in pr:
                p.grad = None
            pyr = U.scale_pyramid(st, 4)
            rec = U.reconstruct_pyramid(pr, pyr)
            dl, el = fn(pyr, pr, rec, 0, None)
            (dl + el).backward()
            torch.cuda.synchronize()
        print(lt, float(a), float(c), float(dl.detach()), float(el.detach()),
              float(pr[0].grad.abs().sum()))


if __name__ == '__main__':
    main()
