"""Duration of the consistency-scatter stage alone (config-2 shape) for the
benchmark's white-noise disparities and for smooth, realistic ones.

    python tools/scatter_bench.py [--workload c2] [--reps 50]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from bench import WORKLOADS, loss_config, synth  # noqa: E402
from uncertainty_model_b200 import functional as K  # noqa: E402
from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def smooth(preds, k):
    ker = torch.ones(4, 1, k, k) / (k * k)
    return [F.conv2d(F.pad(p, (k // 2,) * 4, mode='replicate'), ker, groups=4)
            for p in preds]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--reps', type=int, default=50)
    args = ap.parse_args()
    b, h, w, lt, scale = WORKLOADS[args.workload]
    dev = torch.device('cuda:0')
    fn = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
    st = fn._settings()
    one = torch.ones((), device=dev)
    out = {}
    for name, k in (('noise', 0), ('smooth9', 9), ('smooth31', 31)):
        sets = []
        for s in range(4):
            stereo, preds = synth(b, h, w, scale, s)
            if k:
                preds = smooth(preds, k)
            stereo = stereo.to(dev)
            pyr = U.scale_pyramid(stereo, 4)
            cfgs, bsc, keep = [], [], []
            for i in range(4):
                pd = preds[i].to(dev)
                bb, _, hh, ww = pd.shape
                g = torch.empty_like(pd)
                keep.append((pd, g))
                cfgs.append(K.make_config(st.terms(), st,
                                          st.coefs(i, bb * hh * ww)))
                bsc.append(K.make_scale(pyr[i], pd[:, 0:2], pd[:, 2:4],
                                        shape=(bb, hh, ww), grad_disp=g[:, 0:2],
                                        grad_unc=g[:, 2:4]))
            sets.append((cfgs, bsc, keep, pyr))
        for i in range(4):
            K.loss_backward(sets[i][0], sets[i][1], one, one, dev, 1)
        torch.cuda.synchronize()
        total = 0.0
        for r in range(args.reps):
            c, s_, _, _ = sets[r % 4]
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            K.loss_backward(c, s_, one, one, dev, 1)
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        out[name] = round(total / args.reps * 1e3, 1)
    print(json.dumps({'scatter_us': out}))


if __name__ == '__main__':
    main()
