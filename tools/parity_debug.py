"""Where the CUDA gradients differ from the fp64 oracle outside the kink mask:
lists the worst unmasked elements of a benchmark shape with the distances of
their pixel to every kink (tools for the parity tests, GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import parity  # noqa: E402
from oracle import kinks as K  # noqa: E402
from oracle.make_golden import loss_config, make_inputs  # noqa: E402
from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def main():
    b, h, w = [int(v) for v in sys.argv[1:4]]
    lt = sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 12
    dev = torch.device('cuda:0')
    cfg = loss_config(lt)
    left, right, preds = make_inputs(b, h, w, 0.3, 0)
    stereo = torch.cat([left, right], 1)
    gp = [p.to(dev).requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(stereo.to(dev), 4)
    dl, el = L.TukraUncertaintyLoss(**cfg)(pyr, gp, U.reconstruct_pyramid(gp, pyr), 0, None)
    (dl + el).backward()
    ref = parity.oracle_reference(stereo, preds, cfg)
    from oracle import loss_port as P
    _, _, g32 = P.step(stereo, preds, cfg)
    for i in range(4):
        g = gp[i].grad.cpu().double().numpy()
        r = ref['grads'][i].numpy()
        m = ref['masks'][i].numpy()
        p = ref['preds'][i]
        im = ref['pyramid'][i]
        hh, ww = p.shape[-2:]
        for ch in range(4):
            d = np.abs(g[:, ch] - r[:, ch])
            d[m[:, ch]] = 0
            scale = np.abs(r[:, ch]).max()
            order = np.argsort(d.ravel())[::-1][:top]
            bad = [(np.unravel_index(o, d.shape), d.ravel()[o] / scale) for o in order
                   if d.ravel()[o] / scale > 2e-4]
            if not bad:
                continue
            print(f'--- scale {i} ch {ch} ({hh}x{ww}) max|ref| {scale:.3e}: '
                  f'{(d / scale > 2e-4).sum()} unmasked elements above 2e-4')
            v = ch % 2
            sign = -1.0 if v == 0 else 1.0
            a = p[:, ch:ch + 1]
            opp_img = im[:, 3:6] if v == 0 else im[:, 0:3]
            own_img = im[:, 0:3] if v == 0 else im[:, 3:6]
            wi = K.explicit_warp(sign * a, opp_img)
            wd = K.explicit_warp(sign * a, p[:, 1 - v:2 - v])
            for (bb, y, x), e in bad:
                ix = float(wi['ix'][bb, y, x])
                line = (f'  b{bb} y{y} x{x}: err {e:.2e} ours {g[bb, ch, y, x]:+.4e} '
                        f'ref {r[bb, ch, y, x]:+.4e} ref32 {float(g32[i][bb, ch, y, x]):+.4e} frac(ix) {ix - np.floor(ix):.6f}')
                if ch < 2:
                    line += f' I-rec {[float(t) for t in (own_img - wi["out"])[bb, :, y, x]]} slope {[float(t) for t in wi["slope"][bb, :, y, x]]} slope_y {[float(t) for t in wi["slope_y"][bb, :, y, x]]}'
                if ch < 2:
                    l1 = (own_img - wi['out'])[bb, :, y, x].abs().min()
                    f = float((a - wd['out'])[bb, 0, y, x])
                    line += f' min|I-rec| {float(l1):.2e} d-warp(d) {f:+.2e}'
                else:
                    f = float((a - wd['out'])[bb, 0, y, x])
                    line += f' u-warp(d;u) {f:+.2e}'
                # neighbours masked? (scatter taps)
                nb = m[bb, ch, max(0, y - 1):y + 2, max(0, x - 2):x + 3].sum()
                line += f' masked-neighbours {int(nb)}'
                print(line)


if __name__ == '__main__':
    main()
