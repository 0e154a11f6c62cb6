"""Gradient parity metric shared by the GPU and the CPU tests.

north_star: gradients within 1e-4 relative in fp32.  The comparison is against
the fp64 run of the oracle, element by element, with the elements where the
loss is not differentiable within fp32 rounding (oracle/kinks.py: integer
crossings of a sampling coordinate, sign changes of an absolute value) set
aside by an EXPLICIT mask computed from the oracle's fp64 intermediates -- not
by the size of the error.  Per channel of every scale:

  * every element outside the mask: relative L2 error <= GRAD_REL, and no
    single element off by more than ELEM_REL of the largest gradient value;
  * the masked elements are counted (fraction <= MASK_MAX: the margins are
    2.5e-7 * width wide on both sides of an integer -- 2.6e-4 of the pixels at
    w = 512, twice that at 1024 -- plus the sign kinks and the four taps each
    of those scatters to: ~1e-3 .. 2e-3 in all) and must still be finite and
    bounded by the size of a one-sided jump;
  * the untrimmed relative L2 error is reported (`stats`), never asserted: k
    legitimately one-sided elements out of n put ~sqrt(k/n) on it.
"""
import json
import os

import numpy as np
import torch

GRAD_REL = 1e-4
ELEM_REL = 5e-4
MASK_MAX = 5e-3


def oracle_reference(stereo, preds, cfg):
    """fp64 oracle step + kink masks for seeded inputs."""
    from oracle import kinks as K
    from oracle import loss_port as P
    ref = P.step_detailed(stereo.double(), [p.double() for p in preds], cfg)
    assert K.clamp_kinks(ref['pyramid'], ref['recons']) == 0
    # how far an fp32 pyramid (the reference's, ours) is from the fp64 one
    pyr32 = P.pyramid(stereo.float(), len(preds))
    image_eps = [float((a.double() - b).abs().max())
                 for a, b in zip(pyr32, ref['pyramid'])]
    ref['masks'] = K.kink_masks(ref['pyramid'], ref['preds'], cfg,
                                ref['errors'], image_eps)
    # `pooling`: a kink of a 3x3 mean touches the nine inputs under it
    pooled = bool((cfg.get('error_loss_config') or {}).get('pooling'))
    ref['mask_max'] = MASK_MAX * (9 if pooled else 1)
    return ref


def masks_for(stereo, preds, cfg):
    """Kink masks alone (the reference gradients come from a fixture)."""
    return oracle_reference(stereo, preds, cfg)['masks']


def channel_stats(mine, ref, mask):
    mine = np.asarray(mine, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    mask = np.asarray(mask, dtype=bool)
    assert mine.shape == ref.shape == mask.shape
    assert np.isfinite(mine).all()
    d = np.abs(mine - ref)
    keep = ~mask
    scale = max(np.abs(ref).max(), 1e-300)
    return dict(
        full=float(np.linalg.norm(d) / max(np.linalg.norm(ref), 1e-300)),
        outside=float(np.linalg.norm(d[keep]) /
                      max(np.linalg.norm(ref[keep]), 1e-300)),
        max_outside=float(d[keep].max() / scale) if keep.any() else 0.0,
        max_inside=float(d[mask].max() / scale) if mask.any() else 0.0,
        masked=float(mask.mean()))


def check_grads(grads, ref_grads, masks, what='', grad_rel=GRAD_REL,
                elem_rel=ELEM_REL, mask_max=MASK_MAX):
    """Asserts the parity bar above for every (scale, channel); returns the
    per-channel statistics."""
    out = []
    for i, (g, r, m) in enumerate(zip(grads, ref_grads, masks)):
        assert g is not None, (what, i)
        g = g.detach().cpu().numpy() if isinstance(g, torch.Tensor) else g
        r = r.detach().cpu().numpy() if isinstance(r, torch.Tensor) else r
        m = m.cpu().numpy() if isinstance(m, torch.Tensor) else m
        for ch in range(g.shape[1]):
            if not np.any(r[:, ch]):
                assert not np.any(g[:, ch]), (what, i, ch)
                continue
            st = channel_stats(g[:, ch], r[:, ch], m[:, ch])
            st.update(scale=i, channel=ch)
            out.append(st)
            assert st['outside'] <= grad_rel, (what, st)
            assert st['max_outside'] <= elem_rel, (what, st)
            assert st['masked'] <= mask_max, (what, st)
            assert st['max_inside'] <= 4.0, (what, st)
    return out


def record(name, stats):
    """Appends the per-channel statistics of a benchmark-shape comparison to
    gpurun_out/parity_stats.jsonl (copied to profiles/ by hand: the untrimmed
    norms and mask sizes are part of the evidence, not of the pass/fail)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'parity_stats.jsonl'), 'a') as f:
            f.write(json.dumps({'case': name, 'channels': stats}) + '\n')
    except OSError:
        pass
