"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the REAL
reference (imported from /root/reference through oracle/ref_shim.py).

Run in the build container only (the reference tree does not travel):

    python oracle/make_golden.py

The reference has no tests or golden vectors of its own (SURVEY.md section
8c), so these fixtures -- outputs of the unmodified reference code on seeded
synthetic inputs, inputs included -- are the pin for the oracle port
(`oracle/loss_port.py`, `oracle/spars_port.py`) and, through it, for the CUDA
path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_shim import import_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

CONFIG_YML_LOSS = dict(   # config.yml:103-117
    wssim_weight=1.0, consistency_weight=1.0, smoothness_weight=1.0,
    adversarial_weight=0.85, perceptual_weight=0.05,
    predictive_error_weight=1.0, wssim_alpha=0.85, perceptual_start=5,
    adversarial_loss_type='mse',
    error_loss_config=dict(loss_type='l1', smoothness_weight=0,
                           consistency_weight=0.5, pooling=False))


def loss_config(loss_type='l1', **err_overrides):
    cfg = {k: (dict(v) if isinstance(v, dict) else v)
           for k, v in CONFIG_YML_LOSS.items()}
    cfg['error_loss_config']['loss_type'] = loss_type
    cfg['error_loss_config'].update(err_overrides)
    return cfg


def make_inputs(b, h, w, scale, seed, smooth=False):
    """SURVEY.md section 8d / Appendix C draw order."""
    g = torch.Generator().manual_seed(seed)
    left = torch.rand(b, 3, h, w, generator=g)
    right = torch.rand(b, 3, h, w, generator=g)
    preds = [scale * torch.sigmoid(torch.randn(b, 4, h // 2 ** i, w // 2 ** i,
                                               generator=g))
             for i in range(4)]
    if smooth:
        k = torch.ones(3, 1, 5, 5) / 25
        left = torch.nn.functional.conv2d(
            torch.nn.functional.pad(left, (2, 2, 2, 2), mode='reflect'),
            k, groups=3)
        right = torch.roll(left, 3, dims=3)
        ks = torch.ones(4, 1, 5, 5) / 25
        preds = [torch.nn.functional.conv2d(
            torch.nn.functional.pad(p, (2, 2, 2, 2), mode='replicate'),
            ks, groups=4) for p in preds]
    return left, right, preds


def run_reference_loss(L, u, left, right, preds, cfg, dtype):
    images = torch.cat([left, right], dim=1).to(dtype)
    preds = [p.to(dtype).clone().requires_grad_(True) for p in preds]
    pyr = u.scale_pyramid(images, 4)
    rec = u.reconstruct_pyramid(preds, pyr)
    fn = L.TukraUncertaintyLoss(**cfg)
    errors = []
    # collect the per-scale error maps the way loss.py:544-548 exposes them
    orig = fn.wssim.forward

    def spy(images_, recon_):
        out = orig(images_, recon_)
        errors.append(fn.wssim.previous_image_error.detach().clone())
        return out
    fn.wssim.forward = spy
    dl, el = fn(pyr, preds, rec, 0, None)
    (dl + el).backward()
    return dict(pyr=pyr, rec=[r.detach() for r in rec], errors=errors,
                disp_loss=dl.detach(), error_loss=el.detach(),
                grads=[p.grad for p in preds])


def loss_case(L, u, name, b, h, w, scale, seed, cfg, smooth=False):
    left, right, preds = make_inputs(b, h, w, scale, seed, smooth)
    r32 = run_reference_loss(L, u, left, right, preds, cfg, torch.float32)
    r64 = run_reference_loss(L, u, left, right, preds, cfg, torch.float64)
    out = dict(left=left.numpy(), right=right.numpy(),
               disp_loss_f32=r32['disp_loss'].numpy(),
               error_loss_f32=r32['error_loss'].numpy(),
               disp_loss_f64=r64['disp_loss'].numpy(),
               error_loss_f64=r64['error_loss'].numpy())
    for i in range(4):
        out[f'pred{i}'] = preds[i].numpy()
        out[f'pyr{i}_f32'] = r32['pyr'][i].numpy()
        out[f'rec{i}_f32'] = r32['rec'][i].numpy()
        out[f'err{i}_f32'] = r32['errors'][i].numpy()
        out[f'grad{i}_f32'] = r32['grads'][i].numpy()
        out[f'grad{i}_f64'] = r64['grads'][i].numpy()
    np.savez_compressed(os.path.join(GOLDEN, f'loss_{name}.npz'), **out)
    print(f'loss_{name}: disp={float(r32["disp_loss"]):.7f} '
          f'err={float(r32["error_loss"]):.7f}')


def component_case(L, u):
    """Each public loss class on its own (evaluate.py:124,151 and the
    standalone modules of loss.py), fp32 + fp64, small shape."""
    g = torch.Generator().manual_seed(11)
    b, h, w = 2, 20, 36
    images = torch.rand(b, 6, h, w, generator=g)
    recon = torch.rand(b, 6, h, w, generator=g)
    pred = 0.5 * torch.sigmoid(torch.randn(b, 4, h, w, generator=g))
    error = torch.rand(b, 2, h, w, generator=g)
    out = dict(images=images.numpy(), recon=recon.numpy(), pred=pred.numpy(),
               error=error.numpy())
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        im = images.to(dt)
        rc = recon.to(dt).requires_grad_(True)
        pr = pred.to(dt).requires_grad_(True)
        er = error.to(dt)
        for alpha in (0.85, 1.0):
            ws = L.WeightedSSIMLoss(alpha)
            out[f'image_error_a{alpha}_{tag}'] = \
                ws.image_error(im, rc).detach().numpy()
            val = ws(im, rc)
            grad, = torch.autograd.grad(val, rc)
            out[f'wssim_a{alpha}_{tag}'] = val.detach().numpy()
            out[f'wssim_a{alpha}_grad_{tag}'] = grad.numpy()
        val = L.ConsistencyLoss()(pr[:, 0:2])
        grad, = torch.autograd.grad(val, pr)
        out[f'cons_{tag}'] = val.detach().numpy()
        out[f'cons_grad_{tag}'] = grad.numpy()
        val = L.ConsistencyLoss()(pr[:, 2:4], pr[:, 0:2])
        grad, = torch.autograd.grad(val, pr)
        out[f'cons_ab_{tag}'] = val.detach().numpy()
        out[f'cons_ab_grad_{tag}'] = grad.numpy()
        val = L.SmoothnessLoss()(pr[:, 0:2], im)
        grad, = torch.autograd.grad(val, pr)
        out[f'smooth_{tag}'] = val.detach().numpy()
        out[f'smooth_grad_{tag}'] = grad.numpy()
        for lt in ('l1', 'bayesian', 'log_bayesian'):
            for pooling in (False, True):
                fn = L.ReprojectionErrorLoss(lt, 0.7, 0.3, pooling)
                val = fn(pr, im, er)
                grad, = torch.autograd.grad(val, pr)
                key = f'reproj_{lt}_{"pool" if pooling else "nopool"}'
                out[f'{key}_{tag}'] = val.detach().numpy()
                out[f'{key}_grad_{tag}'] = grad.numpy()
        dl = pr[:, 0:1].detach()
        out[f'recon_left_{tag}'] = \
            u.reconstruct_left_image(dl, im[:, 3:6]).numpy()
        out[f'recon_right_{tag}'] = \
            u.reconstruct_right_image(pr[:, 1:2].detach(), im[:, 0:3]).numpy()
    np.savez_compressed(os.path.join(GOLDEN, 'components.npz'), **out)
    print('components written')


def anchors(L, u):
    """Full-size seeded anchor scalars: inputs are
    regenerated from the seed at test time, only the scalars are stored."""
    out = {}
    for name, (b, h, w, lt) in dict(
            c1=(2, 256, 512, 'l1'), c2=(16, 256, 512, 'bayesian'),
            c3_shard=(8, 192, 384, 'l1'),
            c4=(2, 512, 1024, 'l1')).items():
        left, right, preds = make_inputs(b, h, w, 0.3, 0)
        r = run_reference_loss(L, u, left, right, preds, loss_config(lt),
                               torch.float32)
        out[f'{name}_shape'] = np.array([b, h, w])
        out[f'{name}_disp_loss'] = r['disp_loss'].numpy()
        out[f'{name}_error_loss'] = r['error_loss'].numpy()
        out[f'{name}_grad_l2'] = np.array(
            [float(gr.double().norm()) for gr in r['grads']])
        out[f'{name}_grad_sum'] = np.array(
            [float(gr.double().sum()) for gr in r['grads']])
        print(name, float(r['disp_loss']), float(r['error_loss']))
    np.savez_compressed(os.path.join(GOLDEN, 'anchors.npz'), **out)


def spars_cases(S):
    from oracle.spars_port import synthetic_maps
    out = {}
    for name, (f, h, w, ties) in dict(
            small=(2, 40, 56, False), ties=(2, 40, 56, True),
            ragged=(1, 23, 31, False)).items():
        err, unc = synthetic_maps(f, h, w, seed=3, ties=ties)
        oc = S.curve(err, err)
        pc = S.curve(err, unc)
        pooled = torch.nn.AvgPool2d(11, stride=1)
        out[f'{name}_err'] = err.numpy()
        out[f'{name}_unc'] = unc.numpy()
        out[f'{name}_oracle_curve'] = oc.numpy()
        out[f'{name}_pred_curve'] = pc.numpy()
        out[f'{name}_ause'] = S.ause(oc, pc).numpy()
        out[f'{name}_pooled_err'] = pooled(err).numpy()
        out[f'{name}_pooled_unc'] = pooled(unc).numpy()
        out[f'{name}_stable_order_unc'] = pooled(unc).view(f, 2, -1) \
            .argsort(dim=2, descending=True, stable=True).numpy()
    # Appendix C anchor, stored as scalars only
    err, unc = synthetic_maps(4, 1024, 1280, seed=0)
    oc = S.curve(err, err)
    pc = S.curve(err, unc)
    out['c5_4frames_oracle_curve'] = oc.numpy()
    out['c5_4frames_pred_curve'] = pc.numpy()
    out['c5_4frames_ause'] = S.ause(oc, pc).numpy()
    np.savez_compressed(os.path.join(GOLDEN, 'spars.npz'), **out)
    print('spars written, c5 ause', float(out['c5_4frames_ause']))


def post_case(u):
    """train/utils.py:199-245 (combine_disparity: numpy only, runs as is)."""
    g = torch.Generator().manual_seed(21)
    left = 0.3 * torch.rand(1, 48, 80, generator=g)
    right = 0.3 * torch.rand(1, 48, 80, generator=g)
    out = u.combine_disparity(left, right)
    np.savez_compressed(os.path.join(GOLDEN, 'post.npz'), left=left.numpy(),
                        right=right.numpy(), combined=out.numpy())
    print('post written')


def main():
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)
    L, u, S = import_reference()
    loss_case(L, u, 'l1_default', 2, 32, 64, 0.3, 1, loss_config('l1'))
    loss_case(L, u, 'bayesian_default', 2, 32, 64, 0.3, 2,
              loss_config('bayesian'))
    loss_case(L, u, 'log_bayesian_scale1', 1, 48, 40, 1.0, 3,
              loss_config('log_bayesian'))
    loss_case(L, u, 'l1_allterms', 1, 40, 72, 0.5, 4,
              loss_config('l1', smoothness_weight=0.6,
                          consistency_weight=0.8))
    loss_case(L, u, 'bayesian_pooling', 1, 40, 72, 0.5, 5,
              loss_config('bayesian', smoothness_weight=0.6,
                          consistency_weight=0.8, pooling=True))
    loss_case(L, u, 'bayesian_smooth', 2, 32, 64, 0.3, 6,
              loss_config('bayesian'), smooth=True)
    component_case(L, u)
    anchors(L, u)
    spars_cases(S)
    post_case(u)


if __name__ == '__main__':
    main()
