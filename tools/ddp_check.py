"""Multi-GPU parity check of the batch-sharded loss path (run under torchrun).

Every rank evaluates its shard with `shard_loss`; against the single-GPU
full-batch run on the same data:

  * the reported losses are the global ones (both gradient modes);
  * gradients='global': the local gradients are the matching slices of the
    full-batch gradients, bit for bit (same kernels, same per-sample work);
  * gradients='ddp': they are those slices times the world size (exact for a
    power-of-two world), and a toy network wrapped in torch
    DistributedDataParallel -- what the reference's launcher does
    (parallel_main.py:158) -- ends up with the parameter gradients of the
    single-GPU full-batch step;
  * `distributed.sharded_curve`: frames sharded over ranks give the
    single-GPU sparsification curve bit for bit.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        tools/ddp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

from bench import loss_config, synth  # noqa: E402
from uncertainty_model_b200.distributed import (shard_bounds, shard_loss,  # noqa: E402
                                                sharded_curve)
from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import sparsification as S  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def run(fn, stereo, preds):
    preds = [p.clone().requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(stereo, 4)
    rec = U.reconstruct_pyramid(preds, pyr)
    dl, el = fn(pyr, preds, rec, 0, None)
    (dl + el).backward()
    return dl.detach(), el.detach(), [p.grad for p in preds]


class ToyNet(torch.nn.Module):
    """Stands in for the disparity network: left image -> 4-scale predictions
    (B,4,h,w) in (0, 0.3), layers/decoder.py:239-246."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.heads = torch.nn.ModuleList(
            [torch.nn.Conv2d(3, 4, 3, padding=1) for _ in range(4)])

    def forward(self, left):
        out = []
        for i, head in enumerate(self.heads):
            x = left if i == 0 else torch.nn.functional.avg_pool2d(left, 2 ** i)
            out.append(0.3 * torch.sigmoid(head(x)))
        return out


def net_step(net, fn, stereo):
    net.zero_grad(set_to_none=True)
    preds = net(stereo[:, 0:3])
    pyr = U.scale_pyramid(stereo, 4)
    dl, el = fn(pyr, preds, U.reconstruct_pyramid(preds, pyr), 0, None)
    (dl + el).backward()
    return dl.detach(), el.detach()


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    b = 4 * world
    stereo, preds = synth(b, 64, 128, 0.3, 7)
    stereo = stereo.to(dev)
    preds = [p.to(dev) for p in preds]
    lo, hi = shard_bounds(b, rank, world)
    ok = True
    for lt in ('l1', 'bayesian'):
        full = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
        fdl, fel, fg = run(full, stereo, preds)
        for mode, factor in (('global', 1.0), ('ddp', float(world))):
            sharded = shard_loss(
                L.TukraUncertaintyLoss(**loss_config(lt)).to(dev),
                gradients=mode)
            sdl, sel, sg = run(sharded, stereo[lo:hi].contiguous(),
                               [p[lo:hi].contiguous() for p in preds])
            torch.cuda.synchronize()
            e1 = abs(float(sdl) - float(fdl)) / abs(float(fdl))
            e2 = abs(float(sel) - float(fel)) / abs(float(fel))
            if world & (world - 1) == 0:
                same = all(torch.equal(a, c[lo:hi] * factor)
                           for a, c in zip(sg, fg))
            else:
                same = all(torch.allclose(a, c[lo:hi] * factor, rtol=1e-6,
                                          atol=0) for a, c in zip(sg, fg))
            good = e1 < 1e-6 and e2 < 1e-6 and same
            ok = ok and good
            print(f'rank {rank} {lt} gradients={mode}: loss rel err {e1:.2e} '
                  f'{e2:.2e} grads match {same}', flush=True)

    # ---- torch DDP around a toy network: parameter gradients ---------------
    ref_net = ToyNet().to(dev)
    net_step(ref_net, L.TukraUncertaintyLoss(**loss_config('bayesian')).to(dev),
             stereo)
    ref_grads = [p.grad.clone() for p in ref_net.parameters()]
    ddp_net = DDP(ToyNet().to(dev), device_ids=[local])
    fn = shard_loss(L.TukraUncertaintyLoss(**loss_config('bayesian')).to(dev),
                    gradients='ddp')
    net_step(ddp_net, fn, stereo[lo:hi].contiguous())
    torch.cuda.synchronize()
    worst = 0.0
    for p, r in zip(ddp_net.module.parameters(), ref_grads):
        worst = max(worst, float((p.grad - r).norm() / r.norm()))
    good = worst < 2e-5      # cuDNN picks other algorithms for other batches
    ok = ok and good
    print(f'rank {rank} DDP toy network: parameter gradients rel err '
          f'{worst:.2e}', flush=True)

    # ---- sparsification: frames sharded over ranks --------------------------
    g = torch.Generator().manual_seed(11)
    frames = 2 * world
    err = torch.rand(frames, 2, 96, 160, generator=g)
    unc = (err + 0.2 * torch.rand(frames, 2, 96, 160, generator=g)).clamp(0, 1)
    err, unc = err.to(dev), unc.to(dev)
    flo, fhi = shard_bounds(frames, rank, world)
    for name, a, c in (('oracle', err, err), ('predicted', err, unc)):
        single = S.curve(a, c, device=dev)
        shard = sharded_curve(a[flo:fhi].contiguous(), c[flo:fhi].contiguous(),
                              device=dev)
        same = torch.equal(single, shard)
        ok = ok and same
        print(f'rank {rank} sharded {name} curve identical {same}', flush=True)

    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print('DDP CHECK', 'OK' if flag.item() == 1.0 else 'FAILED')
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
