"""Multi-GPU parity check of the batch-sharded loss path (run under torchrun):
every rank evaluates its shard with `shard_loss`; the global losses must equal
the single-GPU full-batch losses, the local gradients the matching slices of
the full-batch gradients (bit for bit: same kernels, same per-sample work).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        tools/ddp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import loss_config, synth  # noqa: E402
from uncertainty_model_b200.distributed import shard_bounds, shard_loss  # noqa: E402
from uncertainty_model_b200.train import loss as L  # noqa: E402
from uncertainty_model_b200.train import utils as U  # noqa: E402


def run(fn, stereo, preds):
    preds = [p.clone().requires_grad_(True) for p in preds]
    pyr = U.scale_pyramid(stereo, 4)
    rec = U.reconstruct_pyramid(preds, pyr)
    dl, el = fn(pyr, preds, rec, 0, None)
    (dl + el).backward()
    return dl.detach(), el.detach(), [p.grad for p in preds]


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    b = 4 * world
    stereo, preds = synth(b, 64, 128, 0.3, 7)
    stereo = stereo.to(dev)
    preds = [p.to(dev) for p in preds]
    ok = True
    for lt in ('l1', 'bayesian'):
        full = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
        fdl, fel, fg = run(full, stereo, preds)
        lo, hi = shard_bounds(b, rank, world)
        sharded = shard_loss(L.TukraUncertaintyLoss(**loss_config(lt)).to(dev))
        sdl, sel, sg = run(sharded, stereo[lo:hi].contiguous(),
                           [p[lo:hi].contiguous() for p in preds])
        torch.cuda.synchronize()
        e1 = abs(float(sdl) - float(fdl)) / abs(float(fdl))
        e2 = abs(float(sel) - float(fel)) / abs(float(fel))
        same = all(torch.equal(a, c[lo:hi]) for a, c in zip(sg, fg))
        good = e1 < 1e-6 and e2 < 1e-6 and same
        ok = ok and good
        print(f'rank {rank} {lt}: loss rel err {e1:.2e} {e2:.2e} '
              f'grads identical {same}', flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print('DDP CHECK', 'OK' if flag.item() == 1.0 else 'FAILED')
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
