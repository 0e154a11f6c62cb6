"""Batch-sharded data parallelism for the loss path.

The path shards over samples with no data exchange: every rank runs the fused
kernels on its own shard.  The only collective is one NCCL all-reduce of the
raw per-term sums (fp64, scales x 6 values) per step -- issued behind the
column kernels on a high-priority side stream, beside the transposed warps,
with no host synchronisation -- so that every rank reports the loss of the GLOBAL
batch (the reference leaves the loss un-reduced, parallel_main.py:156-160:
rank 0 logs its own shard's).

Gradients come in two normalisations (`shard_loss(..., gradients=)`):

  'ddp'     gradients of the LOCAL mean, 1/(B_local*h*w) -- what the
            reference's launcher needs: it wraps the model in
            DistributedDataParallel (parallel_main.py:158), which AVERAGES the
            parameter gradients over ranks, and the average of the local-mean
            gradients is the global-mean gradient.  Default.
  'global'  gradients of the GLOBAL mean, 1/(B_global*h*w): every rank's
            gradient w.r.t. its predictions is the matching slice of the
            full-batch gradient.  For a gradient SUM all-reduce, or when the
            prediction gradients are consumed directly -- NOT under torch DDP
            (the parameter gradients would come out 1/world too small).

Sparsification shards over frames the same way: all-reduce the per-step sums
of normalised tail means and the row count.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` samples owned by `rank`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank/world')
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_loss(loss_module, group: Optional['dist.ProcessGroup'] = None,
               gradients: str = 'ddp'):
    """Make a `TukraUncertaintyLoss` report global-batch losses; `gradients`
    selects their normalisation (module docstring): 'ddp' | 'global'.

    Requires equal shard sizes on all ranks (as `DistributedSampler` gives)."""
    if gradients not in ('ddp', 'global'):
        raise ValueError("gradients must be 'ddp' or 'global'")
    if not dist.is_initialized():
        raise RuntimeError('torch.distributed is not initialised')
    loss_module.reduce_group = group if group is not None else dist.group.WORLD
    loss_module.world_size = dist.get_world_size(group)
    loss_module.grad_world_size = \
        loss_module.world_size if gradients == 'global' else 1
    return loss_module


def reduce_term_sums(sums: Tensor, group=None) -> Tensor:
    """all-reduce(sum) of raw per-term sums (any backend; fp64)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, group=group)
    return sums


def combine_terms(sums: Tensor, coefs: Tensor) -> Tuple[Tensor, Tensor]:
    """Host-side mirror of usl_loss_combine for CPU tests of the sharding
    algebra: sums, coefs (scales, 6) -> (disp_loss, error_loss)."""
    prod = sums.double() * coefs.double()
    return prod[:, :3].sum(), prod[:, 3:].sum()


def sharded_curve(oracle_error: Tensor, predicted_error: Tensor,
                  kernel_size: int = 11, steps: int = 100, group=None,
                  device='cpu') -> Tensor:
    """`sparsification.curve` over frames sharded across ranks."""
    from .train import sparsification as S
    acc, rows, _ = S.curve_sums(oracle_error, predicted_error, kernel_size,
                                steps)
    count = torch.tensor([float(rows)], dtype=torch.float64, device=acc.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, group=group)
        dist.all_reduce(count, group=group)
    return S.finish_curve(acc, int(count.item())).to(device)
