"""Multi-GPU plumbing for the OT distillation path: one process per GPU, images sharded across ranks.

Every image's OT problems are independent (reference ``losses/loss_libs.py:22-50`` never mixes images), so the
path shards by image with NO data-path collective: each rank runs the fused kernel on its block of images.
Two optional exchanges exist around it:

* :func:`global_mean_loss` -- all-reduce of ``(sum_i F_i, n_valid)`` (two scalars) when the caller wants the mean
  over the non-empty images of the GLOBAL batch instead of the reference's per-process mean
  (``losses/kd_loss.py:99-100`` divides by the local count; its own multi-GPU runs never exchange it);
* :func:`allreduce_student_grads` -- the data-parallel SUM/mean of the student's parameter gradients in one flat
  bucket (the reference wraps the model in DDP and immediately unwraps it, ``libs/train_libs.py:124-130``, so it
  performs no gradient exchange at all; this is the collective a correct data-parallel run needs).

Backend: NCCL over NVLink/NVSwitch on GPUs; the same code runs on gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items for ``rank`` (sizes differ by at most one; earlier ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_cells(pos_per_img: Sequence[int], rank: int, world: int):
    """Image block of ``rank`` plus the matching slice of a flat per-cell array: ``(img_lo, img_hi, cell_lo, cell_hi)``."""
    lo, hi = shard_range(len(pos_per_img), rank, world)
    c_lo = int(sum(pos_per_img[:lo]))
    return lo, hi, c_lo, c_lo + int(sum(pos_per_img[lo:hi]))


def global_mean_loss(loss_sum: torch.Tensor, n_valid: int, group=None) -> Tuple[torch.Tensor, int]:
    """``sum_i F_i / n_valid`` over ALL ranks.  ``loss_sum`` is this rank's sum over its non-skipped images
    (keeps its autograd graph: the all-reduce acts on a detached copy and the local term is re-attached, so
    ``backward`` yields d(global mean)/d(local inputs) = local grads / global count)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return (loss_sum / max(n_valid, 1), n_valid)
    buf = torch.stack([loss_sum.detach().to(torch.float32), torch.tensor(float(n_valid), device=loss_sum.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    total, count = buf[0], int(round(float(buf[1])))
    count = max(count, 1)
    # value = global mean; gradient flows only through the local sum
    return ((total - loss_sum.detach()) + loss_sum) / count, count


def allreduce_student_grads(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> int:
    """One flat-bucket all-reduce of the gradients of ``params`` (student only: the teacher is frozen,
    reference ``train_kd.py:89,107``).  Returns the number of elements reduced."""
    grads: List[torch.Tensor] = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off
