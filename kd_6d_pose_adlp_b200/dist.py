"""Multi-GPU plumbing for the OT distillation path: one process per GPU, images sharded across ranks.

Every image's OT problems are independent (reference ``losses/loss_libs.py:22-50`` never mixes images), so the
path shards by image with NO data-path collective: each rank runs the fused kernel on its block of images.
Two exchanges exist around it:

* :func:`global_mean_loss` -- all-reduce of ``(sum_i F_i, n_valid)`` (two scalars) when the caller wants the mean
  over the non-empty images of the GLOBAL batch instead of the reference's per-process mean
  (``losses/kd_loss.py:99-100`` divides by the local count; its own multi-GPU runs never exchange it);
* :class:`GradBucket` / :func:`allreduce_student_grads` -- the data-parallel mean of the STUDENT's parameter
  gradients (the teacher is frozen, reference ``train_kd.py:89,107``).  The reference wraps the model in DDP and
  immediately unwraps it (``libs/train_libs.py:124-130``) on a gloo group (``train_kd.py:50``), so it performs no
  gradient exchange at all; this is the collective a correct data-parallel run needs.  The bucket is ONE persistent
  flat buffer whose slices ARE the parameters' ``.grad`` tensors: backward accumulates straight into it and the
  all-reduce runs on it in place -- no gather copy before, no scatter copy after (2.3 M fp32 for darknet_tiny_h,
  8.5 M for darknet_tiny).

Backend: NCCL over NVLink/NVSwitch on GPUs (``ReduceOp.AVG``: the division happens inside the collective); the same
code runs on gloo for the CPU tests (SUM followed by one in-place scale).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

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


def _initialised(group=None) -> bool:
    return dist.is_available() and dist.is_initialized()


def global_mean_loss(loss_sum: torch.Tensor, n_valid: int, group=None) -> Tuple[torch.Tensor, int]:
    """``sum_i F_i / n_valid`` over ALL ranks.  ``loss_sum`` is this rank's sum over its non-skipped images
    (keeps its autograd graph: the all-reduce acts on a detached copy and the local term is re-attached, so
    ``backward`` yields d(global mean)/d(local inputs) = local grads / global count)."""
    if not _initialised(group) or dist.get_world_size(group) == 1:
        return (loss_sum / max(n_valid, 1), n_valid)
    buf = torch.stack([loss_sum.detach().to(torch.float32), torch.tensor(float(n_valid), device=loss_sum.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    total, count = buf[0], int(round(float(buf[1])))
    count = max(count, 1)
    # value = global mean; gradient flows only through the local sum
    return ((total - loss_sum.detach()) + loss_sum) / count, count


class GradBucket:
    """Persistent flat gradient buffer of a set of parameters.

    Every parameter with ``requires_grad`` gets ``p.grad = flat[offset : offset + numel].view_as(p)`` -- also the ones a
    given step does not touch (their slice stays zero), so every rank reduces the same number of elements in the same
    order whatever subset of heads produced gradients on it.  Use :meth:`zero` instead of
    ``optimizer.zero_grad(set_to_none=True)`` (which would drop the views; ``set_to_none=False`` is fine).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype: torch.dtype = torch.float32):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no parameter requires a gradient")
        dev = self.params[0].device
        if any(p.device != dev for p in self.params):
            raise ValueError("GradBucket: all parameters must live on one device")
        self.offsets: List[int] = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) & ~3  # 16-byte aligned slices
        self.flat = torch.zeros(n, dtype=dtype, device=dev)
        self.numel = sum(p.numel() for p in self.params)
        self.attach()

    def attach(self, keep_values: bool = True) -> None:
        """(Re-)point every ``.grad`` at its slice; existing gradient values are carried over once."""
        for p, o in zip(self.params, self.offsets):
            view = self.flat[o:o + p.numel()].view_as(p)
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr() and keep_values:
                view.copy_(p.grad)
            p.grad = view

    def attached(self) -> bool:
        return all(p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + o * self.flat.element_size()
                   for p, o in zip(self.params, self.offsets))

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce(self, average: bool = True, group=None, async_op: bool = False):
        """In-place all-reduce of the whole bucket (one collective).  Returns the work handle when ``async_op``."""
        if not _initialised(group) or dist.get_world_size(group) == 1:
            return None
        if not self.attached():
            self.attach()
        world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        if average and backend == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
            if async_op:
                work.wait()
                work = None
            self.flat.div_(world)
        return work


_buckets = {}


def allreduce_student_grads(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> int:
    """All-reduce the gradients of ``params`` in one flat bucket and return the number of gradient elements.

    The bucket is created on first use for this parameter set and kept: from then on ``.grad`` of every parameter is a
    view of it and the call costs exactly one collective.  Parameters without a gradient on this rank take part with
    zeros, so ranks that exercised different heads still issue the same collective."""
    plist = [p for p in params if p.requires_grad]
    if not plist:
        return 0
    key = tuple(id(p) for p in plist)
    bucket: Optional[GradBucket] = _buckets.get(key)
    if bucket is None:
        bucket = GradBucket(plist)
        _buckets.clear()  # one live parameter set per process is the use case; do not pin old models
        _buckets[key] = bucket
    elif not bucket.attached():
        bucket.attach()
    bucket.allreduce(average=average, group=group)
    return bucket.numel
