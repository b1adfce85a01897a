"""world_size-2 gloo tests of the multi-GPU host logic (image sharding, global mean, gradient bucket)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kd_6d_pose_adlp_b200.dist import GradBucket, allreduce_student_grads, global_mean_loss, shard_cells, shard_range
from kd_6d_pose_adlp_b200.synthetic import cu_seqlens, ot_batch
from oracle import sinkhorn_analytic


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = ot_batch(7, seed=3, p_empty_teacher=0.3)
    lo, hi, c_lo, c_hi = shard_cells(batch["pos_per_img"], rank, world)
    _, _, t_lo, t_hi = shard_cells(batch["pos_per_img_t"], rank, world)
    # per-image losses of this rank's block (CPU oracle stands in for the kernel: host logic under test)
    o = sinkhorn_analytic.kdot_fwd_bwd_f64(batch["xs"][c_lo:c_hi], batch["ws"][c_lo:c_hi], batch["xt"][t_lo:t_hi],
                                           batch["wt"][t_lo:t_hi], cu_seqlens(batch["pos_per_img"][lo:hi]),
                                           cu_seqlens(batch["pos_per_img_t"][lo:hi]), 8, 2)
    local = torch.tensor(o["loss_per_img"].sum(), dtype=torch.float32, requires_grad=True)
    mean, count = global_mean_loss(local * 1.0, int(o["valid"].sum()))
    mean.backward()
    # gradient bucket: two "parameters" with rank-dependent grads
    p1, p2 = torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 3))
    p1.grad, p2.grad = torch.full((5,), float(rank + 1)), torch.full((2, 3), 10.0 * (rank + 1))
    n = allreduce_student_grads([p1, p2], average=True)
    # persistent bucket: .grad tensors ARE slices of the flat buffer, backward accumulates into them in place, a
    # parameter this rank never touched (p5 on rank 1) takes part with zeros so both ranks reduce the same buffer
    p3, p4, p5 = (torch.nn.Parameter(torch.ones(4)), torch.nn.Parameter(torch.ones(3, 2)), torch.nn.Parameter(torch.ones(6)))
    bucket = GradBucket([p3, p4, p5])
    flat_ptr = bucket.flat.data_ptr()
    loss = (p3 * (rank + 1)).sum() + (p4 * 2.0).sum() + ((p5 * 3.0).sum() if rank == 0 else 0.0)
    loss.backward()
    views_kept = bucket.attached() and p3.grad.data_ptr() == flat_ptr
    bucket.allreduce(average=True)
    same_buffer = bucket.flat.data_ptr() == flat_ptr and bucket.attached()
    q.put((rank, float(mean), count, float(local.grad), n, p1.grad.tolist(), p2.grad.flatten().tolist(),
           views_kept, same_buffer, p3.grad.tolist(), p4.grad.flatten().tolist(), p5.grad.tolist()))
    dist.destroy_process_group()


def test_global_mean_and_grad_bucket_world2():
    world, port = 2, 29000 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = ot_batch(7, seed=3, p_empty_teacher=0.3)
    o = sinkhorn_analytic.kdot_fwd_bwd_f64(batch["xs"], batch["ws"], batch["xt"], batch["wt"],
                                           cu_seqlens(batch["pos_per_img"]), cu_seqlens(batch["pos_per_img_t"]), 8, 2)
    want = o["loss_per_img"].sum() / o["valid"].sum()
    for rank, mean, count, g, n, g1, g2, views_kept, same_buffer, g3, g4, g5 in res:
        assert count == int(o["valid"].sum())
        assert abs(mean - want) < 1e-5 * abs(want)      # every rank holds the single-process mean
        assert abs(g - 1.0 / count) < 1e-7               # d(mean)/d(local sum) = 1 / global count
        assert n == 11 and np.allclose(g1, 1.5) and np.allclose(g2, 15.0)
        assert views_kept and same_buffer               # zero-copy: backward wrote into the bucket, all-reduce ran in place
        assert np.allclose(g3, 1.5) and np.allclose(g4, 2.0) and np.allclose(g5, 1.5)   # p5: (3 + 0) / 2
