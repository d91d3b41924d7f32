"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed over NCCL.

* MC predictive loop: the S samples are split into contiguous ranges of GLOBAL sample indices, parameters are
  replicated, and the only exchange is one all-reduce of the two fp64 moment buffers (sum y, sum y^2).
* Data-parallel training: every rank uses the same weight-noise key, activation-space noise is indexed by the global
  batch index, and gradients are averaged with bucketed all-reduces.

The reference has no such path (its DDP / all-gather helpers in distributed_training.py:27-142 and
research/distributed_pno.py:206-214, 327 are never reached from PNOTrainer or predict_with_uncertainty).
"""
import torch
import torch.distributed as dist

from . import functional as F_


def _active(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def mc_shard(num_samples, group=None):
    rank, ws = world(group)
    return shard_range(num_samples, rank, ws)


def shared_seed(seed, device, group=None):
    """The Philox seed of one MC pass / training step: the caller's, or a fresh one drawn on rank 0, same on all ranks."""
    if seed is not None:
        return int(seed)            # an explicit seed is the caller's promise that every rank passes the same one: no exchange
    seed = F_.fresh_seed()
    if _active(group):
        t = torch.tensor([int(seed) & (2 ** 62 - 1)], dtype=torch.int64, device=device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        seed = int(t.item())
    return int(seed)


def allreduce_moments(s1, s2, group=None):
    """Sum the per-rank moment buffers in place (one fused NCCL all-reduce of [2, ...] fp64)."""
    if not _active(group):
        return
    # the two buffers are the halves of one [2, ...] allocation wherever this package creates them: reduce it in place
    if (s2.data_ptr() == s1.data_ptr() + s1.numel() * s1.element_size() and s1.is_contiguous() and s2.is_contiguous()):
        buf = torch.as_strided(s1, (2 * s1.numel(),), (1,))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        return
    buf = torch.stack([s1, s2])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    s1.copy_(buf[0])
    s2.copy_(buf[1])


def allreduce_gradients(params, group=None, bucket_bytes=256 << 20):
    """Average .grad over ranks with large flat buckets (few launches; NVSwitch bandwidth is not per-link bound).

    Complex gradients travel as their real views.  Parameters without a gradient contribute zeros so that all ranks
    issue identical collectives."""
    if not _active(group):
        return
    ws = dist.get_world_size(group)
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(ws)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        bucket, size = [], 0

    for p in params:
        if not p.requires_grad:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        g = torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad
        if not g.is_contiguous():
            # kernel-native spectral grads are permuted views of contiguous storage: reduce the storage order
            g = g.permute(*sorted(range(g.dim()), key=lambda d: -g.stride(d)))
        nbytes = g.numel() * g.element_size()
        if size + nbytes > bucket_bytes:
            flush()
        bucket.append(g)
        size += nbytes
    flush()
