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


def _grad_storage(p):
    """The gradient of p as a flat real view of its storage (spectral gradients are permuted views of kernel-native storage;
    complex ones travel as (re, im) pairs), or None when the gradient is not dense."""
    g = p.grad
    g = torch.view_as_real(g) if g.is_complex() else g
    return F_.storage_order(g)


def _avg_inplace(t, group, async_op=False):
    """Average t over the group in place.  NCCL averages inside the collective (ncclAvg); gloo sums and the division follows."""
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op), None
    work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return work, dist.get_world_size(group)


class GradientSynchronizer:
    """Data-parallel gradient averaging overlapped with backward (the reference's side module relies on DDP's bucket overlap,
    distributed_training.py:96-101).

    Every parameter gets a post-accumulate-grad hook.  A gradient of at least `large_bytes` (the spectral weights: 210-420 MB
    each at cfg 2) is all-reduced IN ITS OWN STORAGE the moment autograd has produced it -- asynchronously on NCCL's stream, while
    backward continues on the compute stream -- with no flattening copy.  The small ones (1x1 conv weights and biases) are
    gathered into one flat bucket at `finish()`.  All ranks run the same autograd graph, so the collectives are issued in the
    same order everywhere.  Parameters that received no gradient take no part (a single-GPU run would skip them, too)."""

    def __init__(self, params, group=None, large_bytes=4 << 20):
        self.group, self.large_bytes = group, large_bytes
        # mode-sharded spectral parameters (mode_parallel.ModeParallelPNO) live on one rank each: nothing to average
        self.params = [p for p in params if p.requires_grad and not getattr(p, "pno_mode_sharded", False)]
        self.pending, self.small = [], []
        self.enabled = True
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def has_large(self):
        """True when some parameter's gradient is all-reduced on its own, concurrently with backward (data parallelism over
        replicated spectral weights; with mode-sharded weights only the small 1x1-conv gradients are left)."""
        if not (self.enabled and _active(self.group)):
            return False
        return any(p.numel() * p.element_size() >= self.large_bytes for p in self.params)

    def _on_grad(self, p):
        if not self.enabled or not _active(self.group) or p.grad is None:
            return
        flat = _grad_storage(p)
        if flat is not None and flat.numel() * flat.element_size() >= self.large_bytes:
            work, div = _avg_inplace(flat, self.group, async_op=True)
            self.pending.append((work, flat, div))
        else:
            self.small.append(p)

    def finish(self):
        """Reduce the small gradients, then make the compute stream wait for every outstanding collective."""
        if self.small:
            views = []
            for p in self.small:
                g = torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad
                views.append(g)
            flat = torch.cat([g.reshape(-1) for g in views])
            work, div = _avg_inplace(flat, self.group, async_op=False)
            if div:
                flat.div_(div)
            off = 0
            for g in views:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
            self.small = []
        for work, flat, div in self.pending:
            work.wait()
            if div:
                flat.div_(div)
        self.pending = []

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def allreduce_gradients(params, group=None, large_bytes=4 << 20):
    """Average .grad over ranks after backward has finished (the un-overlapped form of GradientSynchronizer, for callers that own
    their backward pass): large gradients in their own storage, small ones through one flat bucket.  Parameters without a
    gradient are skipped."""
    if not _active(group):
        return
    sync = GradientSynchronizer([], group, large_bytes)
    for p in params:
        if p.requires_grad and p.grad is not None:
            sync._on_grad(p)
    sync.finish()
