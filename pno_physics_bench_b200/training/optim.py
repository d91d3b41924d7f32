"""Optimiser step of the trainer on device kernels (`pno_sumsq_accumulate`, `pno_adamw_step`, csrc/pno_uq.cu).

`FusedAdamW` has torch.optim.AdamW's update rule (decoupled weight decay, bias correction, `sqrt(v_hat) + eps`) and state
(`step`, `exp_avg`, `exp_avg_sq`), and folds `torch.nn.utils.clip_grad_norm_` (reference training/trainer.py:283-288) into
the same pass: ONE multi-tensor launch for the global gradient norm and ONE multi-tensor update launch that reads the norm from
device memory (`pno_sumsq_accumulate_multi`, `pno_adamw_step_multi`: a table of tensor pointers travels as a kernel parameter) --
no host sync, two launches and two passes over the 10 GB of parameter / gradient / moment state of the cfg-2 model instead of
torch's clip (norm + scale) and multi-pass foreach update.  Complex parameters are updated through their real view in
storage order (the spectral weights are permuted views of kernel-native storage)."""
from typing import Optional

import torch

from .. import _lib


from ..functional import storage_order as _storage_order  # noqa: E402


def _real(t):
    return torch.view_as_real(t) if t.is_complex() else t


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None, write_clipped_grads: bool = False):
        if lr < 0 or eps < 0 or weight_decay < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.write_clipped_grads = write_clipped_grads      # leave .grad scaled like clip_grad_norm_ does (one more write)
        self._gsq = None
        self.shard_group = None  # process group over which mode-sharded parameters (mode_parallel.py) are spread
        self._kl = {}            # id(param) -> 1 (spectral mean) / 2 (spectral log-variance)
        self._kl_gscale = None

    def set_analytic_kl(self, kl_params, gscale):
        """For the next step(): the KL term's gradient w.r.t. the spectral weights is added inside the optimiser kernels
        (it is analytic in the parameter: gs * mu, gs * (-0.5 (1 - exp(lv))), models.py:100-104) instead of being materialised
        and accumulated into .grad by a separate pass.  kl_params = (mu_0, lv_0, mu_1, lv_1, ...); gscale: 0-d float32 CUDA
        tensor (d loss / d KL).  `.grad` then holds the data-term gradient only."""
        self._kl = {}
        for k in range(0, len(kl_params), 2):
            self._kl[id(kl_params[k])] = 1
            self._kl[id(kl_params[k + 1])] = 2
        self._kl_gscale = gscale.detach().to(torch.float32).contiguous()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        work = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None and not self._kl.get(id(p), 0):
                    continue
                _lib.require_cuda(p, "parameter")
                if _real(p).dtype != torch.float32:
                    raise RuntimeError(f"FusedAdamW: float32 / complex64 parameters only, got {p.dtype}")
                pf = _storage_order(_real(p))
                if pf is None:
                    raise RuntimeError("FusedAdamW: parameter storage must be dense")
                kind = self._kl.get(id(p), 0)
                if p.grad is None and kind:
                    p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad
                if g.stride() != p.stride() or g.dtype != p.dtype:
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(p.grad)
                    if self.write_clipped_grads:
                        p.grad = g
                gf = _storage_order(_real(g))
                state = self.state[p]
                if not state:
                    state["step"] = torch.tensor(0.0)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                for key in ("exp_avg", "exp_avg_sq"):          # e.g. loaded from a reference (torch.optim.AdamW) checkpoint:
                    t = state[key]                             # same values, but laid out for the logical shape
                    if t.stride() != p.stride() or t.dtype != p.dtype or t.device != p.device:
                        state[key] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(t)
                work.append((group, state, pf, gf, _storage_order(_real(state["exp_avg"])),
                             _storage_order(_real(state["exp_avg_sq"])), kind, bool(getattr(p, "pno_mode_sharded", False))))
        if not work:
            return loss
        import ctypes
        dev = work[0][2].device

        def ptrs(ts):
            return (ctypes.c_void_p * len(ts))(*[_lib.ptr(t) for t in ts])

        n = len(work)
        ns = (ctypes.c_uint64 * n)(*[w[2].numel() for w in work])
        kinds = (ctypes.c_int * n)(*[w[6] for w in work])
        p_arr, g_arr = ptrs([w[2] for w in work]), ptrs([w[3] for w in work])
        kl_ptr = _lib.ptr(self._kl_gscale) if self._kl_gscale is not None else None
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        with _lib.on_device(work[0][2]):
            stream = _lib.stream_ptr(dev)
            if clip:
                if self._gsq is None or self._gsq.device != dev:
                    self._gsq = torch.zeros(1, dtype=torch.float64, device=dev)
                self._gsq.zero_()
                import torch.distributed as dist
                sharded = [k for k, w in enumerate(work) if w[7]]
                spread = bool(sharded) and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.shard_group) > 1
                if not spread:
                    # one launch: global gradient norm over every tensor (KL gradient of the spectral weights added on the fly)
                    _lib.check(lib.pno_sumsq_accumulate_multi(n, g_arr, ns, p_arr, kinds, kl_ptr, _lib.ptr(self._gsq), stream),
                               "pno_sumsq_accumulate_multi")
                else:
                    # mode-sharded parameters: every rank holds different shards, so their squared norms are summed over the
                    # group; the replicated parameters (identical, already averaged gradients) are counted once
                    for idx, reduce_it in ((sharded, True), ([k for k in range(n) if not work[k][7]], False)):
                        if not idx:
                            continue
                        acc = torch.zeros(1, dtype=torch.float64, device=dev)
                        c = len(idx)
                        pick = lambda arr, typ: (typ * c)(*[arr[k] for k in idx])      # noqa: E731
                        _lib.check(lib.pno_sumsq_accumulate_multi(c, pick(g_arr, ctypes.c_void_p), pick(ns, ctypes.c_uint64),
                                                                  pick(p_arr, ctypes.c_void_p), pick(kinds, ctypes.c_int), kl_ptr,
                                                                  _lib.ptr(acc), stream), "pno_sumsq_accumulate_multi")
                        if reduce_it:
                            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.shard_group)
                        self._gsq += acc
            # one launch per run of tensors that share hyper-parameters and step count (normally: one launch)
            start = 0
            while start < n:
                group, state = work[start][0], work[start][1]
                step_now = int(state["step"]) + 1
                end = start
                while end < n and work[end][0] is group and int(work[end][1]["step"]) + 1 == step_now:
                    end += 1
                for w in work[start:end]:
                    w[1]["step"] += 1
                b1, b2 = group["betas"]
                c = end - start
                sub = lambda arr, typ: (typ * c)(*arr[start:end])      # noqa: E731
                _lib.check(lib.pno_adamw_step_multi(
                    c, sub(p_arr, ctypes.c_void_p), sub(g_arr, ctypes.c_void_p), ptrs([w[4] for w in work[start:end]]),
                    ptrs([w[5] for w in work[start:end]]), sub(ns, ctypes.c_uint64), sub(kinds, ctypes.c_int), group["lr"], b1, b2,
                    group["eps"], group["weight_decay"], step_now, _lib.ptr(self._gsq) if clip else None,
                    float(self.max_grad_norm) if clip else 0.0, int(self.write_clipped_grads), kl_ptr, stream), "pno_adamw_step_multi")
                start = end
        self._kl, self._kl_gscale = {}, None          # one step only: the scale belongs to this step's loss
        return loss

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (device scalar; what clip_grad_norm_ returns)."""
        if self._gsq is None:
            raise RuntimeError("FusedAdamW.grad_norm() before a clipped step")
        return self._gsq.sqrt().float()[0]
