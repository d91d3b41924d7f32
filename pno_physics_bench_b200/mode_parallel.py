"""Mode-parallel (spectral tensor-parallel) training of the PNO.

Data parallelism replicates the spectral weights and all-reduces their gradients every step: 2.5 GB at BASELINE config 2 and
14.5 GB at config 5 (hid 256, 48 modes), where that exchange is several times the compute.  Here the RETAINED MODES are sharded
instead: rank r keeps mu / log_var, their gradients and their AdamW state for its own window of the 2*m1*m2 modes only (1/G of the
memory), and the batch stays sharded as in data parallelism.  Per spectral layer and direction the ranks exchange the truncated
spectrum twice (all-to-all: [modes, local batch] <-> [local modes, global batch]; 19 MB per rank at config 5) -- the transforms
and the 1x1 branches run on the local batch, the per-mode channel mix on the local modes for the GLOBAL batch:

    x_loc --pno_dft_fwd--> xhat[all modes][B_loc] --all-to-all--> xhat[my modes][B_glob] --pno_mix_fwd_window--> yhat[my modes][B_glob]
          --all-to-all--> yhat[all modes][B_loc] --pno_dft_inv (+ local branch, activation)--> y_loc

The weight noise is keyed by the GLOBAL mode index and the activation noise by the global batch index, so a sharded step computes
exactly the single-GPU step on the global batch.  The reference has no such path (closest: the layer-wise partitioning sketch in
scaling/distributed_inference_optimization.py:589-711); the 1x1 convolutions stay replicated and are averaged as in data
parallelism.  Tensor-core engines only (the mode-window kernels are tcgen05 kernels).
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from . import functional as F_
from .distributed import world


# ----------------------------------------------------------------------------------------------------------------------
# mode windows and the two exchanges
# ----------------------------------------------------------------------------------------------------------------------
def mode_windows(m1, m2, rank, world_size):
    """[(first global mode, count)] of this rank.  A window never crosses the boundary between the two weight blocks
    (mode m1*m2): one rank keeps both blocks as two windows; G > 1 ranks need G even and 2*m1*m2 divisible by G."""
    m1m2 = m1 * m2
    if world_size == 1:
        return [(0, m1m2), (m1m2, m1m2)]
    if world_size % 2 or (2 * m1m2) % world_size:
        raise ValueError(f"mode parallelism over {world_size} ranks needs an even rank count dividing the {2 * m1m2} retained modes")
    n = 2 * m1m2 // world_size
    return [(rank * n, n)]


def _a2a(send, group):
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def modes_from_batch(spec, group=None):
    """[P, M, B_loc, C] (all modes, local batch) -> [P, M/G, G*B_loc, C] (this rank's modes, global batch).  P = planes (re, im)."""
    _, ws = world(group)
    if ws == 1:
        return spec
    P, M, B, C = spec.shape
    n = M // ws
    send = spec.view(P, ws, n, B, C).permute(1, 0, 2, 3, 4).contiguous()          # [dest, P, n, B, C]
    recv = _a2a(send, group)                                                       # [src, P, n, B, C]
    return recv.permute(1, 2, 0, 3, 4).reshape(P, n, ws * B, C)


def batch_from_modes(spec, group=None):
    """[P, M/G, G*B_loc, C] (this rank's modes, global batch) -> [P, M, B_loc, C] (all modes, local batch)."""
    _, ws = world(group)
    if ws == 1:
        return spec
    P, n, BG, C = spec.shape
    B = BG // ws
    send = spec.view(P, n, ws, B, C).permute(2, 0, 1, 3, 4).contiguous()           # [dest, P, n, B, C]
    recv = _a2a(send, group)                                                       # [src (= owner of the modes), P, n, B, C]
    return recv.permute(1, 0, 2, 3, 4).reshape(P, ws * n, B, C)


MIX_MAX_BATCH = 128          # batch rows one launch of the forward / input-gradient channel-mix kernels takes at most (their MMA N)


def _mix_chunk_rows(stage, engine, plan_shape, bg, ci, co):
    """Largest batch chunk <= min(bg, MIX_MAX_BATCH) the channel-mix kernel `stage` tiles for this engine (the fp32-parity engine
    keeps hi / lo copies of its operand tiles: 64 rows; single pass: 128)."""
    h, w, m1, m2 = plan_shape
    for rows in (128, 96, 64, 32, 16, 8):
        if rows <= MIX_MAX_BATCH and rows < bg and _lib.engine_supports(stage, engine & 0xFF, h, w, m1, m2, rows, ci, co):
            return rows
    return min(bg, MIX_MAX_BATCH)


def _mix_window_batched(fn, name, plan, engine, src, bg, ci, co, mu_ptr, lv_ptr, key, first, count, dst, st, plan_shape):
    """One mode window of pno_mix_fwd_window / pno_mix_bwd_dx_window over a global batch of any size: src / dst are the
    [2, count, bg, C] slices of this window.  Up to MIX_MAX_BATCH rows go through in place; a larger global batch (8 ranks x 32) is
    cut into batch chunks, each staged through a dense copy (the kernels take dense [mode][b][c] spectra)."""
    seed, sample_idx, site = key
    stage = "mix_fwd" if name == "pno_mix_fwd_window" else "mix_bwd_dx"
    if bg <= MIX_MAX_BATCH and _lib.engine_supports(stage, engine & 0xFF, *plan_shape, bg, ci, co):
        _lib.check(fn(plan, engine, src[0].data_ptr(), src[1].data_ptr(), bg, ci, co, mu_ptr, lv_ptr, seed, sample_idx, site, first,
                      count, dst[0].data_ptr(), dst[1].data_ptr(), st), name)
        return
    rows = _mix_chunk_rows(stage, engine, plan_shape, bg, ci, co)
    for b0 in range(0, bg, rows):
        b1 = min(bg, b0 + rows)
        part = src[:, :, b0:b1].contiguous()
        out = torch.empty(2, count, b1 - b0, dst.shape[3], dtype=torch.float32, device=src.device)
        _lib.check(fn(plan, engine, part[0].data_ptr(), part[1].data_ptr(), b1 - b0, ci, co, mu_ptr, lv_ptr, seed, sample_idx, site, first,
                      count, out[0].data_ptr(), out[1].data_ptr(), st), name)
        dst[:, :, b0:b1].copy_(out)


# ----------------------------------------------------------------------------------------------------------------------
# the sharded spectral layer
# ----------------------------------------------------------------------------------------------------------------------
class _ModeParallelSpectralFn(torch.autograd.Function):
    """act(spectral(x; W) + addend) with W sharded by modes over `group` (see the module docstring)."""

    @staticmethod
    def forward(ctx, x, mu, lv, addend, meta):
        lib = _lib.load()
        (m1, m2, windows, seed, sample_idx, site, act, engine, group, sampled, grad_mode) = meta
        F_._check_x(x, mu.shape[1])
        _lib.require_param(mu, x, "spectral weight shard", torch.complex64)
        _lib.require_param(lv, x, "spectral log-variance shard", torch.float32)
        x = x.contiguous()
        b, ci, h, w = x.shape
        co = mu.shape[2]
        M = 2 * m1 * m2
        plan = _lib.plan(h, w, m1, m2, x.device)
        st = _lib.stream_ptr(x.device)
        needs_grad = grad_mode and any(ctx.needs_input_grad[:4])
        with _lib.on_device(x):
            xh = torch.empty(2, M, b, ci, dtype=torch.float32, device=x.device)
            _lib.check(lib.pno_dft_fwd(plan, engine, x.data_ptr(), b, ci, 0, xh[0].data_ptr(), xh[1].data_ptr(), st), "pno_dft_fwd")
            xs = modes_from_batch(xh, group)                                        # [2, n_local, B_glob, Ci]
            bg = xs.shape[2]
            ys = torch.empty(2, xs.shape[1], bg, co, dtype=torch.float32, device=x.device)
            mu_r = torch.view_as_real(mu.detach())
            off = 0
            for first, count in windows:
                _mix_window_batched(lib.pno_mix_fwd_window, "pno_mix_fwd_window", plan, engine, xs[:, off:off + count], bg, ci, co,
                                    mu_r[off:off + count].data_ptr(), lv[off:off + count].data_ptr() if sampled else None,
                                    (seed, sample_idx, site), first, count, ys[:, off:off + count], st, (h, w, m1, m2))
                off += count
            yh = batch_from_modes(ys, group)                                        # [2, M, B_loc, Co]
            y = torch.empty(b, co, h, w, dtype=torch.float32, device=x.device)
            pre = torch.empty_like(y) if (needs_grad and act != _lib.ACT_NONE) else None
            add = addend.contiguous() if addend is not None else None
            _lib.check(lib.pno_dft_inv(plan, engine, yh[0].data_ptr(), yh[1].data_ptr(), b, co, 0, _lib.ptr(add), act, y.data_ptr(),
                                       _lib.ptr(pre), st), "pno_dft_inv")
        if needs_grad:
            ctx.save_for_backward(mu, lv, pre, xs)
            ctx.meta = meta
            ctx.shape = (b, ci, co, h, w, bg, addend is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        mu, lv, pre, xs = ctx.saved_tensors
        (m1, m2, windows, seed, sample_idx, site, act, engine, group, sampled, _) = ctx.meta
        b, ci, co, h, w, bg, has_add = ctx.shape
        _, ws = world(group)
        M = 2 * m1 * m2
        dev = g.device
        plan = _lib.plan(h, w, m1, m2, dev)
        st = _lib.stream_ptr(dev)
        g = g.contiguous()
        with _lib.on_device(g):
            if act != _lib.ACT_NONE:
                gz = torch.empty_like(g)
                _lib.check(lib.pno_act_bwd(pre.data_ptr(), g.data_ptr(), g.numel(), act, gz.data_ptr(), st), "pno_act_bwd")
            else:
                gz = g
            gh = torch.empty(2, M, b, co, dtype=torch.float32, device=dev)
            _lib.check(lib.pno_dft_fwd(plan, engine, gz.data_ptr(), b, co, 1, gh[0].data_ptr(), gh[1].data_ptr(), st), "pno_dft_fwd")
            gs = modes_from_batch(gh, group)                                        # [2, n_local, B_glob, Co]
            need_w = ctx.needs_input_grad[1] or (sampled and ctx.needs_input_grad[2])
            need_lv = sampled and ctx.needs_input_grad[2]
            dmu = torch.empty(mu.shape + (2,), dtype=torch.float32, device=dev) if need_w else None
            dlv = torch.empty(lv.shape, dtype=torch.float32, device=dev) if need_lv else None
            dxs = torch.empty(2, xs.shape[1], bg, ci, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
            mu_r = torch.view_as_real(mu.detach())
            off = 0
            for first, count in windows:
                sl = slice(off, off + count)
                if need_w:
                    _lib.check(lib.pno_mix_bwd_dw_window(plan, engine, xs[0, sl].data_ptr(), xs[1, sl].data_ptr(), gs[0, sl].data_ptr(),
                                                         gs[1, sl].data_ptr(), bg, ci, co, lv[sl].data_ptr() if sampled else None, seed,
                                                         sample_idx, site, first, count, dmu[sl].data_ptr(),
                                                         dlv[sl].data_ptr() if need_lv else None, st), "pno_mix_bwd_dw_window")
                if dxs is not None:
                    _mix_window_batched(lib.pno_mix_bwd_dx_window, "pno_mix_bwd_dx_window", plan, engine, gs[:, sl], bg, ci, co,
                                        mu_r[sl].data_ptr(), lv[sl].data_ptr() if sampled else None, (seed, sample_idx, site), first,
                                        count, dxs[:, sl], st, (h, w, m1, m2))
                off += count
            dx = None
            if dxs is not None:
                dxh = batch_from_modes(dxs, group)                                  # [2, M, B_loc, Ci]
                dx = torch.empty(b, ci, h, w, dtype=torch.float32, device=dev)
                _lib.check(lib.pno_dft_inv(plan, engine, dxh[0].data_ptr(), dxh[1].data_ptr(), b, ci, 1, None, _lib.ACT_NONE,
                                           dx.data_ptr(), None, st), "pno_dft_inv")
            # every rank's loss is the mean over ITS batch shard; the training loss is the mean over ranks: the weight gradients
            # (summed over the global batch here) carry that 1/G, like the averaged gradients of the replicated parameters
            if ws > 1:
                if dmu is not None:
                    dmu.mul_(1.0 / ws)
                if dlv is not None:
                    dlv.mul_(1.0 / ws)
        return (dx, torch.view_as_complex(dmu) if (dmu is not None and ctx.needs_input_grad[1]) else None, dlv,
                gz if (has_add and ctx.needs_input_grad[3]) else None, None)


class ModeParallelPNO(nn.Module):
    """A ProbabilisticNeuralOperator whose spectral weights are sharded by modes over `group`.

    Built from a full model that is identical on every rank (same seed or same checkpoint): this rank's mode windows are copied
    into `spectral_mu[l]` / `spectral_log_var[l]` ([local modes, Ci, Co], kernel-native) and the full spectral parameters are
    released.  The 1x1 convolutions are shared with (and owned by) the wrapped model and stay replicated.  Same forward contract as
    the model (`forward(x, sample, return_kl)`, `kl_divergence`, `kl_params`), so PNOTrainer drives it unchanged."""

    def __init__(self, model, group=None, engine="tf32x3"):
        super().__init__()
        rank, ws = world(group)
        self.group, self.rank, self.world_size = group, rank, ws
        self.engine = _lib.ENGINES[engine] if isinstance(engine, str) else int(engine)
        if (self.engine & 0xFF) == _lib.ENGINE_FP32:
            raise ValueError("mode parallelism runs on the tensor-core engines ('tf32x3' or 'tf32')")
        self.model = model
        self.modes, self.hidden_dim, self.num_layers = model.modes, model.hidden_dim, model.num_layers
        self.windows = mode_windows(model.modes, model.modes, rank, ws)
        m1m2 = model.modes * model.modes
        mus, lvs = [], []
        for layer in model.fourier_layers:
            full_mu = torch.cat([F_.native_spectral(w.detach())[0].reshape(m1m2, *w.shape[:2])
                                 for w in (layer.weights1, layer.weights2)])                      # [2 m1m2, Ci, Co]
            full_lv = torch.cat([F_.native_spectral(w.detach())[0].reshape(m1m2, *w.shape[:2])
                                 for w in (layer.weights1_log_var, layer.weights2_log_var)])
            mus.append(nn.Parameter(torch.cat([full_mu[f:f + c] for f, c in self.windows]).clone()))
            lvs.append(nn.Parameter(torch.cat([full_lv[f:f + c] for f, c in self.windows]).clone()))
        self.spectral_mu, self.spectral_log_var = nn.ParameterList(mus), nn.ParameterList(lvs)
        for p in [*self.spectral_mu, *self.spectral_log_var]:
            p.pno_mode_sharded = True           # GradientSynchronizer / FusedAdamW: not replicated, never all-reduced
        model.fourier_layers = nn.ModuleList()  # release the full spectral parameters (the point of the exercise)

    def sharded_parameters(self):
        return [*self.spectral_mu, *self.spectral_log_var]

    def forward(self, x, sample: bool = True, return_kl: bool = False):
        from .models import _ACT_CODES
        m = self.model
        m._validate(x)
        stochastic = sample and self.training
        sc = F_.scope_or_fresh() if stochastic else None
        act = _ACT_CODES[m.activation_name]
        kw = dict(seed=sc.seed, sample_idx=sc.sample_idx) if stochastic else {}
        boff = sc.batch_offset if stochastic else 0
        eng = self.engine
        h = F_.local_branch(x, m.lift_mean.weight, m.lift_mean.bias, m.lift_log_var.weight if stochastic else None,
                            m.lift_log_var.bias if stochastic else None, site=0, batch_offset=boff, engine=eng, **kw)
        for l in range(self.num_layers):
            local = F_.local_branch(h, m.linear_mean[l].weight, m.linear_mean[l].bias,
                                    m.linear_log_var[l].weight if stochastic else None,
                                    m.linear_log_var[l].bias if stochastic else None, site=2 + 3 * l, batch_offset=boff, engine=eng,
                                    **kw)
            meta = (self.modes, self.modes, self.windows, sc.seed if stochastic else 0, sc.sample_idx if stochastic else 0, 1 + 3 * l,
                    act, eng, self.group, stochastic, torch.is_grad_enabled())
            h = _ModeParallelSpectralFn.apply(h, self.spectral_mu[l], self.spectral_log_var[l], local, meta)
        if return_kl:
            mean = F_.local_branch(h, m.project_mean.weight, m.project_mean.bias, engine=eng)
            log_var = F_.local_branch(h, m.project_log_var.weight, m.project_log_var.bias, engine=eng)
            return mean, torch.clamp(log_var, -10.0, 10.0), self.kl_divergence()
        kw = dict(seed=sc.seed, sample_idx=sc.sample_idx, batch_offset=sc.batch_offset) if stochastic else {}
        return F_.local_branch(h, m.project_mean.weight, m.project_mean.bias, m.project_log_var.weight if stochastic else None,
                               m.project_log_var.bias if stochastic else None, site=1 + 3 * self.num_layers, engine=eng, **kw)

    def kl_divergence(self):
        """models.py:381-389 over the sharded weights: the local shards' KL, summed over the group (a value; its gradient with
        respect to the weights is analytic and is added by FusedAdamW.set_analytic_kl, as in the unsharded trainer step)."""
        lib = _lib.load()
        dev = self.spectral_mu[0].device
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        with _lib.on_device(acc):
            for mu, lv in zip(self.spectral_mu, self.spectral_log_var):
                _lib.check(lib.pno_kl_fwd(torch.view_as_real(mu.detach()).data_ptr(), lv.detach().data_ptr(), lv.numel(), acc.data_ptr(),
                                          _lib.stream_ptr(dev)), "pno_kl_fwd")
        if self.world_size > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        return acc.to(torch.float32)

    compute_kl_divergence = kl_divergence

    def kl_params(self):
        out = []
        for mu, lv in zip(self.spectral_mu, self.spectral_log_var):
            out += [mu, lv]
        return out

    @torch.no_grad()
    def gather_state_dict(self):
        """The reference-layout state_dict (full [Ci,Co,m1,m2] spectral parameters) assembled from every rank's shards."""
        sd = {k: v for k, v in self.model.state_dict().items()}
        m1m2, m = self.modes * self.modes, self.modes
        for l, (mu, lv) in enumerate(zip(self.spectral_mu, self.spectral_log_var)):
            parts = []
            for t in (torch.view_as_real(mu.detach()).contiguous(), lv.detach().contiguous()):
                if self.world_size > 1:
                    bufs = [torch.empty_like(t) for _ in range(self.world_size)]
                    dist.all_gather(bufs, t, group=self.group)
                    t = torch.cat(bufs)
                parts.append(t)
            full_mu, full_lv = torch.view_as_complex(parts[0]), parts[1]                       # [2 m1m2, Ci, Co]
            for blk, name in enumerate(("weights1", "weights2")):
                wmu = full_mu[blk * m1m2:(blk + 1) * m1m2].reshape(m, m, *full_mu.shape[1:]).permute(2, 3, 0, 1)
                wlv = full_lv[blk * m1m2:(blk + 1) * m1m2].reshape(m, m, *full_lv.shape[1:]).permute(2, 3, 0, 1)
                sd[f"fourier_layers.{l}.{name}"] = wmu.contiguous()
                sd[f"fourier_layers.{l}.{name}_log_var"] = wlv.contiguous()
        return sd
