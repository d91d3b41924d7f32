"""torch.autograd glue around the C ABI (include/pno_b200.h): every tensor op of the reference's hot path
(models.py:63-104, 264-275, 293-306) is one or two calls into libpno_b200.so on torch's current CUDA stream.

Noise is never drawn with torch.randn_like: each call carries a Philox (seed, sample_idx, site) triple and the kernels
draw eps in place (forward) and re-draw it (backward).  `materialize_*` expose the same numbers for parity tests.
"""
import contextlib
import threading

import torch

from . import _lib

_tls = threading.local()


# ----------------------------------------------------------------------------------------------------------------------
# noise scope
# ----------------------------------------------------------------------------------------------------------------------
class NoiseScope:
    """(seed, sample_idx, batch_offset) used by every stochastic site of the forward passes run inside the scope.

    batch_offset is the GLOBAL index of the first local batch item, so a batch sharded across GPUs draws exactly the
    activation-space noise the unsharded batch would."""

    def __init__(self, seed, sample_idx=0, batch_offset=0):
        self.seed, self.sample_idx, self.batch_offset = int(seed) & (2 ** 64 - 1), int(sample_idx), int(batch_offset)


@contextlib.contextmanager
def noise_scope(seed, sample_idx=0, batch_offset=0):
    prev = getattr(_tls, "scope", None)
    _tls.scope = NoiseScope(seed, sample_idx, batch_offset)
    try:
        yield _tls.scope
    finally:
        _tls.scope = prev


def current_scope():
    return getattr(_tls, "scope", None)


def fresh_seed():
    """A new 63-bit seed from torch's global CPU generator (so torch.manual_seed makes runs repeatable)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def scope_or_fresh():
    s = current_scope()
    return s if s is not None else NoiseScope(fresh_seed(), 0, 0)


# ----------------------------------------------------------------------------------------------------------------------
# parameter layout
# ----------------------------------------------------------------------------------------------------------------------
def native_spectral(w):
    """[Ci,Co,m1,m2] logical tensor -> ([m1,m2,Ci,Co] contiguous tensor sharing storage if possible, was_view)."""
    p = w.permute(2, 3, 0, 1)
    if p.is_contiguous():
        return p, True
    return p.contiguous(), False


def alloc_spectral(ci, co, m1, m2, dtype, device=None):
    """Kernel-native [m1,m2,Ci,Co] storage exposed with the reference's logical shape [Ci,Co,m1,m2]."""
    return torch.empty(m1, m2, ci, co, dtype=dtype, device=device).permute(2, 3, 0, 1)


def storage_order(t):
    """Flat view of a dense (non-overlapping, gap-free) real tensor in memory order, or None (complex: pass view_as_real)."""
    dims = sorted((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1)
    expect = 1
    for st, sz in dims:
        if st != expect:
            return None
        expect *= sz
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


def _check_x(x, ci):
    _lib.require_cuda(x, "input")
    if x.dtype != torch.float32:
        # reference: out_ft is hard-wired to cfloat, so fp64/fp16/bf16 inputs raise RuntimeError
        raise RuntimeError(f"expected a float32 input, got {x.dtype}")
    if x.dim() != 4:
        raise RuntimeError(f"expected a 4-D [B,C,H,W] input, got {x.dim()}-D")
    if x.size(1) != ci:
        raise RuntimeError(f"expected {ci} input channels, got {x.size(1)}")


# ----------------------------------------------------------------------------------------------------------------------
# spectral convolution (+ fused block epilogue)
# ----------------------------------------------------------------------------------------------------------------------
class _SpectralConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, w2, lv1, lv2, addend, seed, sample_idx, site, act, engine, grad_mode):
        lib = _lib.load()
        ci, co, m1, m2 = w1.shape
        _check_x(x, ci)
        for name, t, dt in (("weights1", w1, torch.complex64), ("weights2", w2, torch.complex64),
                            ("weights1_log_var", lv1, torch.float32), ("weights2_log_var", lv2, torch.float32),
                            ("addend", addend, torch.float32)):
            _lib.require_param(t, x, name, dt)
        x = x.contiguous()
        b, _, h, w = x.shape
        plan = _lib.plan(h, w, m1, m2, x.device)
        n1, _ = native_spectral(w1.detach())
        n2, _ = native_spectral(w2.detach())
        sampled = lv1 is not None
        l1 = native_spectral(lv1.detach())[0] if sampled else None
        l2 = native_spectral(lv2.detach())[0] if sampled else None
        add = addend.contiguous() if addend is not None else None
        needs_grad = grad_mode and any(ctx.needs_input_grad)   # under no_grad nothing is produced or kept for backward
        y = torch.empty(b, co, h, w, dtype=torch.float32, device=x.device)
        pre = torch.empty_like(y) if (needs_grad and act != _lib.ACT_NONE) else None
        ws = torch.empty(lib.pno_spectral_workspace_bytes(plan, b, ci, co) // 4, dtype=torch.float32, device=x.device)
        with _lib.on_device(x):
            _lib.check(lib.pno_spectral_fwd(plan, engine, x.data_ptr(), b, ci, co, n1.data_ptr(), n2.data_ptr(), _lib.ptr(l1),
                                            _lib.ptr(l2), seed, sample_idx, site, _lib.ptr(add), act, y.data_ptr(),
                                            _lib.ptr(pre), ws.data_ptr(), _lib.stream_ptr(x.device)), "pno_spectral_fwd")
        if needs_grad:
            xhat = ws[: 2 * 2 * m1 * m2 * b * ci].clone() if any(ctx.needs_input_grad[1:5]) else None
            ctx.save_for_backward(x, n1, n2, l1, l2, pre, xhat)
            ctx.meta = (plan, b, ci, co, h, w, m1, m2, seed, sample_idx, site, act, engine, sampled, addend is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, n1, n2, l1, l2, pre, xhat = ctx.saved_tensors
        plan, b, ci, co, h, w, m1, m2, seed, sample_idx, site, act, engine, sampled, has_add = ctx.meta
        g = g.contiguous()
        if act != _lib.ACT_NONE:
            gz = torch.empty_like(g)
            with _lib.on_device(g):
                _lib.check(lib.pno_act_bwd(pre.data_ptr(), g.data_ptr(), g.numel(), act, gz.data_ptr(),
                                           _lib.stream_ptr(g.device)), "pno_act_bwd")
        else:
            gz = g
        need_x, need_w, need_lv = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2], \
            sampled and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        dx = torch.empty_like(x) if need_x else None
        dev = x.device
        want_w = need_w or need_lv
        dmu1 = torch.empty(m1, m2, ci, co, dtype=torch.complex64, device=dev) if want_w else None
        dmu2 = torch.empty(m1, m2, ci, co, dtype=torch.complex64, device=dev) if want_w else None
        dlv1 = torch.empty(m1, m2, ci, co, dtype=torch.float32, device=dev) if need_lv else None
        dlv2 = torch.empty(m1, m2, ci, co, dtype=torch.float32, device=dev) if need_lv else None
        ws = torch.empty(lib.pno_spectral_workspace_bytes(plan, b, ci, co) // 4, dtype=torch.float32, device=dev)
        with _lib.on_device(x):
            _lib.check(lib.pno_spectral_bwd(plan, engine, x.data_ptr(), _lib.ptr(xhat), gz.data_ptr(), b, ci, co,
                                            n1.data_ptr(), n2.data_ptr(), _lib.ptr(l1), _lib.ptr(l2), seed, sample_idx, site,
                                            _lib.ptr(dx), _lib.ptr(dmu1), _lib.ptr(dmu2), _lib.ptr(dlv1), _lib.ptr(dlv2),
                                            ws.data_ptr(), _lib.stream_ptr(x.device)), "pno_spectral_bwd")

        def logical(t):
            return t.permute(2, 3, 0, 1) if t is not None else None

        return (dx, logical(dmu1) if ctx.needs_input_grad[1] else None, logical(dmu2) if ctx.needs_input_grad[2] else None,
                logical(dlv1) if ctx.needs_input_grad[3] else None, logical(dlv2) if ctx.needs_input_grad[4] else None,
                gz if (has_add and ctx.needs_input_grad[5]) else None, None, None, None, None, None, None)


def spectral_conv(x, w1, w2, lv1=None, lv2=None, *, addend=None, act=_lib.ACT_NONE, seed=0, sample_idx=0, site=0,
                  engine=_lib.ENGINE_FP32):
    """act(spectral(x; W) + addend) with W = mu (lv None) or mu + exp(lv/2)*eps(seed, sample_idx, site)."""
    return _SpectralConvFn.apply(x, w1, w2, lv1, lv2, addend, int(seed), int(sample_idx), int(site), int(act), int(engine),
                                  torch.is_grad_enabled())


# ----------------------------------------------------------------------------------------------------------------------
# local branch: 1x1 conv mean (+ log-var conv + reparameterised noise)
# ----------------------------------------------------------------------------------------------------------------------
class _LocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wm, bm, wl, bl, seed, sample_idx, site, batch_offset, engine, grad_mode):
        lib = _lib.load()
        _lib.require_cuda(x, "input")
        if x.dtype != torch.float32:
            # reference: nn.Conv2d with float32 weights raises RuntimeError for a half / double input
            raise RuntimeError(f"expected a float32 input, got {x.dtype}")
        if x.dim() < 3 or x.size(1) != wm.shape[1]:
            raise RuntimeError(f"expected an input with {wm.shape[1]} channels, got shape {tuple(x.shape)}")
        for name, t in (("weight", wm), ("bias", bm), ("log_var weight", wl), ("log_var bias", bl)):
            _lib.require_param(t, x, name)
        x = x.contiguous()
        b, ci = x.shape[0], x.shape[1]
        hw = x.numel() // (b * ci)
        co = wm.shape[0]
        noisy = wl is not None
        out = torch.empty((b, co) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
        needs_grad = grad_mode and any(ctx.needs_input_grad)   # under no_grad nothing is produced or kept for backward
        dlv = torch.empty_like(out) if (noisy and needs_grad) else None
        wm_c = wm.detach().reshape(co, ci).contiguous()
        wl_c = wl.detach().reshape(co, ci).contiguous() if noisy else None
        with _lib.on_device(x):
            _lib.check(lib.pno_local_fwd(engine, x.data_ptr(), b, ci, co, hw, wm_c.data_ptr(), _lib.ptr(bm),
                                         _lib.ptr(wl_c), _lib.ptr(bl) if noisy else 0, seed, sample_idx, site, batch_offset,
                                         out.data_ptr(), _lib.ptr(dlv), _lib.stream_ptr(x.device)), "pno_local_fwd")
        if needs_grad:
            ctx.save_for_backward(x, wm_c, wl_c, dlv)
            ctx.meta = (b, ci, co, hw, noisy, wm.shape, bm is not None, bl is not None, engine)
        return out

    @staticmethod
    def backward(ctx, g):
        """dx, [dWm; dWl] and the bias sums from pno_local_bwd: tensor-core kernels where the engine tiles the shape
        (csrc/pno_tc_local_bwd.cu), streaming kernels for lift- / projection-shaped layers, CUDA-core GEMM kernels for the fp32
        engine and fallback shapes (csrc/pno_local_simt.cu).  No torch / cuBLAS GEMM is involved."""
        x, wm, wl, dlv = ctx.saved_tensors
        b, ci, co, hw, noisy, wshape, has_bm, has_bl, engine = ctx.meta
        ni = ctx.needs_input_grad
        lib = _lib.load()
        gc = g.contiguous()
        rows = 2 * co if noisy else co
        dx = torch.empty_like(x) if ni[0] else None
        want_w = ni[1] or (noisy and ni[3])
        dw = torch.empty(rows, ci, dtype=torch.float32, device=x.device) if want_w else None
        want_b = (ni[2] and has_bm) or (noisy and ni[4] and has_bl)
        db = torch.empty(rows, dtype=torch.float32, device=x.device) if want_b else None
        ws = torch.empty(lib.pno_local_bwd_workspace_bytes(b, ci, co, hw, int(noisy)) // 4 + 4, dtype=torch.float32, device=x.device)
        with _lib.on_device(x):
            _lib.check(lib.pno_local_bwd(engine, x.data_ptr(), gc.data_ptr(), _lib.ptr(dlv), b, ci, co, hw, wm.data_ptr(),
                                         _lib.ptr(wl), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), ws.data_ptr(),
                                         _lib.stream_ptr(x.device)), "pno_local_bwd")
        return (dx, dw[:co].reshape(wshape) if ni[1] else None, db[:co] if (ni[2] and has_bm) else None,
                dw[co:].reshape(wshape) if (noisy and ni[3]) else None, db[co:] if (noisy and ni[4] and has_bl) else None,
                None, None, None, None, None, None)


def local_branch(x, wm, bm, wl=None, bl=None, *, seed=0, sample_idx=0, site=0, batch_offset=0, engine=_lib.ENGINE_FP32):
    """conv1x1(x; wm, bm) [+ eps * clamp(exp(0.5*clamp(conv1x1(x; wl, bl), -10, 10)), 1e-6)]  (models.py:264-275)."""
    return _LocalFn.apply(x, wm, bm, wl, bl, int(seed), int(sample_idx), int(site), int(batch_offset), int(engine),
                          torch.is_grad_enabled())


# ----------------------------------------------------------------------------------------------------------------------
# KL divergence of the spectral weights (models.py:100-104)
# ----------------------------------------------------------------------------------------------------------------------
class _KLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *params):
        """params = (mu_0, lv_0, mu_1, lv_1, ...) logical [Ci,Co,m1,m2] tensors; returns the summed KL (0-d fp32)."""
        lib = _lib.load()
        dev = params[0].device
        _lib.require_cuda(params[0], "parameters")
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        natives = []
        for k in range(0, len(params), 2):
            mu, _ = native_spectral(params[k].detach())
            lv, _ = native_spectral(params[k + 1].detach())
            natives += [mu, lv]
            with _lib.on_device(mu):
                _lib.check(lib.pno_kl_fwd(mu.data_ptr(), lv.data_ptr(), lv.numel(), acc.data_ptr(),
                                          _lib.stream_ptr(mu.device)), "pno_kl_fwd")
        ctx.save_for_backward(*natives)
        return acc.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        natives = ctx.saved_tensors
        gs = g.detach().to(torch.float32).contiguous()
        grads = []
        for k in range(0, len(natives), 2):
            mu, lv = natives[k], natives[k + 1]
            dmu, dlv = torch.empty_like(mu), torch.empty_like(lv)
            with _lib.on_device(mu):
                _lib.check(lib.pno_kl_bwd(mu.data_ptr(), lv.data_ptr(), lv.numel(), gs.data_ptr(), 0, dmu.data_ptr(),
                                          dlv.data_ptr(), _lib.stream_ptr(mu.device)), "pno_kl_bwd")
            grads += [dmu.permute(2, 3, 0, 1), dlv.permute(2, 3, 0, 1)]
        return tuple(grads)


def kl_divergence(*params):
    return _KLFn.apply(*params)


def kl_grad_accumulate_(params, gscale):
    """In-place form of `_KLFn.backward` for a training loop that owns the step: adds gscale * dKL/d(mu, log_var) straight into
    the existing `.grad` buffers (one kernel per block) instead of letting autograd materialise a second gradient per
    parameter and add the two (4 full-size adds per layer).  params = (mu_0, lv_0, ...); gscale: 0-d float32 CUDA tensor."""
    lib = _lib.load()
    gs = gscale.detach().to(torch.float32).contiguous()
    for k in range(0, len(params), 2):
        mu_p, lv_p = params[k], params[k + 1]
        mu, _ = native_spectral(mu_p.detach())
        lv, _ = native_spectral(lv_p.detach())
        grads, fresh = [], False
        for p in (mu_p, lv_p):
            if p.grad is None:
                p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
            gn, view = native_spectral(p.grad)
            if not view:                       # a gradient in another layout: re-home it in the parameter's layout
                p.grad = torch.empty_like(p, memory_format=torch.preserve_format).copy_(p.grad)
                gn, view = native_spectral(p.grad)
            grads.append(gn)
        with _lib.on_device(mu):
            _lib.check(lib.pno_kl_bwd(mu.data_ptr(), lv.data_ptr(), lv.numel(), gs.data_ptr(), 1, grads[0].data_ptr(),
                                      grads[1].data_ptr(), _lib.stream_ptr(mu.device)), "pno_kl_bwd")


# ----------------------------------------------------------------------------------------------------------------------
# MC moments and noise materialisation
# ----------------------------------------------------------------------------------------------------------------------
def mc_accumulate(y, s1, s2):
    lib = _lib.load()
    y = y.contiguous()
    with _lib.on_device(y):
        _lib.check(lib.pno_mc_accumulate(y.data_ptr(), y.numel(), s1.data_ptr(), s2.data_ptr(), _lib.stream_ptr(y.device)),
                   "pno_mc_accumulate")


def mc_finalize(s1, s2, num_samples):
    lib = _lib.load()
    mean = torch.empty(s1.shape, dtype=torch.float32, device=s1.device)
    std = torch.empty_like(mean)
    with _lib.on_device(s1):
        _lib.check(lib.pno_mc_finalize(s1.data_ptr(), s2.data_ptr(), s1.numel(), num_samples, mean.data_ptr(), std.data_ptr(),
                                       _lib.stream_ptr(s1.device)), "pno_mc_finalize")
    return mean, std


def materialize_activation_noise(seed, sample_idx, site, shape, batch_offset=0, device="cuda"):
    """eps the kernels draw for an activation-shaped site, as a tensor (parity tests inject it into the oracle)."""
    lib = _lib.load()
    out = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    per_item = out.numel() // shape[0] if shape[0] else 0
    with torch.cuda.device(out.device):
        _lib.check(lib.pno_philox_normal_activation(int(seed), sample_idx, site, batch_offset * per_item, out.numel(),
                                                    out.data_ptr(), _lib.stream_ptr()), "pno_philox_normal_activation")
    return out


def materialize_spectral_noise(seed, sample_idx, site, ci, co, m1, m2, device="cuda"):
    """(eps_re, eps_im), each [2,Ci,Co,m1,m2] (block 0 = weights1), as the kernels draw them."""
    lib = _lib.load()
    re = torch.empty(2, ci, co, m1, m2, dtype=torch.float32, device=device)
    im = torch.empty_like(re)
    with torch.cuda.device(re.device):
        _lib.check(lib.pno_philox_normal_spectral(int(seed), sample_idx, site, ci, co, m1, m2, re.data_ptr(), im.data_ptr(),
                                                  _lib.stream_ptr()), "pno_philox_normal_spectral")
    return re, im


def philox_bits(seed, sample_idx, site, q0, nq, device="cuda"):
    lib = _lib.load()
    out = torch.empty(nq, 4, dtype=torch.int32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(lib.pno_philox_bits(int(seed), sample_idx, site, q0, nq, out.data_ptr(), _lib.stream_ptr()),
                   "pno_philox_bits")
    return out
