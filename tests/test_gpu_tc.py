"""Tensor-core (tcgen05 / TMEM / TMA) engine tests on a B200: descriptor self-tests first, then the fused kernels
against the fp32 CUDA-core engine and the oracle."""
import itertools

import pytest
import torch

from conftest import rel_l2
from pno_physics_bench_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def eng(stage, engine, h, w, m1, m2, b, ci, co):
    """Engine code for a direct C-ABI call: STRICT (the call raises unless the tensor-core kernel runs) wherever the tiling covers
    the shape -- `pno_engine_supports` is the tilings' own validation code -- and the explicit CUDA-core fallback elsewhere, so
    that no test can pass by silently comparing the CUDA-core engine with itself."""
    if engine == "fp32":
        return _lib.ENGINES["fp32"]
    stages = (stage,) if isinstance(stage, str) else stage
    if all(_lib.engine_supports(st, engine, h, w, m1, m2, b, ci, co) for st in stages):
        return _lib.ENGINES[engine]
    return _lib.ENGINES[engine + "+fallback"]


def torch_local_reference(x, wm, bm, wl, bl, seed, sidx, site, up, batch_offset=0):
    """out and gradients of the local branch by torch autograd in FLOAT64 with the kernels' own noise (independent of every
    kernel here; fp64 because cuDNN / cuBLAS fp32 convolutions may run in TF32)."""
    from pno_physics_bench_b200 import functional as F_
    xs = [None if t is None else t.detach().double().requires_grad_(True) for t in (x, wm, bm, wl, bl)]

    def conv(inp, w, b):
        return torch.einsum("oi,bihw->bohw", w.reshape(w.shape[0], w.shape[1]), inp) + b.view(1, -1, 1, 1)
    out = conv(xs[0], xs[1], xs[2])
    if wl is not None:
        eps = F_.materialize_activation_noise(seed, sidx, site, out.shape, batch_offset).double()
        lv = torch.clamp(conv(xs[0], xs[3], xs[4]), -10, 10)
        out = out + eps * torch.exp(0.5 * lv).clamp(min=1e-6)
    out.backward(up.double())
    return out.detach(), [None if t is None else t.grad for t in xs]


def tf32_trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("flags,n,k", [(f, n, k) for f in range(16) for n, k in ((64, 64), (128, 32))]
                         + [(0b110000 | 0b1010, 96, 96), (0b010000, 32, 128), (0b100101, 256, 64),
                            (0b1000000, 64, 40), (0b1010000, 160, 40), (0b1000000, 48, 24), (0b1100000, 16, 8)])
def test_umma_selftest(flags, n, k):
    lib = _lib.load()
    gen = torch.Generator().manual_seed(flags * 1000 + n + k)
    a = tf32_trunc(torch.randn(128, k, generator=gen)).to(DEV)
    b = tf32_trunc(torch.randn(k, n, generator=gen)).to(DEV)
    a_src = a.t().contiguous() if flags & 1 else a
    b_src = b if flags & 2 else b.t().contiguous()
    d = torch.full((128, n), float("nan"), device=DEV)
    _lib.check(lib.pno_selftest_umma(flags, n, k, a.data_ptr(), b.data_ptr(), a_src.data_ptr(), b_src.data_ptr(),
                                     d.data_ptr(), _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    sign = (-1.0 if flags & 16 else 1.0) * (-1.0 if flags & 32 else 1.0)
    want = sign * (a.double() @ b.double())
    assert rel_l2(d, want) < 1e-6, f"flags={flags:06b}"


@pytest.mark.parametrize("flags,n,k", [(0, 64, 64), (1, 64, 64), (2, 64, 32), (3, 128, 64), (1, 128, 32), (0, 16, 96), (2, 48, 64),
                                       (3, 192, 64)])
def test_umma_selftest_mixed_fp32_parity_product(flags, n, k):
    """kind::tf32 on the RAW fp32 tiles + two kind::f16 (bf16) cross-term MMAs reproduce the fp32 product to ~1e-6
    (pno_sm100.cuh, "x3b"); operands are full-precision normals (NOT pre-truncated to tf32)."""
    lib = _lib.load()
    gen = torch.Generator().manual_seed(flags * 77 + n + k)
    a = torch.randn(128, k, generator=gen).to(DEV)
    b = torch.randn(k, n, generator=gen).to(DEV)
    d = torch.full((128, n), float("nan"), device=DEV)
    scratch = torch.empty(512 * k, dtype=torch.uint8, device=DEV)
    _lib.check(lib.pno_selftest_umma_mixed(flags, n, k, a.data_ptr(), b.data_ptr(), d.data_ptr(), scratch.data_ptr(),
                                           _lib.stream_ptr()), "selftest mixed")
    torch.cuda.synchronize()
    err = rel_l2(d, a.double() @ b.double())
    print(f"\n[x3b selftest flags={flags:03b} N={n} K={k}] rel-L2 vs fp64 product: {err:.2e}")
    assert err < 2e-6


def test_tf32_mma_truncates_raw_fp32_operands():
    """The main term of the x3b product relies on kind::tf32 IGNORING the 13 low mantissa bits of a raw fp32 operand
    (truncation, not rounding): lo = v - trunc(v) is then exactly what the main term missed."""
    lib = _lib.load()
    gen = torch.Generator().manual_seed(9)
    a = torch.randn(128, 64, generator=gen).to(DEV)
    b = torch.randn(64, 64, generator=gen).to(DEV)
    outs = []
    for aa, bb in ((a, b), (tf32_trunc(a.cpu()).to(DEV), tf32_trunc(b.cpu()).to(DEV))):
        d = torch.full((128, 64), float("nan"), device=DEV)
        _lib.check(lib.pno_selftest_umma_mixed(4, 64, 64, aa.data_ptr(), bb.data_ptr(), d.data_ptr(), None, _lib.stream_ptr()),
                   "selftest mixed")
        torch.cuda.synchronize()
        outs.append(d)
    assert torch.equal(outs[0], outs[1])


def _local_case(b, ci, co, h, w, noisy, engine, gen):
    from pno_physics_bench_b200 import functional as F_
    x = torch.randn(b, ci, h, w, generator=gen).to(DEV).requires_grad_(True)
    wm = (torch.randn(co, ci, 1, 1, generator=gen) / ci ** 0.5).to(DEV).requires_grad_(True)
    bm = torch.randn(co, generator=gen).to(DEV).requires_grad_(True)
    wl = (torch.randn(co, ci, 1, 1, generator=gen) * 0.3).to(DEV).requires_grad_(True) if noisy else None
    bl = (torch.randn(co, generator=gen) * 2 - 4).to(DEV).requires_grad_(True) if noisy else None
    out = F_.local_branch(x, wm, bm, wl, bl, seed=11, sample_idx=2, site=5, batch_offset=3,
                          engine=eng(("local_fwd", "local_bwd"), engine, h, w, 1, 1, b, ci, co))
    up = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).to(DEV)
    out.backward(up)
    grads = [t.grad.clone() for t in (x, wm, bm) + ((wl, bl) if noisy else ())]
    return out.detach(), grads


@pytest.mark.parametrize("engine,tol", [("tf32x3", 1e-5), ("tf32", 5e-3)])
@pytest.mark.parametrize("noisy", [False, True])
@pytest.mark.parametrize("shape", [(2, 32, 128, 16, 16), (3, 256, 256, 64, 64), (1, 64, 384, 8, 48), (5, 128, 256, 16, 32),
                                   (2, 256, 128, 8, 16)])
def test_local_branch_tensor_core_vs_cuda_core(engine, tol, noisy, shape):
    b, ci, co, h, w = shape
    ref_out, ref_g = _local_case(b, ci, co, h, w, noisy, "fp32", torch.Generator().manual_seed(sum(shape)))
    out, g = _local_case(b, ci, co, h, w, noisy, engine, torch.Generator().manual_seed(sum(shape)))
    assert rel_l2(out, ref_out) < tol
    # single-pass TF32 can flip clamp decisions of elements that sit on the +-10 boundary: loose bound on its gradients
    for a_, b_ in zip(g, ref_g):
        assert rel_l2(a_, b_) < (10 * tol if engine == "tf32x3" else 0.2)


@pytest.mark.parametrize("mode", ["1", "2", "3"])
@pytest.mark.parametrize("engine,tol", [("tf32x3", 1e-5), ("tf32", 5e-3)])
def test_local_branch_cluster_modes(monkeypatch, engine, tol, mode):
    """The three launch forms of the local-branch kernel agree with the CUDA-core engine: single CTAs (1), 2-CTA clusters that
    multicast the weight tiles (2), CTA pairs running 2-CTA MMAs (tcgen05 cta_group::2, M = 256; 3).  PNO_LOCAL_CL selects the
    form; several tiles per CTA (persistent loop), both 1x1 convs with noise."""
    shape = (3, 256, 256, 64, 64)
    ref_out, _ = _local_case(*shape, True, "fp32", torch.Generator().manual_seed(77))
    monkeypatch.setenv("PNO_LOCAL_CL", mode)
    out, _ = _local_case(*shape, True, engine, torch.Generator().manual_seed(77))
    assert rel_l2(out, ref_out) < tol


def test_full_size_inverse_every_tile_is_right(monkeypatch):
    """Single pass, full cfg-3 shape, repeated launches: stage 2 reads Z from tensor memory behind stage 1 with no barrier in
    between (in-order MMA pipe) and y leaves by TMA store from the staged addend buffers.  A protocol race would corrupt
    single 128-row tiles, which a whole-tensor rel-L2 could hide: check the worst element against the CUDA-core kernel, and
    the tensor-memory form against the transposer form."""
    b, c, h, w, m1, m2 = 64, 256, 64, 64, 20, 20
    gen = torch.Generator().manual_seed(99)
    m = 2 * m1 * m2
    yr = (torch.randn(m, b, c, generator=gen) / 8).to(DEV)
    yi = (torch.randn(m, b, c, generator=gen) / 8).to(DEV)
    add = torch.randn(b, c, h, w, generator=gen).to(DEV)
    ref, _ = _dft_inv(yr, yi, h, w, m1, m2, "fp32", 0, add, 1)
    scale = float(ref.abs().max())
    for rep in range(3):
        got, _ = _dft_inv(yr, yi, h, w, m1, m2, "tf32", 0, add, 1)
        assert not torch.isnan(got).any()
        assert float((got - ref).abs().max()) < 4e-3 * scale, rep
    monkeypatch.setenv("PNO_INV_TS", "0")
    old, _ = _dft_inv(yr, yi, h, w, m1, m2, "tf32", 0, add, 1)
    assert float((got - old).abs().max()) < 4e-3 * scale


@pytest.mark.parametrize("noisy", [False, True])
@pytest.mark.parametrize("shape", [(4, 3, 256, 32, 32), (2, 1, 128, 8, 12), (3, 4, 96, 16, 16)])
def test_lift_shaped_backward_vs_torch(noisy, shape):
    """Weight / bias gradients of a lift-shaped 1x1 conv pair (Ci <= 4, input is data): the streaming kernels of pno_local_fwd /
    pno_local_bwd against torch autograd with the same noise."""
    from pno_physics_bench_b200 import functional as F_
    b, ci, co, h, w = shape
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(b, ci, h, w, generator=gen).to(DEV)
    wm = (torch.randn(co, ci, 1, 1, generator=gen)).to(DEV).requires_grad_(True)
    bm = torch.randn(co, generator=gen).to(DEV).requires_grad_(True)
    wl = (torch.randn(co, ci, 1, 1, generator=gen) * 0.3).to(DEV).requires_grad_(True) if noisy else None
    bl = (torch.randn(co, generator=gen) - 3).to(DEV).requires_grad_(True) if noisy else None
    up = torch.randn(b, co, h, w, generator=torch.Generator().manual_seed(2)).to(DEV)
    out = F_.local_branch(x, wm, bm, wl, bl, seed=5, sample_idx=1, site=0, engine=_lib.ENGINES["tf32x3"])
    out.backward(up)
    ref_out, ref_g = torch_local_reference(x.requires_grad_(False), wm, bm, wl, bl, 5, 1, 0, up)
    assert rel_l2(out, ref_out) < 1e-5
    for t, g in zip((wm, bm) + ((wl, bl) if noisy else ()), ref_g[1:]):
        assert rel_l2(t.grad, g) < 2e-5, rel_l2(t.grad, g)


@pytest.mark.parametrize("noisy", [False, True])
@pytest.mark.parametrize("shape", [(4, 256, 1, 32, 32), (2, 64, 3, 8, 12), (3, 96, 4, 16, 16)])
def test_projection_shaped_backward_vs_torch(noisy, shape):
    """dx, weight and bias gradients of a projection-shaped 1x1 conv pair (Co <= 4): the streaming kernels of pno_local_fwd /
    pno_local_bwd against torch autograd with the same noise."""
    from pno_physics_bench_b200 import functional as F_
    b, ci, co, h, w = shape
    gen = torch.Generator().manual_seed(sum(shape) + 3)
    x = torch.randn(b, ci, h, w, generator=gen).to(DEV).requires_grad_(True)
    wm = (torch.randn(co, ci, 1, 1, generator=gen) / ci ** 0.5).to(DEV).requires_grad_(True)
    bm = torch.randn(co, generator=gen).to(DEV).requires_grad_(True)
    wl = (torch.randn(co, ci, 1, 1, generator=gen) * 0.1).to(DEV).requires_grad_(True) if noisy else None
    bl = (torch.randn(co, generator=gen) - 3).to(DEV).requires_grad_(True) if noisy else None
    up = torch.randn(b, co, h, w, generator=torch.Generator().manual_seed(2)).to(DEV)
    out = F_.local_branch(x, wm, bm, wl, bl, seed=5, sample_idx=1, site=13, engine=_lib.ENGINES["tf32x3"])
    out.backward(up)
    ref_out, ref_g = torch_local_reference(x, wm, bm, wl, bl, 5, 1, 13, up)
    assert rel_l2(out, ref_out) < 1e-5
    for t, g in zip((x, wm, bm) + ((wl, bl) if noisy else ()), ref_g):
        assert rel_l2(t.grad, g) < 2e-5, rel_l2(t.grad, g)


def _dft_fwd(x, m1, m2, engine, grad_scale=0):
    lib = _lib.load()
    b, c, h, w = x.shape
    plan = _lib.plan(h, w, m1, m2, x.device)
    m = 2 * m1 * m2
    xr = torch.full((m, b, c), float("nan"), device=x.device)
    xi = torch.full((m, b, c), float("nan"), device=x.device)
    _lib.check(lib.pno_dft_fwd(plan, eng("dft_fwd", engine, h, w, m1, m2, b, c, c), x.data_ptr(), b, c, grad_scale, xr.data_ptr(),
                               xi.data_ptr(), _lib.stream_ptr()), "pno_dft_fwd")
    torch.cuda.synchronize()
    return xr, xi


@pytest.mark.parametrize("engine,tol", [("tf32x3", 2e-6), ("tf32", 2e-3)])
@pytest.mark.parametrize("shape", [(2, 8, 64, 64, 20, 20), (1, 16, 64, 64, 12, 12), (3, 8, 128, 128, 32, 32),
                                   (2, 16, 32, 96, 8, 12), (64, 256, 64, 64, 20, 20),
                                   # two / one image per stage-2 batch (channel counts that only allow short super tiles): the
                                   # scalar-store branches of the ky-major epilogue
                                   (2, 6, 128, 128, 8, 8), (1, 3, 128, 64, 12, 12)])
def test_dft_fwd_tensor_core_vs_cuda_core(engine, tol, shape):
    b, c, h, w, m1, m2 = shape
    x = torch.randn(b, c, h, w, generator=torch.Generator().manual_seed(sum(shape))).to(DEV)
    for gs in (0, 1):
        rr, ri = _dft_fwd(x, m1, m2, "fp32", gs)
        tr, ti = _dft_fwd(x, m1, m2, engine, gs)
        assert rel_l2(tr, rr) < tol and rel_l2(ti, ri) < tol, (gs, rel_l2(tr, rr), rel_l2(ti, ri))
    # and against torch.fft directly (fp64)
    xf = torch.fft.rfft2(x.double())
    want = torch.cat([xf[:, :, :m1, :m2], xf[:, :, h - m1:, :m2]], dim=2)          # [b,c,2m1,m2]
    want = want.permute(2, 3, 0, 1).reshape(2 * m1 * m2, b, c)
    tr, ti = _dft_fwd(x, m1, m2, engine, 0)
    assert rel_l2(torch.complex(tr.double(), ti.double()), want) < max(tol, 5e-7)


def _dft_inv(yr, yi, h, w, m1, m2, engine, adjoint=0, addend=None, act=0, want_pre=False):
    lib = _lib.load()
    m, b, c = yr.shape
    plan = _lib.plan(h, w, m1, m2, yr.device)
    y = torch.full((b, c, h, w), float("nan"), device=yr.device)
    pre = torch.full((b, c, h, w), float("nan"), device=yr.device) if want_pre else None
    _lib.check(lib.pno_dft_inv(plan, eng("dft_inv", engine, h, w, m1, m2, b, c, c), yr.data_ptr(), yi.data_ptr(), b, c, adjoint,
                               _lib.ptr(addend), act, y.data_ptr(), _lib.ptr(pre), _lib.stream_ptr()), "pno_dft_inv")
    torch.cuda.synchronize()
    return y, pre


@pytest.mark.parametrize("engine,tol", [("tf32x3", 2e-6), ("tf32", 2e-3)])
@pytest.mark.parametrize("shape", [(2, 8, 64, 64, 20, 20), (1, 16, 64, 64, 12, 12), (3, 8, 128, 128, 32, 32),
                                   (2, 16, 32, 96, 8, 12), (1, 4, 64, 48, 36, 8), (64, 256, 64, 64, 20, 20)])
def test_dft_inv_tensor_core_vs_cuda_core(engine, tol, shape):
    b, c, h, w, m1, m2 = shape
    gen = torch.Generator().manual_seed(sum(shape))
    m = 2 * m1 * m2
    yr = torch.randn(m, b, c, generator=gen).to(DEV)
    yi = torch.randn(m, b, c, generator=gen).to(DEV)
    add = torch.randn(b, c, h, w, generator=gen).to(DEV)
    for adjoint, addend, act in ((0, None, 0), (1, None, 0), (0, add, 1), (0, add, 2)):
        ref, ref_pre = _dft_inv(yr, yi, h, w, m1, m2, "fp32", adjoint, addend, act, True)
        out, pre = _dft_inv(yr, yi, h, w, m1, m2, engine, adjoint, addend, act, True)
        assert rel_l2(out, ref) < tol and rel_l2(pre, ref_pre) < tol, (adjoint, act, rel_l2(out, ref))


def _mix(xr, xi, layer, sampled, engine):
    lib = _lib.load()
    m, b, ci = xr.shape
    co = layer.out_channels
    plan = _lib.plan(64, 64, layer.modes1, layer.modes2, xr.device)
    n = [t.detach().permute(2, 3, 0, 1) for t in (layer.weights1, layer.weights2, layer.weights1_log_var, layer.weights2_log_var)]
    assert all(t.is_contiguous() for t in n)
    yr = torch.full((m, b, co), float("nan"), device=xr.device)
    yi = torch.full((m, b, co), float("nan"), device=xr.device)
    _lib.check(lib.pno_mix_fwd(plan, eng("mix_fwd", engine, 64, 64, layer.modes1, layer.modes2, b, ci, co), xr.data_ptr(), xi.data_ptr(), b, ci, co, n[0].data_ptr(),
                               n[1].data_ptr(), n[2].data_ptr() if sampled else 0, n[3].data_ptr() if sampled else 0,
                               777, 3, 9, yr.data_ptr(), yi.data_ptr(), _lib.stream_ptr()), "pno_mix_fwd")
    torch.cuda.synchronize()
    return yr, yi


@pytest.mark.parametrize("engine,tol", [("tf32x3", 5e-6), ("tf32", 2e-3)])   # fp32 CUDA-core sums of 512 terms are ~2e-6 themselves
@pytest.mark.parametrize("sampled", [False, True])
@pytest.mark.parametrize("shape", [(16, 128, 128, 4, 4), (64, 256, 256, 20, 20), (32, 64, 256, 5, 3), (128, 128, 128, 2, 2),
                                   (2, 128, 128, 3, 4), (20, 64, 128, 2, 2)])       # batch smaller than / not a multiple of the MMA N step
def test_mix_tensor_core_vs_cuda_core(engine, tol, sampled, shape):
    import pno_physics_bench_b200 as pno
    b, ci, co, m1, m2 = shape
    gen = torch.Generator().manual_seed(sum(shape))
    layer = pno.SpectralConv2d_Probabilistic(ci, co, m1, m2)
    with torch.no_grad():
        for w_, lv in ((layer.weights1, layer.weights1_log_var), (layer.weights2, layer.weights2_log_var)):
            w_.copy_(torch.complex(torch.randn(w_.shape, generator=gen), torch.randn(w_.shape, generator=gen)) / ci ** 0.5)
            lv.copy_(torch.rand(lv.shape, generator=gen) * 4 - 6)
    layer = layer.to(DEV)
    m = 2 * m1 * m2
    xr = torch.randn(m, b, ci, generator=gen).to(DEV)
    xi = torch.randn(m, b, ci, generator=gen).to(DEV)
    rr, ri = _mix(xr, xi, layer, sampled, "fp32")
    tr, ti = _mix(xr, xi, layer, sampled, engine)
    assert rel_l2(tr, rr) < tol and rel_l2(ti, ri) < tol, (rel_l2(tr, rr), rel_l2(ti, ri))
    # a pipeline race (TMA-staged weight tiles, operand rings) would corrupt single (mode, 128-channel) tiles that the
    # whole-tensor norm can hide: bound the worst element too, over repeated launches
    scale = float(rr.abs().max())
    for _ in range(2):
        tr, ti = _mix(xr, xi, layer, sampled, engine)
        assert float((tr - rr).abs().max()) < 20 * tol * scale and float((ti - ri).abs().max()) < 20 * tol * scale


@pytest.mark.parametrize("engine,tol", [("tf32x3", 1e-5), ("tf32", 2e-2)])
def test_model_tensor_core_engines_against_oracle(engine, tol):
    """A model whose every stage is tiled by the tensor-core kernels (hid 128, 32x32, modes 8, batch 16): forward + backward."""
    import pno_physics_bench_b200 as pno
    from pno_physics_bench_b200 import functional as F_
    from oracle import pno_oracle as orc
    gen = torch.Generator().manual_seed(17)
    torch.manual_seed(17)
    hid, nl, modes, b, hw = 128, 2, 8, 16, 32
    model = pno.ProbabilisticNeuralOperator(3, hid, nl, modes, engine=engine)
    with torch.no_grad():
        for layer in model.fourier_layers:
            for w_, lv in ((layer.weights1, layer.weights1_log_var), (layer.weights2, layer.weights2_log_var)):
                w_.copy_(torch.complex(torch.randn(w_.shape, generator=gen), torch.randn(w_.shape, generator=gen)) / hid ** 0.5)
                lv.copy_(torch.rand(lv.shape, generator=gen) * 4 - 6)
        for conv in [model.lift_log_var, model.project_log_var, *model.linear_log_var]:
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen) * 0.05)
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=gen) - 3.0)
    model = model.to(DEV).train()
    x = torch.randn(b, 3, hw, hw, generator=gen)
    tgt = torch.randn(b, 1, hw, hw, generator=gen)
    seed, sidx = 99, 4
    with pno.noise_scope(seed, sidx):
        y = model(x.to(DEV), sample=True)
    loss = torch.mean((y - tgt.to(DEV)) ** 2) + 1e-4 * model.kl_divergence()
    loss.backward()
    table = {0: F_.materialize_activation_noise(seed, sidx, 0, (b, hid, hw, hw)).cpu(),
             1 + 3 * nl: F_.materialize_activation_noise(seed, sidx, 1 + 3 * nl, (b, 1, hw, hw)).cpu()}
    for l in range(nl):
        re, im = F_.materialize_spectral_noise(seed, sidx, 1 + 3 * l, hid, hid, modes, modes)
        table[1 + 3 * l] = (re.cpu(), im.cpu())
        table[2 + 3 * l] = F_.materialize_activation_noise(seed, sidx, 2 + 3 * l, (b, hid, hw, hw)).cpu()
    p = {k: v.detach().cpu().contiguous().requires_grad_(True) for k, v in model.state_dict().items()}
    yo = orc.pno_forward(p, x, orc.InjectedNoise(table), True, True, "gelu")
    lo = torch.mean((yo - tgt) ** 2) + 1e-4 * orc.pno_kl(p)
    lo.backward()
    assert rel_l2(y, yo) < tol and rel_l2(loss, lo) < tol
    for k, v in model.named_parameters():
        assert rel_l2(v.grad, p[k].grad) < 5 * tol, k


def _mix_bwd_dx(gr, gi, layer, sampled, engine):
    lib = _lib.load()
    m, b, co = gr.shape
    ci = layer.in_channels
    plan = _lib.plan(64, 64, layer.modes1, layer.modes2, gr.device)
    n = [t.detach().permute(2, 3, 0, 1) for t in (layer.weights1, layer.weights2, layer.weights1_log_var, layer.weights2_log_var)]
    dxr = torch.full((m, b, ci), float("nan"), device=gr.device)
    dxi = torch.full((m, b, ci), float("nan"), device=gr.device)
    _lib.check(lib.pno_mix_bwd_dx(plan, eng("mix_bwd_dx", engine, 64, 64, layer.modes1, layer.modes2, b, ci, co), gr.data_ptr(), gi.data_ptr(), b, ci, co, n[0].data_ptr(),
                                  n[1].data_ptr(), n[2].data_ptr() if sampled else 0, n[3].data_ptr() if sampled else 0,
                                  777, 3, 9, dxr.data_ptr(), dxi.data_ptr(), _lib.stream_ptr()), "pno_mix_bwd_dx")
    torch.cuda.synchronize()
    return dxr, dxi


@pytest.mark.parametrize("engine,tol", [("tf32x3", 5e-6), ("tf32", 2e-3)])
@pytest.mark.parametrize("sampled", [False, True])
@pytest.mark.parametrize("shape", [(16, 128, 128, 4, 4), (64, 256, 256, 20, 20), (20, 128, 64, 2, 2), (3, 256, 32, 3, 4)])
def test_mix_backward_dx_tensor_core_vs_cuda_core(engine, tol, sampled, shape):
    """Input gradient of the channel mix (W re-sampled in-kernel, conjugated): tensor-core engines vs the fp32 CUDA-core kernel,
    which the layer-level tests pin to the oracle's autograd."""
    import pno_physics_bench_b200 as pno
    b, ci, co, m1, m2 = shape
    gen = torch.Generator().manual_seed(sum(shape) + 1)
    layer = pno.SpectralConv2d_Probabilistic(ci, co, m1, m2)
    with torch.no_grad():
        for w_, lv in ((layer.weights1, layer.weights1_log_var), (layer.weights2, layer.weights2_log_var)):
            w_.copy_(torch.complex(torch.randn(w_.shape, generator=gen), torch.randn(w_.shape, generator=gen)) / co ** 0.5)
            lv.copy_(torch.rand(lv.shape, generator=gen) * 4 - 6)
    layer = layer.to(DEV)
    m = 2 * m1 * m2
    gr = torch.randn(m, b, co, generator=gen).to(DEV)
    gi = torch.randn(m, b, co, generator=gen).to(DEV)
    rr, ri = _mix_bwd_dx(gr, gi, layer, sampled, "fp32")
    tr, ti = _mix_bwd_dx(gr, gi, layer, sampled, engine)
    assert rel_l2(tr, rr) < tol and rel_l2(ti, ri) < tol, (rel_l2(tr, rr), rel_l2(ti, ri))


def _mix_bwd_dw(xr, xi, gr, gi, layer, sampled, engine):
    lib = _lib.load()
    m, b, ci = xr.shape
    co = gr.shape[2]
    m1, m2 = layer.modes1, layer.modes2
    plan = _lib.plan(64, 64, m1, m2, gr.device)
    lv = [t.detach().permute(2, 3, 0, 1) for t in (layer.weights1_log_var, layer.weights2_log_var)]
    dmu = [torch.full((m1, m2, ci, co, 2), float("nan"), device=gr.device) for _ in range(2)]
    dlv = [torch.full((m1, m2, ci, co), float("nan"), device=gr.device) for _ in range(2)]
    _lib.check(lib.pno_mix_bwd_dw(plan, eng("mix_bwd_dw", engine, 64, 64, m1, m2, b, ci, co), xr.data_ptr(), xi.data_ptr(), gr.data_ptr(), gi.data_ptr(), b, ci, co,
                                  lv[0].data_ptr() if sampled else 0, lv[1].data_ptr() if sampled else 0, 777, 3, 9,
                                  dmu[0].data_ptr(), dmu[1].data_ptr(), dlv[0].data_ptr() if sampled else 0,
                                  dlv[1].data_ptr() if sampled else 0, _lib.stream_ptr()), "pno_mix_bwd_dw")
    torch.cuda.synchronize()
    return dmu + (dlv if sampled else [])


@pytest.mark.parametrize("engine,tol", [("tf32x3", 5e-6), ("tf32", 2e-3)])
@pytest.mark.parametrize("sampled", [False, True])
@pytest.mark.parametrize("shape", [(32, 256, 256, 6, 5), (5, 128, 256, 3, 4), (64, 128, 128, 2, 2), (12, 256, 128, 4, 4),
                                   (72, 128, 128, 2, 3), (200, 128, 128, 1, 2)])      # the last two: several K chunks of the batch
def test_mix_backward_dw_tensor_core_vs_cuda_core(engine, tol, sampled, shape):
    """Weight gradient of the channel mix (dmu and, with the noise re-drawn, dlv): tensor-core kernel vs the fp32 CUDA-core
    kernel, which the layer-level tests pin to the oracle's autograd.  Ragged batches exercise the TMA zero fill."""
    import pno_physics_bench_b200 as pno
    b, ci, co, m1, m2 = shape
    gen = torch.Generator().manual_seed(sum(shape) + 2)
    layer = pno.SpectralConv2d_Probabilistic(ci, co, m1, m2)
    with torch.no_grad():
        for lv in (layer.weights1_log_var, layer.weights2_log_var):
            lv.copy_(torch.rand(lv.shape, generator=gen) * 4 - 6)
    layer = layer.to(DEV)
    m = 2 * m1 * m2
    xr, xi = torch.randn(m, b, ci, generator=gen).to(DEV), torch.randn(m, b, ci, generator=gen).to(DEV)
    gr, gi = torch.randn(m, b, co, generator=gen).to(DEV), torch.randn(m, b, co, generator=gen).to(DEV)
    ref = _mix_bwd_dw(xr, xi, gr, gi, layer, sampled, "fp32")
    got = _mix_bwd_dw(xr, xi, gr, gi, layer, sampled, engine)
    for a_, b_ in zip(got, ref):
        assert not torch.isnan(a_).any()
        assert rel_l2(a_, b_) < tol, rel_l2(a_, b_)


# ----------------------------------------------------------------------------------------------------------------------
# size-independent properties at the full BASELINE config-3 layer shape (B 64, C 256, 64x64, modes 20), where the oracle is
# too slow to be the checker
# ----------------------------------------------------------------------------------------------------------------------
FULL = dict(B=64, C=256, H=64, W=64, m=20)


@pytest.mark.parametrize("engine,tol", [("tf32x3", 1e-5), ("tf32", 5e-3)])
def test_full_size_transform_pair(engine, tol):
    """forward + inverse transform at the full shape: the tensor-core engines against the fp32 CUDA-core engine (itself
    checked against the oracle at small sizes), and the pair reproduces a field made of conjugate-complete retained modes."""
    lib = _lib.load()
    B, C, H, W, m = (FULL[k] for k in ("B", "C", "H", "W", "m"))
    M = 2 * m * m
    plan = _lib.plan(H, W, m, m, torch.device(DEV, 0))
    eng = _lib.ENGINES[engine]
    gen = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(B, C, H, W, device=DEV, generator=gen)
    st = _lib.stream_ptr()

    def pair(e, xin):
        xr, xi = torch.empty(M, B, C, device=DEV), torch.empty(M, B, C, device=DEV)
        out = torch.empty_like(xin)
        _lib.check(lib.pno_dft_fwd(plan, e, xin.data_ptr(), B, C, 0, xr.data_ptr(), xi.data_ptr(), st))
        _lib.check(lib.pno_dft_inv(plan, e, xr.data_ptr(), xi.data_ptr(), B, C, 0, 0, _lib.ACT_NONE, out.data_ptr(), 0, st))
        return out, xr, xi

    y_ref, xr_ref, xi_ref = pair(_lib.ENGINE_FP32, x)
    y, xr, xi = pair(eng, x)
    assert rel_l2(xr, xr_ref) < tol and rel_l2(xi, xi_ref) < tol
    assert rel_l2(y, y_ref) < tol
    # a field built from retained modes whose conjugate partners are retained too (|kx| < m1, ky < m2) is a fixed point
    hh = torch.arange(H, device=DEV, dtype=torch.float32).view(H, 1)
    ww = torch.arange(W, device=DEV, dtype=torch.float32).view(1, W)
    f = torch.cos(2 * torch.pi * (3 * hh / H + 5 * ww / W)) + 0.5 * torch.sin(2 * torch.pi * (-7 * hh / H + 11 * ww / W)) + 0.25
    xf = f.expand(B, C, H, W).contiguous()
    yf, _, _ = pair(eng, xf)
    assert rel_l2(yf, xf) < tol


@pytest.mark.parametrize("engine,tol", [("tf32x3", 1e-5), ("tf32", 5e-3)])
def test_full_size_layer_is_linear_in_x_for_fixed_noise(engine, tol):
    """For one weight sample (fixed Philox key) the sampled layer is linear: L(a x1 + b x2) = a L(x1) + b L(x2)."""
    import pno_physics_bench_b200 as pno
    B, C, H, W, m = (FULL[k] for k in ("B", "C", "H", "W", "m"))
    torch.manual_seed(2)
    layer = pno.SpectralConv2d_Probabilistic(C, C, m, m, engine=engine).to(DEV).train()
    gen = torch.Generator(device=DEV).manual_seed(5)
    x1 = torch.randn(B, C, H, W, device=DEV, generator=gen)
    x2 = torch.randn(B, C, H, W, device=DEV, generator=gen)
    with torch.no_grad():
        outs = []
        for xin in (x1, x2, 0.75 * x1 - 1.5 * x2):
            with pno.noise_scope(77, 3):
                outs.append(layer(xin, sample=True))
    assert rel_l2(outs[2], 0.75 * outs[0] - 1.5 * outs[1]) < tol
    # and the draw really is a sample: another key gives another output
    with torch.no_grad(), pno.noise_scope(77, 4):
        other = layer(x1, sample=True)
    assert rel_l2(other, outs[0]) > 0.1


def test_full_size_engines_agree_on_one_block():
    """One PNO block at the headline shape: the single-pass engine stays within the reduced-precision budget of the
    fp32-parity engine (same noise), and the MC moments are finite."""
    import pno_physics_bench_b200 as pno
    B, C, H, W, m = (FULL[k] for k in ("B", "C", "H", "W", "m"))
    torch.manual_seed(4)
    blk = pno.PNOBlock(C, C, m, m, activation="gelu", local_noise="activation", engine="tf32x3").to(DEV).train()
    with torch.no_grad():
        blk.w_log_var.weight.mul_(0.02)
        blk.w_log_var.bias.fill_(-2.0)
    x = torch.randn(B, C, H, W, device=DEV)
    with torch.no_grad():
        with pno.noise_scope(9, 1):
            ref = blk(x, sample=True)
        blk.engine = _lib.ENGINES["tf32"]
        blk.conv.engine = blk.engine
        with pno.noise_scope(9, 1):
            fast = blk(x, sample=True)
    assert torch.isfinite(ref).all() and rel_l2(fast, ref) < 2e-2


def test_local_branch_wide_tile_variant(monkeypatch):
    """The optional 256-pixel-tile variant of the local-branch kernel (PNO_LOCAL_TN=256) computes the same thing."""
    b, ci, co, h, w = 2, 64, 128, 16, 32
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(b, ci, h, w, generator=gen).to(DEV)
    wm = (torch.randn(co, ci, generator=gen) / ci ** 0.5).to(DEV)
    wl = (torch.randn(co, ci, generator=gen) * 0.05).to(DEV)
    bm, bl = torch.randn(co, generator=gen).to(DEV), (torch.randn(co, generator=gen) - 3).to(DEV)
    lib = _lib.load()
    outs = []
    for tn in ("128", "256"):
        monkeypatch.setenv("PNO_LOCAL_TN", tn)
        out = torch.empty(b, co, h, w, device=DEV)
        _lib.check(lib.pno_local_fwd(_lib.ENGINES["tf32"], x.data_ptr(), b, ci, co, h * w, wm.data_ptr(), bm.data_ptr(), wl.data_ptr(),
                                     bl.data_ptr(), 5, 1, 2, 0, out.data_ptr(), 0, _lib.stream_ptr()))
        outs.append(out)
    ref = torch.empty(b, co, h, w, device=DEV)
    _lib.check(lib.pno_local_fwd(_lib.ENGINE_FP32, x.data_ptr(), b, ci, co, h * w, wm.data_ptr(), bm.data_ptr(), wl.data_ptr(),
                                 bl.data_ptr(), 5, 1, 2, 0, ref.data_ptr(), 0, _lib.stream_ptr()))
    assert rel_l2(outs[0], ref) < 5e-3 and rel_l2(outs[1], ref) < 5e-3
    assert rel_l2(outs[1], outs[0]) < 1e-6


# ----------------------------------------------------------------------------------------------------------------------
# two-pass tensor-core transforms (grids beyond the fused kernels: 256-row images; fp32-parity engine at 128x128 / 32 modes)
# ----------------------------------------------------------------------------------------------------------------------
def _fft_reference(x, m1, m2):
    b, c, h, w = x.shape
    xf = torch.fft.rfft2(x.double())
    want = torch.cat([xf[:, :, :m1, :m2], xf[:, :, h - m1:, :m2]], dim=2)          # [b,c,2m1,m2]
    return want.permute(2, 3, 0, 1).reshape(2 * m1 * m2, b, c)


def _ifft_reference(yr, yi, h, w, m1, m2):
    m, b, c = yr.shape
    yh = torch.complex(yr.double(), yi.double()).reshape(2 * m1, m2, b, c).permute(2, 3, 0, 1)
    full = torch.zeros(b, c, h, w // 2 + 1, dtype=torch.complex128, device=yr.device)
    full[:, :, :m1, :m2] = yh[:, :, :m1]
    full[:, :, h - m1:, :m2] = yh[:, :, m1:]
    return torch.fft.irfft2(full, s=(h, w))


@pytest.mark.parametrize("engine,tol", [("tf32x3", 2e-6), ("tf32", 2e-3)])
@pytest.mark.parametrize("shape,forced", [((2, 16, 256, 256, 48, 48), False), ((1, 8, 256, 128, 20, 33), False),
                                          ((3, 8, 128, 128, 32, 32), True), ((2, 4, 128, 64, 17, 20), True),
                                          ((1, 5, 128, 96, 64, 24), True)])
def test_two_pass_transforms_against_fft(engine, tol, shape, forced, monkeypatch):
    """Forward and inverse two-pass transforms (k_tc_rowgemm + k_tc_local<EPI=1> + the spectrum transpositions) against torch.fft in
    fp64; `forced` shapes are also tiled by the fused kernels and take the two-pass path through PNO_DFT_TWOPASS=1."""
    b, c, h, w, m1, m2 = shape
    if forced:
        monkeypatch.setenv("PNO_DFT_TWOPASS", "1")
    assert _lib.engine_supports("dft_fwd", engine, h, w, m1, m2, b, c, c) and _lib.engine_supports("dft_inv", engine, h, w, m1, m2, b, c, c)
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(b, c, h, w, generator=gen).to(DEV)
    want = _fft_reference(x, m1, m2)
    for gs in (0, 1):
        tr, ti = _dft_fwd(x, m1, m2, engine, gs)
        if gs == 0:       # 256-point fp32 sums (two passes, fp32 intermediate) measure 2.6e-6 .. 3.5e-6 in the fp32-parity engine
            assert rel_l2(torch.complex(tr.double(), ti.double()), want) < max(tol, 5e-6 if h > 128 else 5e-7)
        else:       # adjoint scaling = alive * c_ky / (HW): compare with the CUDA-core engine
            rr, ri = _dft_fwd(x, m1, m2, "fp32", gs)
            lim = max(tol, 5e-6 if h > 128 else 2e-6)
            assert rel_l2(tr, rr) < lim and rel_l2(ti, ri) < lim
    m = 2 * m1 * m2
    yr = torch.randn(m, b, c, generator=gen).to(DEV)
    yi = torch.randn(m, b, c, generator=gen).to(DEV)
    add = torch.randn(b, c, h, w, generator=gen).to(DEV)
    if 2 * m1 <= h:                                     # no overlapping rows: irfft2 of the scattered spectrum is the reference
        out, _ = _dft_inv(yr, yi, h, w, m1, m2, engine, 0, None, 0)
        # irfft2 ignores the imaginary part of the self-conjugate columns exactly as the kernels do
        assert rel_l2(out, _ifft_reference(yr, yi, h, w, m1, m2)) < max(tol, 5e-7)
    for adjoint, addend, act in ((1, None, 0), (0, add, 1), (0, add, 2)):
        ref, ref_pre = _dft_inv(yr, yi, h, w, m1, m2, "fp32", adjoint, addend, act, True)
        out, pre = _dft_inv(yr, yi, h, w, m1, m2, engine, adjoint, addend, act, True)
        lim = max(tol, 5e-6 if h > 128 else 2e-6)
        assert rel_l2(out, ref) < lim and rel_l2(pre, ref_pre) < lim, (adjoint, act, rel_l2(out, ref))


@pytest.mark.parametrize("engine,tol", [("tf32x3", 1e-5), ("tf32", 2e-2)])
def test_cfg5_shaped_layer_against_oracle(engine, tol):
    """A sampled spectral layer at the BASELINE config-5 grid (256 x 256, 48 modes; 128 channels, batch 2) against the oracle:
    forward and every gradient, all stages on the tensor cores (strict engine)."""
    import pno_physics_bench_b200 as pno
    from pno_physics_bench_b200 import functional as F_
    from oracle import pno_oracle as orc
    B, C, H, m = 2, 128, 256, 48
    gen = torch.Generator().manual_seed(55)
    torch.manual_seed(55)
    layer = pno.SpectralConv2d_Probabilistic(C, C, m, m, engine=engine)
    with torch.no_grad():
        for w_, lv in ((layer.weights1, layer.weights1_log_var), (layer.weights2, layer.weights2_log_var)):
            w_.copy_(torch.complex(torch.randn(w_.shape, generator=gen), torch.randn(w_.shape, generator=gen)) / (2 * C) ** 0.5)
            lv.copy_(torch.rand(lv.shape, generator=gen) * 4 - 10)
    layer.noise_site = 7
    x = torch.randn(B, C, H, H, generator=gen)
    up = torch.randn(B, C, H, H, generator=gen)
    p = {k: v.detach().clone().contiguous().requires_grad_(True) for k, v in layer.state_dict().items()}
    layer = layer.to(DEV).train()
    xg = x.to(DEV).requires_grad_(True)
    c0 = _lib.stage_counts()
    with pno.noise_scope(17, 2):
        y = layer(xg, sample=True)
    y.backward(up.to(DEV))
    counts = _lib.tensor_core_calls(c0)
    for st in ("dft_fwd", "mix_fwd", "dft_inv", "mix_bwd_dx", "mix_bwd_dw"):
        assert counts[st][0] == 0 and counts[st][1] > 0, (st, counts[st])
    re, im = F_.materialize_spectral_noise(17, 2, 7, C, C, m, m)
    xo = x.clone().requires_grad_(True)
    yo = orc.spectral_layer(xo, p, "", orc.InjectedNoise({7: (re.cpu(), im.cpu())}), site=7)
    yo.backward(up)
    errs = {"y": rel_l2(y, yo), "dx": rel_l2(xg.grad, xo.grad)}
    for k, v in layer.named_parameters():
        errs["d" + k] = rel_l2(v.grad, p[k].grad)
    print(f"\n[cfg5_layer | {engine}] " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < (tol if engine == "tf32" else 2 * tol), (k, v)
