// extern "C" surface of libpno_b200.so (see include/pno_b200.h): plans, argument validation, engine dispatch and the
// whole-layer compositions.
#include <stdarg.h>

#include <vector>

#include "pno_common.cuh"
#include "pno_tc.cuh"

static thread_local char g_err[512] = "";

void pno_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

unsigned long long g_pno_launches = 0;
int g_pno_sm_reserve = 0;
long long* g_pno_debug_buffer = nullptr;
thread_local const uint32_t* g_pno_dyn_key = nullptr;

// which engine ran each stage: g_stage_counts[stage][0 = CUDA cores, 1 = tensor cores] (pno_stage_counts)
static unsigned long long g_stage_counts[PNO_N_STAGES][2];

extern "C" int pno_stage_counts(uint64_t* out) {
    if (!out) PNO_FAIL(PNO_E_INVALID, "pno_stage_counts: null output");
    for (int s = 0; s < PNO_N_STAGES; ++s) {
        out[2 * s] = g_stage_counts[s][0];
        out[2 * s + 1] = g_stage_counts[s][1];
    }
    return PNO_OK;
}

thread_local int g_pno_dry_run = 0;

struct ScratchEntry { int dev, slot; cudaStream_t st; float* buf; size_t bytes; };
static ScratchEntry g_scratch[1024];
static int g_nscratch = 0;
float* pno_scratch(int slot, size_t bytes, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int i = 0; i < g_nscratch; ++i) {
        ScratchEntry& e = g_scratch[i];
        if (e.dev == dev && e.st == st && e.slot == slot) {
            if (e.bytes < bytes) {
                float* bigger = nullptr;
                if (cudaMalloc(&bigger, bytes) != cudaSuccess) return nullptr;
                e.buf = bigger;          // the outgrown buffer stays allocated (see pno_common.cuh)
                e.bytes = bytes;
            }
            return e.buf;
        }
    }
    if (g_nscratch == 1024) return nullptr;
    ScratchEntry& e = g_scratch[g_nscratch];
    e.dev = dev; e.slot = slot; e.st = st; e.bytes = bytes;
    if (cudaMalloc(&e.buf, bytes) != cudaSuccess) return nullptr;
    ++g_nscratch;
    return e.buf;
}

// Host-only (no CUDA call): would `engine` run `stage` on its tensor-core kernel at this shape?  The tilings' own validation code
// answers (dry run), so this can never disagree with what a launch would do.
extern "C" int pno_engine_supports(int stage, int engine, int H, int W, int m1, int m2, int B, int Ci, int Co) {
    const int eng = engine & 0xFF;
    if (eng == PNO_ENGINE_FP32) return 0;
    if (eng != PNO_ENGINE_TC_TF32X3 && eng != PNO_ENGINE_TC_TF32) return 0;
    if (H <= 0 || W <= 0 || m1 <= 0 || m2 <= 0 || B <= 0 || Ci <= 0 || Co <= 0 || m1 > H || m2 > W / 2 + 1) return 0;
    pno_plan plan;
    memset(&plan, 0, sizeof(plan));
    plan.H = H; plan.W = W; plan.m1 = m1; plan.m2 = m2; plan.M = 2 * m1 * m2;
    float* const d = reinterpret_cast<float*>(uintptr_t(4096));      // aligned stand-in for every device pointer
    const PhiloxKey key = make_key(0, 0, 0);
    g_pno_dry_run = 1;
    bool ok = true;
    switch (stage) {
    case PNO_STAGE_DFT_FWD:
        ok = tc_dft_fwd(&plan, eng, d, B, Ci, 0, d, d, nullptr) == PNO_OK;
        break;
    case PNO_STAGE_DFT_INV:      // bare layer, block epilogue (addend), block epilogue kept for backward (addend + pre-activation)
        ok = tc_dft_inv(&plan, eng, d, d, B, Co, 0, nullptr, PNO_ACT_NONE, d, nullptr, nullptr) == PNO_OK &&
             tc_dft_inv(&plan, eng, d, d, B, Co, 0, d, PNO_ACT_GELU, d, nullptr, nullptr) == PNO_OK &&
             tc_dft_inv(&plan, eng, d, d, B, Co, 0, d, PNO_ACT_GELU, d, d, nullptr) == PNO_OK;
        break;
    case PNO_STAGE_MIX_FWD:      // mean weights and sampled weights
        ok = tc_mix_fwd(&plan, eng, d, d, B, Ci, Co, d, d, nullptr, nullptr, key, d, d, nullptr) == PNO_OK &&
             tc_mix_fwd(&plan, eng, d, d, B, Ci, Co, d, d, d, d, key, d, d, nullptr) == PNO_OK;
        break;
    case PNO_STAGE_MIX_BWD_DX:
        ok = tc_mix_bwd_dx(&plan, eng, d, d, B, Ci, Co, d, d, d, d, key, d, d, nullptr) == PNO_OK;
        break;
    case PNO_STAGE_MIX_BWD_DW:
        ok = tc_mix_bwd_dw(&plan, eng, d, d, d, d, B, Ci, Co, d, d, key, d, d, d, d, nullptr) == PNO_OK;
        break;
    case PNO_STAGE_LOCAL_FWD:
        ok = Ci > 8 && Co > 4 && tc_local_fwd(eng, d, B, Ci, Co, H * W, d, d, d, d, key, 0, d, d, nullptr) == PNO_OK;
        break;
    case PNO_STAGE_LOCAL_BWD:
        ok = Ci > 4 && Co > 4 && tc_local_bwd(eng, d, d, d, B, Ci, Co, H * W, d, d, d, d, d, d, nullptr) == PNO_OK;
        break;
    default:
        ok = false;
    }
    g_pno_dry_run = 0;
    return ok ? 1 : 0;
}

extern "C" int pno_set_dynamic_noise(const uint32_t* device_key) {
    g_pno_dyn_key = device_key;
    return PNO_OK;
}

// Engine dispatch of one stage.  `engine` = PNO_ENGINE_* code, optionally ORed with PNO_ENGINE_ALLOW_FALLBACK.  A tensor-core
// engine whose tiling does not cover the shape fails with PNO_E_UNSUPPORTED unless the caller opted into the CUDA-core
// fallback; either way the stage counters record which engine actually ran.
template <typename TC, typename SIMT>
static int dispatch(int engine, int stage, const char* who, TC tc, SIMT simt) {
    const int eng = engine & 0xFF;
    if (eng != PNO_ENGINE_FP32 && eng != PNO_ENGINE_TC_TF32X3 && eng != PNO_ENGINE_TC_TF32)
        PNO_FAIL(PNO_E_INVALID, "%s: unknown engine %d", who, eng);
    if (eng != PNO_ENGINE_FP32) {
        const int rc = tc(eng);
        if (rc == PNO_OK) { ++g_stage_counts[stage][1]; return rc; }
        if (rc != PNO_E_UNSUPPORTED || !(engine & PNO_ENGINE_ALLOW_FALLBACK)) return rc;
    }
    const int rc = simt();
    if (rc == PNO_OK) ++g_stage_counts[stage][0];
    return rc;
}

extern "C" int pno_debug_set_buffer(void* device_ptr) {
    g_pno_debug_buffer = static_cast<long long*>(device_ptr);
    return PNO_OK;
}

extern "C" int pno_version(void) { return PNO_VERSION; }
extern "C" uint64_t pno_launch_count(void) { return g_pno_launches; }
extern "C" int pno_set_sm_reserve(int n) {
    if (n < 0 || n > 64) PNO_FAIL(PNO_E_INVALID, "pno_set_sm_reserve: 0 <= n <= 64 (got %d)", n);
    g_pno_sm_reserve = n;
    return PNO_OK;
}
extern "C" const char* pno_last_error(void) { return g_err; }

extern "C" int pno_check_device(void) {
    int dev = 0;
    PNO_CUDA(cudaGetDevice(&dev));
    int major = 0;
    PNO_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) PNO_FAIL(PNO_E_ARCH, "libpno_b200 is built for sm_100a only; device %d is sm_%d*", dev, major);
    return PNO_OK;
}

extern "C" int pno_plan_create(int H, int W, int m1, int m2, pno_plan_t** out) {
    if (!out) PNO_FAIL(PNO_E_INVALID, "pno_plan_create: null output");
    *out = nullptr;
    if (H <= 0 || W <= 0 || m1 <= 0 || m2 <= 0) PNO_FAIL(PNO_E_INVALID, "pno_plan_create: non-positive size");
    if (m1 > H || m2 > W / 2 + 1)
        PNO_FAIL(PNO_E_MODES, "modes (%d, %d) exceed the spectrum of a %dx%d grid (%d, %d)", m1, m2, H, W, H, W / 2 + 1);
    int rc = pno_check_device();
    if (rc) return rc;
    pno_plan* p = new pno_plan();
    p->H = H; p->W = W; p->m1 = m1; p->m2 = m2; p->M = 2 * m1 * m2;
    p->tc_tables = nullptr;
    p->tc_tables_inv = nullptr;
    p->tc_tables2 = nullptr;
    PNO_CUDA(cudaGetDevice(&p->device));
    std::vector<float2> th(H), tw(W);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < H; ++n) th[n] = make_float2((float)cos(two_pi * n / H), (float)sin(two_pi * n / H));
    for (int n = 0; n < W; ++n) tw[n] = make_float2((float)cos(two_pi * n / W), (float)sin(two_pi * n / W));
    std::vector<int> kx(2 * m1);
    std::vector<float> coef((size_t)2 * m1 * m2);
    for (int j = 0; j < m1; ++j) { kx[j] = j; kx[m1 + j] = H - m1 + j; }
    for (int r = 0; r < 2 * m1; ++r) {
        // reference models.py:93-94: the second slice assignment overwrites the first where they overlap
        const bool alive = r >= m1 || kx[r] < H - m1;
        for (int ky = 0; ky < m2; ++ky) {
            const bool self_conj = ky == 0 || (W % 2 == 0 && ky == W / 2);
            coef[(size_t)r * m2 + ky] = alive ? (float)((self_conj ? 1.0 : 2.0) / ((double)H * W)) : 0.f;
        }
    }
    PNO_CUDA(cudaMalloc(&p->tabH, sizeof(float2) * H));
    PNO_CUDA(cudaMalloc(&p->tabW, sizeof(float2) * W));
    PNO_CUDA(cudaMalloc(&p->kx, sizeof(int) * 2 * m1));
    PNO_CUDA(cudaMalloc(&p->coef, sizeof(float) * coef.size()));
    PNO_CUDA(cudaMemcpy(p->tabH, th.data(), sizeof(float2) * H, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMemcpy(p->tabW, tw.data(), sizeof(float2) * W, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMemcpy(p->kx, kx.data(), sizeof(int) * 2 * m1, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMemcpy(p->coef, coef.data(), sizeof(float) * coef.size(), cudaMemcpyHostToDevice));
    *out = p;
    return PNO_OK;
}

extern "C" int pno_plan_destroy(pno_plan_t* p) {
    if (!p) return PNO_OK;
    tc_plan_release(p);
    tc_plan_release_inv(p);
    tc_plan_release2(p);
    cudaFree(p->tabH);
    cudaFree(p->tabW);
    cudaFree(p->kx);
    cudaFree(p->coef);
    delete p;
    return PNO_OK;
}

extern "C" size_t pno_spectral_workspace_bytes(const pno_plan_t* p, int B, int Ci, int Co) {
    if (!p || B <= 0 || Ci <= 0 || Co <= 0) return 0;
    const size_t mb = (size_t)p->M * B;
    return sizeof(float) * (2 * mb * Ci + 2 * mb * Co + 2 * mb * Ci);
}

static int check_common(const pno_plan* p, int B, int Ci, int Co, const char* who) {
    if (!p) PNO_FAIL(PNO_E_INVALID, "%s: null plan", who);
    if (B <= 0 || Ci <= 0 || Co <= 0) PNO_FAIL(PNO_E_INVALID, "%s: non-positive B/Ci/Co (%d, %d, %d)", who, B, Ci, Co);
    return PNO_OK;
}

extern "C" int pno_dft_fwd(const pno_plan_t* p, int engine, const float* x, int B, int C, int grad_scale, float* xr,
                           float* xi, void* stream) {
    int rc = check_common(p, B, C, C, "pno_dft_fwd");
    if (rc) return rc;
    if (!x || !xr || !xi) PNO_FAIL(PNO_E_INVALID, "pno_dft_fwd: null pointer");
    return dispatch(engine, PNO_STAGE_DFT_FWD, "pno_dft_fwd",
                    [&](int eng) { return tc_dft_fwd(p, eng, x, B, C, grad_scale, xr, xi, as_stream(stream)); },
                    [&]() { return simt_dft_fwd(p, x, B, C, grad_scale, xr, xi, as_stream(stream)); });
}

extern "C" int pno_dft_inv(const pno_plan_t* p, int engine, const float* yr, const float* yi, int B, int C, int adjoint,
                           const float* addend, int act, float* y, float* pre_act, void* stream) {
    int rc = check_common(p, B, C, C, "pno_dft_inv");
    if (rc) return rc;
    if (!yr || !yi || !y) PNO_FAIL(PNO_E_INVALID, "pno_dft_inv: null pointer");
    return dispatch(engine, PNO_STAGE_DFT_INV, "pno_dft_inv",
                    [&](int eng) { return tc_dft_inv(p, eng, yr, yi, B, C, adjoint, addend, act, y, pre_act, as_stream(stream)); },
                    [&]() { return simt_dft_inv(p, yr, yi, B, C, adjoint, addend, act, y, pre_act, as_stream(stream)); });
}

static int check_lv(const float* lv1, const float* lv2, const char* who) {
    if ((lv1 == nullptr) != (lv2 == nullptr)) PNO_FAIL(PNO_E_INVALID, "%s: lv1 and lv2 must both be set or both be NULL", who);
    return PNO_OK;
}

extern "C" int pno_mix_fwd(const pno_plan_t* p, int engine, const float* xr, const float* xi, int B, int Ci, int Co,
                           const float* mu1, const float* mu2, const float* lv1, const float* lv2, uint64_t seed,
                           uint32_t sample_idx, uint32_t site, float* yr, float* yi, void* stream) {
    int rc = check_common(p, B, Ci, Co, "pno_mix_fwd");
    if (rc) return rc;
    if ((rc = check_lv(lv1, lv2, "pno_mix_fwd"))) return rc;
    if (!xr || !xi || !mu1 || !mu2 || !yr || !yi) PNO_FAIL(PNO_E_INVALID, "pno_mix_fwd: null pointer");
    const PhiloxKey key = make_key(seed, sample_idx, site);
    return dispatch(engine, PNO_STAGE_MIX_FWD, "pno_mix_fwd",
                    [&](int eng) { return tc_mix_fwd(p, eng, xr, xi, B, Ci, Co, mu1, mu2, lv1, lv2, key, yr, yi, as_stream(stream)); },
                    [&]() { return simt_mix_fwd(p, xr, xi, B, Ci, Co, mu1, mu2, lv1, lv2, key, yr, yi, as_stream(stream)); });
}

extern "C" int pno_mix_bwd_dx(const pno_plan_t* p, int engine, const float* gr, const float* gi, int B, int Ci, int Co,
                              const float* mu1, const float* mu2, const float* lv1, const float* lv2, uint64_t seed,
                              uint32_t sample_idx, uint32_t site, float* dxr, float* dxi, void* stream) {
    int rc = check_common(p, B, Ci, Co, "pno_mix_bwd_dx");
    if (rc) return rc;
    if ((rc = check_lv(lv1, lv2, "pno_mix_bwd_dx"))) return rc;
    if (!gr || !gi || !mu1 || !mu2 || !dxr || !dxi) PNO_FAIL(PNO_E_INVALID, "pno_mix_bwd_dx: null pointer");
    const PhiloxKey key = make_key(seed, sample_idx, site);
    return dispatch(engine, PNO_STAGE_MIX_BWD_DX, "pno_mix_bwd_dx",
                    [&](int eng) { return tc_mix_bwd_dx(p, eng, gr, gi, B, Ci, Co, mu1, mu2, lv1, lv2, key, dxr, dxi, as_stream(stream)); },
                    [&]() { return simt_mix_bwd_dx(p, gr, gi, B, Ci, Co, mu1, mu2, lv1, lv2, key, dxr, dxi, as_stream(stream)); });
}

extern "C" int pno_mix_bwd_dw(const pno_plan_t* p, int engine, const float* xr, const float* xi, const float* gr,
                              const float* gi, int B, int Ci, int Co, const float* lv1, const float* lv2, uint64_t seed,
                              uint32_t sample_idx, uint32_t site, float* dmu1, float* dmu2, float* dlv1, float* dlv2,
                              void* stream) {
    int rc = check_common(p, B, Ci, Co, "pno_mix_bwd_dw");
    if (rc) return rc;
    if ((rc = check_lv(lv1, lv2, "pno_mix_bwd_dw"))) return rc;
    if (!xr || !xi || !gr || !gi) PNO_FAIL(PNO_E_INVALID, "pno_mix_bwd_dw: null pointer");
    if ((dmu1 == nullptr) != (dmu2 == nullptr) || (dlv1 == nullptr) != (dlv2 == nullptr))
        PNO_FAIL(PNO_E_INVALID, "pno_mix_bwd_dw: gradient outputs must be given per pair");
    const PhiloxKey key = make_key(seed, sample_idx, site);
    return dispatch(engine, PNO_STAGE_MIX_BWD_DW, "pno_mix_bwd_dw",
                    [&](int eng) {
                        return tc_mix_bwd_dw(p, eng, xr, xi, gr, gi, B, Ci, Co, lv1, lv2, key, dmu1, dmu2, dlv1, dlv2, as_stream(stream));
                    },
                    [&]() {
                        return simt_mix_bwd_dw(p, xr, xi, gr, gi, B, Ci, Co, lv1, lv2, key, dmu1, dmu2, dlv1, dlv2, as_stream(stream));
                    });
}

// ---- mode windows (mode-parallel sharding of the spectral weights): tensor-core engines only ----
static int check_window(const pno_plan* p, int engine, const char* who) {
    if (!p) PNO_FAIL(PNO_E_INVALID, "%s: null plan", who);
    const int eng = engine & 0xFF;
    if (eng != PNO_ENGINE_TC_TF32X3 && eng != PNO_ENGINE_TC_TF32)
        PNO_FAIL(PNO_E_UNSUPPORTED, "%s: mode windows run on the tensor-core engines (tf32x3 / tf32) only", who);
    return PNO_OK;
}

extern "C" int pno_mix_fwd_window(const pno_plan_t* p, int engine, const float* xr, const float* xi, int B, int Ci, int Co,
                                  const float* mu, const float* lv, uint64_t seed, uint32_t sample_idx, uint32_t site,
                                  int mode_first, int mode_count, float* yr, float* yi, void* stream) {
    int rc = check_window(p, engine, "pno_mix_fwd_window");
    if (rc) return rc;
    if (!xr || !xi || !mu || !yr || !yi || mode_count < 1) PNO_FAIL(PNO_E_INVALID, "pno_mix_fwd_window: bad arguments");
    rc = tc_mix_fwd(p, engine & 0xFF, xr, xi, B, Ci, Co, mu, mu, lv, lv, make_key(seed, sample_idx, site), yr, yi, as_stream(stream),
                    mode_first, mode_count);
    if (rc == PNO_OK) ++g_stage_counts[PNO_STAGE_MIX_FWD][1];
    return rc;
}

extern "C" int pno_mix_bwd_dx_window(const pno_plan_t* p, int engine, const float* gr, const float* gi, int B, int Ci, int Co,
                                     const float* mu, const float* lv, uint64_t seed, uint32_t sample_idx, uint32_t site,
                                     int mode_first, int mode_count, float* dxr, float* dxi, void* stream) {
    int rc = check_window(p, engine, "pno_mix_bwd_dx_window");
    if (rc) return rc;
    if (!gr || !gi || !mu || !dxr || !dxi || mode_count < 1) PNO_FAIL(PNO_E_INVALID, "pno_mix_bwd_dx_window: bad arguments");
    rc = tc_mix_bwd_dx(p, engine & 0xFF, gr, gi, B, Ci, Co, mu, mu, lv, lv, make_key(seed, sample_idx, site), dxr, dxi,
                       as_stream(stream), mode_first, mode_count);
    if (rc == PNO_OK) ++g_stage_counts[PNO_STAGE_MIX_BWD_DX][1];
    return rc;
}

extern "C" int pno_mix_bwd_dw_window(const pno_plan_t* p, int engine, const float* xr, const float* xi, const float* gr,
                                     const float* gi, int B, int Ci, int Co, const float* lv, uint64_t seed, uint32_t sample_idx,
                                     uint32_t site, int mode_first, int mode_count, float* dmu, float* dlv, void* stream) {
    int rc = check_window(p, engine, "pno_mix_bwd_dw_window");
    if (rc) return rc;
    if (!xr || !xi || !gr || !gi || !dmu || mode_count < 1) PNO_FAIL(PNO_E_INVALID, "pno_mix_bwd_dw_window: bad arguments");
    rc = tc_mix_bwd_dw(p, engine & 0xFF, xr, xi, gr, gi, B, Ci, Co, lv, lv, make_key(seed, sample_idx, site), dmu, dmu, dlv, dlv,
                       as_stream(stream), mode_first, mode_count);
    if (rc == PNO_OK) ++g_stage_counts[PNO_STAGE_MIX_BWD_DW][1];
    return rc;
}

extern "C" int pno_spectral_fwd(const pno_plan_t* p, int engine, const float* x, int B, int Ci, int Co, const float* mu1,
                                const float* mu2, const float* lv1, const float* lv2, uint64_t seed, uint32_t sample_idx,
                                uint32_t site, const float* addend, int act, float* y, float* pre_act, void* workspace,
                                void* stream) {
    int rc = check_common(p, B, Ci, Co, "pno_spectral_fwd");
    if (rc) return rc;
    if (!x || !y || !workspace) PNO_FAIL(PNO_E_INVALID, "pno_spectral_fwd: null pointer");
    const size_t mb = (size_t)p->M * B;
    float* xr = static_cast<float*>(workspace);
    float* xi = xr + mb * Ci;
    float* yr = xi + mb * Ci;
    float* yi = yr + mb * Co;
    if ((rc = pno_dft_fwd(p, engine, x, B, Ci, 0, xr, xi, stream))) return rc;
    if ((rc = pno_mix_fwd(p, engine, xr, xi, B, Ci, Co, mu1, mu2, lv1, lv2, seed, sample_idx, site, yr, yi, stream)))
        return rc;
    return pno_dft_inv(p, engine, yr, yi, B, Co, 0, addend, act, y, pre_act, stream);
}

extern "C" int pno_spectral_bwd(const pno_plan_t* p, int engine, const float* x, const float* xhat_saved, const float* g,
                                int B, int Ci, int Co, const float* mu1, const float* mu2, const float* lv1,
                                const float* lv2, uint64_t seed, uint32_t sample_idx, uint32_t site, float* dx,
                                float* dmu1, float* dmu2, float* dlv1, float* dlv2, void* workspace, void* stream) {
    int rc = check_common(p, B, Ci, Co, "pno_spectral_bwd");
    if (rc) return rc;
    if (!g || !workspace) PNO_FAIL(PNO_E_INVALID, "pno_spectral_bwd: null pointer");
    const size_t mb = (size_t)p->M * B;
    float* xr = static_cast<float*>(workspace);
    float* xi = xr + mb * Ci;
    float* gr = xi + mb * Ci;
    float* gi = gr + mb * Co;
    float* dxr = gi + mb * Co;
    float* dxi = dxr + mb * Ci;
    if ((rc = pno_dft_fwd(p, engine, g, B, Co, 1, gr, gi, stream))) return rc;
    if (dmu1 || dlv1) {
        const float *sxr = xr, *sxi = xi;
        if (xhat_saved) {
            sxr = xhat_saved;
            sxi = xhat_saved + mb * Ci;
        } else {
            if (!x) PNO_FAIL(PNO_E_INVALID, "pno_spectral_bwd: need x or xhat_saved for the weight gradients");
            if ((rc = pno_dft_fwd(p, engine, x, B, Ci, 0, xr, xi, stream))) return rc;
        }
        if ((rc = pno_mix_bwd_dw(p, engine, sxr, sxi, gr, gi, B, Ci, Co, lv1, lv2, seed, sample_idx, site, dmu1, dmu2,
                                 dlv1, dlv2, stream)))
            return rc;
    }
    if (dx) {
        if ((rc = pno_mix_bwd_dx(p, engine, gr, gi, B, Ci, Co, mu1, mu2, lv1, lv2, seed, sample_idx, site, dxr, dxi,
                                 stream)))
            return rc;
        if ((rc = pno_dft_inv(p, engine, dxr, dxi, B, Ci, 1, nullptr, PNO_ACT_NONE, dx, nullptr, stream))) return rc;
    }
    return PNO_OK;
}

extern "C" size_t pno_local_bwd_workspace_bytes(int B, int Ci, int Co, int HW, int noisy) {
    if (B <= 0 || Ci <= 0 || Co <= 0 || HW <= 0) return 0;
    return tc_local_bwd_workspace_bytes(B, Ci, Co, HW, noisy);
}

extern "C" int pno_local_bwd(int engine, const float* x, const float* g, const float* dout_dlv, int B, int Ci, int Co, int HW,
                             const float* Wm, const float* Wl, float* dx, float* dw, float* db, void* workspace, void* stream) {
    if (B <= 0 || Ci <= 0 || Co <= 0 || HW <= 0) PNO_FAIL(PNO_E_INVALID, "pno_local_bwd: non-positive size");
    if (!x || !g || !Wm || !workspace) PNO_FAIL(PNO_E_INVALID, "pno_local_bwd: null pointer");
    // lift- (Ci <= 4, no dx) and projection-shaped (Co <= 4) layers are streaming CUDA-core kernels whatever the engine
    const bool thin = HW % 4 == 0 && ((Ci <= 4 && dx == nullptr) || Co <= 4);
    if (thin) {
        const int rc = tc_local_bwd(PNO_ENGINE_TC_TF32X3, x, g, dout_dlv, B, Ci, Co, HW, Wm, Wl, dx, dw, db, workspace, as_stream(stream));
        if (rc == PNO_OK) { ++g_stage_counts[PNO_STAGE_LOCAL_THIN][0]; return rc; }
        if (rc != PNO_E_UNSUPPORTED) return rc;       // (misaligned thin layers take the general kernels below)
        return simt_local_bwd(x, g, dout_dlv, B, Ci, Co, HW, Wm, Wl, dx, dw, db, as_stream(stream));
    }
    return dispatch(engine, PNO_STAGE_LOCAL_BWD, "pno_local_bwd",
                    [&](int eng) { return tc_local_bwd(eng, x, g, dout_dlv, B, Ci, Co, HW, Wm, Wl, dx, dw, db, workspace, as_stream(stream)); },
                    [&]() { return simt_local_bwd(x, g, dout_dlv, B, Ci, Co, HW, Wm, Wl, dx, dw, db, as_stream(stream)); });
}

extern "C" int pno_local_fwd(int engine, const float* x, int B, int Ci, int Co, int HW, const float* Wm, const float* bm,
                             const float* Wl, const float* bl, uint64_t seed, uint32_t sample_idx, uint32_t site,
                             uint64_t batch_offset, float* out, float* dout_dlv, void* stream) {
    if (B <= 0 || Ci <= 0 || Co <= 0 || HW <= 0) PNO_FAIL(PNO_E_INVALID, "pno_local_fwd: non-positive size");
    if (!x || !Wm || !out) PNO_FAIL(PNO_E_INVALID, "pno_local_fwd: null pointer");
    if (B > 65535) PNO_FAIL(PNO_E_INVALID, "pno_local_fwd: batch %d exceeds the grid limit", B);
    const PhiloxKey key = make_key(seed, sample_idx, site);
    // lift- (Ci <= 8) and projection-shaped (Co <= 4) layers are streaming CUDA-core kernels whatever the engine: there is no
    // GEMM in them to put on the tensor cores
    if (Ci <= 8 || Co <= 4) {
        const int rc = simt_local_fwd(x, B, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv, as_stream(stream));
        if (rc == PNO_OK) ++g_stage_counts[PNO_STAGE_LOCAL_THIN][0];
        return rc;
    }
    return dispatch(engine, PNO_STAGE_LOCAL_FWD, "pno_local_fwd",
                    [&](int eng) { return tc_local_fwd(eng, x, B, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv, as_stream(stream)); },
                    [&]() { return simt_local_fwd(x, B, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv, as_stream(stream)); });
}
