// Shared device/host helpers of libpno_b200: error plumbing, the plan, the Philox noise stream, activations.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pno_b200.h"

// ---------------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------------
void pno_set_error(const char* fmt, ...);

#define PNO_FAIL(code, ...)          \
    do {                             \
        pno_set_error(__VA_ARGS__);  \
        return (code);               \
    } while (0)

#define PNO_CUDA(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) PNO_FAIL(PNO_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                        __FILE__, __LINE__);                                            \
    } while (0)

extern long long* g_pno_debug_buffer;        // device buffer for kernel phase counters (NULL = off)
extern unsigned long long g_pno_launches;   // kernels launched by this library (pno_launch_count)
// SMs the persistent kernels leave free (pno_set_sm_reserve): a concurrent NCCL kernel needs somewhere to run -- every kernel here
// owns a whole SM (shared memory, TMEM), so with grids of all 148 SMs a gradient all-reduce launched during backward only runs in
// the gaps between our kernels
extern int g_pno_sm_reserve;
static inline int pno_usable_sms(int sms) { return sms - g_pno_sm_reserve < 1 ? 1 : sms - g_pno_sm_reserve; }
#define PNO_LAUNCH_CHECK()            \
    do {                              \
        ++g_pno_launches;             \
        PNO_CUDA(cudaGetLastError()); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------------------------
struct pno_plan {
    int H, W, m1, m2;
    int M;            // 2*m1*m2 retained modes
    int device;
    // device tables
    float2* tabH;     // [H]   (cos, sin)(2 pi n / H)
    float2* tabW;     // [W]
    int* kx;          // [2*m1] actual kx of retained row kx2
    float* coef;      // [2*m1*m2] alive(kx2) * c_ky / (H*W)
    // dense twiddle matrices for the tensor-core engines (allocated lazily by the engines that use them)
    void* tc_tables;
    void* tc_tables_inv;
    void* tc_tables2;   // tables of the two-pass transforms (pno_tc_dft2.cu)
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Library-owned scratch: one grow-only buffer per (device, stream, slot), so that calls on different streams never share one
// (calls on one stream are ordered).  A buffer that is outgrown is NOT freed: a captured CUDA graph may still hold its address.
// Slots: 0 = tf32 hi / lo parts of 1x1 weights / twiddle tables, 1 and 2 = intermediates of the two-pass transforms.
float* pno_scratch(int slot, size_t bytes, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------------
// single-instruction approximations (flush-to-zero forms: no denormal fix-up code around the MUFU)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_exp(float x) { return fast_ex2(x * 1.4426950408889634f); }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---------------------------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller.  Normative definition: oracle/philox.py.
// ---------------------------------------------------------------------------------------------------------------------
struct PhiloxKey {
    uint32_t k0, k1, sample, site;
    // Device-resident key (pno_set_dynamic_noise): when non-NULL the kernel reads (seed lo, seed hi, sample_idx) from these three
    // words at run time instead of the by-value fields above, so that a captured CUDA graph of a sampled forward can be replayed
    // for every MC sample / training step without re-recording its launch arguments.
    const uint32_t* dyn;
};

extern thread_local int g_pno_dry_run;               // pno_engine_supports: validate the shape / geometry only
extern thread_local const uint32_t* g_pno_dyn_key;   // set by pno_set_dynamic_noise (this host thread), NULL by default

static inline PhiloxKey make_key(uint64_t seed, uint32_t sample, uint32_t site) {
    PhiloxKey k;
    k.k0 = (uint32_t)seed;
    k.k1 = (uint32_t)(seed >> 32);
    k.sample = sample;
    k.site = site;
    k.dyn = g_pno_dyn_key;
    return k;
}

// every kernel that draws noise calls this once (per thread, before its main loop)
__device__ __forceinline__ PhiloxKey resolve_key(PhiloxKey k) {
    if (k.dyn != nullptr) {
        k.k0 = __ldg(k.dyn);
        k.k1 = __ldg(k.dyn + 1);
        k.sample = __ldg(k.dyn + 2);
    }
    return k;
}

// PNO_PHILOX_ROUNDS exists to MEASURE what the round count costs; the normative stream is 10 rounds
// (oracle/philox.py) and every parity test assumes it.
#ifndef PNO_PHILOX_ROUNDS
#define PNO_PHILOX_ROUNDS 10
#endif
__device__ __forceinline__ uint4 philox_quad_bits(const PhiloxKey& key, uint64_t q) {
    uint32_t c0 = (uint32_t)q, c1 = (uint32_t)(q >> 32), c2 = key.sample, c3 = key.site;
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < PNO_PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// u in (0,1): ((x >> 9) + 0.5) * 2^-23 is exact in fp32 (and so is u2 - 0.5).  theta in (-pi, pi) keeps the fast sin/cos in
// their accurate range.  Branch-free: lg2.approx / sqrt.approx / sin.approx / cos.approx are one MUFU each (|error| on the
// normals ~1e-6 absolute, inside the 5e-6 the parity test allows against the fp64 oracle stream).
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = fmaf((float)(a >> 9), 1.1920928955078125e-7f, 5.9604644775390625e-8f);
    const float u2m = fmaf((float)(b >> 9), 1.1920928955078125e-7f, 5.9604644775390625e-8f - 0.5f);
    const float r = fast_sqrt(fast_lg2(u1) * -1.3862943611198906f);      // sqrt(-2 ln u1)
    const float t = 6.283185307179586f * u2m;
    return make_float2(r * fast_cos(t), r * fast_sin(t));
}

__device__ __forceinline__ float4 philox_quad_normals(const PhiloxKey& key, uint64_t q) {
    const uint4 x = philox_quad_bits(key, q);
    const float2 a = box_muller(x.x, x.y), b = box_muller(x.z, x.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

// one normal of an activation-shaped site (slow path for unaligned elements)
__device__ __forceinline__ float philox_normal_at(const PhiloxKey& key, uint64_t e) {
    const float4 z = philox_quad_normals(key, e >> 2);
    const int lane = (int)(e & 3);
    return lane == 0 ? z.x : lane == 1 ? z.y : lane == 2 ? z.z : z.w;
}

// quad index of spectral weight (blk, kx, ky, i, o>>1); mode_in_blk = kx*m2+ky
__device__ __forceinline__ uint64_t spectral_quad(int blk, int m1m2, int mode_in_blk, int Ci, int Cop, int i, int opair) {
    return ((uint64_t)((uint64_t)blk * m1m2 + mode_in_blk) * Ci + i) * Cop + opair;
}

// ---------------------------------------------------------------------------------------------------------------------
// activations (models.py:224-229: exact-erf GELU, ReLU)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == PNO_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    if (act == PNO_ACT_RELU) return v > 0.f ? v : 0.f;
    return v;
}

// GELU for the tensor-core epilogues: erfc through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on erf), evaluated in the
// complementary form on the negative side so that the small tail keeps its relative accuracy.  ~14 instructions vs ~40.
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float ec = p * t * fast_ex2(z * z * -1.4426950408889634f);          // erfc(|x|/sqrt2)
    return fmaf(z * -0.70710678118654752f, ec, fmaxf(x, 0.f));             // max(x,0) - |x|/2 * erfc(|x|/sqrt2)
}
__device__ __forceinline__ float act_apply_fast(float v, int act) {
    if (act == PNO_ACT_GELU) return gelu_fast(v);
    if (act == PNO_ACT_RELU) return v > 0.f ? v : 0.f;
    return v;
}

__device__ __forceinline__ float act_grad(float v, int act) {
    if (act == PNO_ACT_GELU) {
        const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752f));
        const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
        return cdf + v * pdf;
    }
    if (act == PNO_ACT_RELU) return v > 0.f ? 1.f : 0.f;
    return 1.f;
}

// ---------------------------------------------------------------------------------------------------------------------
// engine entry points implemented in the other translation units
// ---------------------------------------------------------------------------------------------------------------------
int simt_dft_fwd(const pno_plan* p, const float* x, int B, int C, int grad_scale, float* xr, float* xi, cudaStream_t st);
int simt_dft_inv(const pno_plan* p, const float* yr, const float* yi, int B, int C, int adjoint, const float* addend,
                 int act, float* y, float* pre_act, cudaStream_t st);
int simt_mix_fwd(const pno_plan* p, const float* xr, const float* xi, int B, int Ci, int Co, const float* mu1,
                 const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* yr, float* yi,
                 cudaStream_t st);
int simt_mix_bwd_dx(const pno_plan* p, const float* gr, const float* gi, int B, int Ci, int Co, const float* mu1,
                    const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* dxr, float* dxi,
                    cudaStream_t st);
int simt_mix_bwd_dw(const pno_plan* p, const float* xr, const float* xi, const float* gr, const float* gi, int B, int Ci,
                    int Co, const float* lv1, const float* lv2, PhiloxKey key, float* dmu1, float* dmu2, float* dlv1,
                    float* dlv2, cudaStream_t st);
int simt_local_fwd(const float* x, int B, int Ci, int Co, int HW, const float* Wm, const float* bm, const float* Wl,
                   const float* bl, PhiloxKey key, uint64_t batch_offset, float* out, float* dout_dlv, cudaStream_t st);
int simt_local_bwd(const float* x, const float* g, const float* dout_dlv, int B, int Ci, int Co, int HW, const float* Wm,
                   const float* Wl, float* dx, float* dw, float* db, cudaStream_t st);
