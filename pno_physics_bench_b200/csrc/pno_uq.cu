// Loss / calibration epilogue on device: one pass over (mean, spread, target) produces every
// sufficient statistic that the reference's losses (training/losses.py:97-149) and CalibrationMetrics
// (metrics.py:32-170, 295-392) reduce to, as fp64 sums that are additive over batches -- so PNOTrainer.evaluate never copies
// predictions to the host (trainer.py:540-552 does, per batch) and the metrics are one small read at the end.
//   pno_uq_stats          the statistics (layout in include/pno_b200.h)
//   pno_loss_data_bwd     gradient of  g_se * sum (mean-target)^2 + g_nll * sum nll  w.r.t. mean and log_var
#include "pno_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxLevels = PNO_UQ_MAX_LEVELS, kMaxAlphas = PNO_UQ_MAX_ALPHAS, kMaxBins = PNO_UQ_MAX_BINS;

struct UqArgs {
    const float *mean, *spread, *target;
    unsigned long long n;
    int spread_kind, n_levels, n_alphas, n_bins;
    float z_levels[kMaxLevels];
    float alphas[kMaxAlphas], z_alphas[kMaxAlphas];
    const float* bin_edges;     // n_bins + 1 boundaries (device), bins are (edge[b], edge[b+1]] like metrics.py:69-75
    double* acc;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kThreads) k_uq_stats(const UqArgs a) {
    __shared__ float s_edges[kMaxBins + 1];
    __shared__ unsigned int s_cnt[kMaxBins], s_hit[kMaxBins];
    __shared__ float s_conf[kMaxBins];
    __shared__ double s_red[kThreads / 32];
    for (int i = threadIdx.x; i < kMaxBins; i += kThreads) { s_cnt[i] = 0; s_hit[i] = 0; s_conf[i] = 0.f; }
    for (int i = threadIdx.x; i <= a.n_bins && a.n_bins > 0; i += kThreads) s_edges[i] = a.bin_edges[i];
    __syncthreads();

    double d_ae = 0, d_se = 0, d_nll = 0, d_s = 0, d_s2 = 0, d_sae = 0, d_is[kMaxAlphas];
    unsigned long long cnt[kMaxLevels];
#pragma unroll
    for (int k = 0; k < kMaxLevels; ++k) cnt[k] = 0;
#pragma unroll
    for (int k = 0; k < kMaxAlphas; ++k) d_is[k] = 0;
    unsigned long long n_mine = 0;

    for (unsigned long long i = blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < a.n;
         i += (unsigned long long)gridDim.x * kThreads) {
        const float mu = a.mean[i], t = a.target[i], sp = a.spread[i];
        const float sd = a.spread_kind ? expf(0.5f * sp) : sp;         // losses.py:60 std = exp(0.5 log_var)
        const float err = fabsf(mu - t);                               // metrics.py:121
        const float d = t - mu;
        const float se = d * d;
        // Gaussian NLL, std clamped at 1e-6 (losses.py:104-115, metrics.py:312-323)
        const float sc = fmaxf(sd, 1e-6f);
        const float nll = 0.5f * (1.8378770664093453f + 2.f * logf(sc) + se / (sc * sc));
        ++n_mine;
        d_ae += err; d_se += se; d_nll += nll; d_s += sd; d_s2 += (double)sd * sd; d_sae += (double)sd * err;
#pragma unroll
        for (int k = 0; k < kMaxLevels; ++k)
            if (k < a.n_levels) cnt[k] += err <= a.z_levels[k] * sd;   // metrics.py:122-125, losses.py:140-143
#pragma unroll
        for (int k = 0; k < kMaxAlphas; ++k)
            if (k < a.n_alphas) {                                      // metrics.py:147-158
                const float lower = mu - a.z_alphas[k] * sd, upper = mu + a.z_alphas[k] * sd;
                const float pen = 2.f / a.alphas[k];
                d_is[k] += (upper - lower) + pen * fmaxf(lower - t, 0.f) + pen * fmaxf(t - upper, 0.f);
            }
        if (a.n_bins > 0) {                                            // metrics.py:58-83
            const float zn = err / (sd + 1e-8f);
            const float pdf = expf(-(zn * zn) / 2.f - 0.91893853320467274f);
            const float conf = fminf(fmaxf(1.f - 2.f * pdf, 0.f), 1.f);
            int b = (int)ceilf(conf * a.n_bins) - 1;                   // first guess, then the reference's own comparisons
            b = max(0, min(a.n_bins - 1, b));
            while (b > 0 && !(conf > s_edges[b])) --b;
            while (b < a.n_bins - 1 && !(conf <= s_edges[b + 1])) ++b;
            if (conf > s_edges[b] && conf <= s_edges[b + 1]) {
                atomicAdd(&s_cnt[b], 1u);
                atomicAdd(&s_conf[b], conf);
                if (zn <= 1.0f) atomicAdd(&s_hit[b], 1u);
            }
        }
    }

    // block reduction of the scalar slots, one double atomic per block and slot
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto block_add = [&](double v, int slot) {
        v = warp_sum(v);
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double w = lane < kThreads / 32 ? s_red[lane] : 0.0;
            w = warp_sum(w);
            if (lane == 0 && w != 0.0) atomicAdd(&a.acc[slot], w);
        }
        __syncthreads();
    };
    block_add((double)n_mine, 0);
    block_add(d_ae, 1);
    block_add(d_se, 2);
    block_add(d_nll, 3);
    block_add(d_s, 4);
    block_add(d_s2, 5);
    block_add(d_sae, 6);
    for (int k = 0; k < a.n_levels; ++k) block_add((double)cnt[k], 7 + k);
    for (int k = 0; k < a.n_alphas; ++k) block_add(d_is[k], 7 + a.n_levels + k);
    const int base = 7 + a.n_levels + a.n_alphas;
    for (int b = threadIdx.x; b < a.n_bins; b += kThreads)
        if (s_cnt[b]) {
            atomicAdd(&a.acc[base + 3 * b], (double)s_cnt[b]);
            atomicAdd(&a.acc[base + 3 * b + 1], (double)s_conf[b]);
            atomicAdd(&a.acc[base + 3 * b + 2], (double)s_hit[b]);
        }
}

__global__ void __launch_bounds__(kThreads) k_loss_data_bwd(const float* __restrict__ mean, const float* __restrict__ lv,
                                                           const float* __restrict__ target, unsigned long long n,
                                                           const double* __restrict__ g, float* __restrict__ dmean,
                                                           float* __restrict__ dlv) {
    const float g_se = (float)g[0], g_nll = (float)g[1];
    for (unsigned long long i = blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * kThreads) {
        const float mu = mean[i], t = target[i];
        const float d = mu - t;
        float gm = g_se * 2.f * d, gl = 0.f;
        if (lv) {
            const float sd = expf(0.5f * lv[i]);
            const float sc = fmaxf(sd, 1e-6f);
            const float iv = 1.f / (sc * sc);
            gm += g_nll * d * iv;                                       // d nll / d mean = (mean - t) / var
            // d nll / d log_var = 0.5 (1 - (t - mean)^2 / var) through std = exp(lv / 2); zero where the clamp is active
            if (sd >= 1e-6f) gl = g_nll * 0.5f * (1.f - d * d * iv);
        }
        dmean[i] = gm;
        if (dlv) dlv[i] = gl;
    }
}

}  // namespace

extern "C" int pno_uq_stats(const float* mean, const float* spread, const float* target, uint64_t n, int spread_kind,
                            const float* z_levels, int n_levels, const float* alphas, const float* z_alphas, int n_alphas,
                            const float* bin_edges, int n_bins, double* acc, void* stream) {
    if (!mean || !spread || !target || !acc) PNO_FAIL(PNO_E_INVALID, "pno_uq_stats: null pointer");
    if (n_levels < 0 || n_levels > kMaxLevels || n_alphas < 0 || n_alphas > kMaxAlphas || n_bins < 0 || n_bins > kMaxBins)
        PNO_FAIL(PNO_E_INVALID, "pno_uq_stats: at most %d coverage levels, %d interval levels, %d bins", kMaxLevels, kMaxAlphas,
                 kMaxBins);
    if ((n_levels && !z_levels) || (n_alphas && (!alphas || !z_alphas)) || (n_bins && !bin_edges))
        PNO_FAIL(PNO_E_INVALID, "pno_uq_stats: missing level / bin arrays");
    if (n == 0) return PNO_OK;
    UqArgs a;
    memset(&a, 0, sizeof(a));
    a.mean = mean; a.spread = spread; a.target = target; a.n = n; a.spread_kind = spread_kind;
    a.n_levels = n_levels; a.n_alphas = n_alphas; a.n_bins = n_bins; a.bin_edges = bin_edges; a.acc = acc;
    for (int k = 0; k < n_levels; ++k) a.z_levels[k] = z_levels[k];           // host arrays (a handful of floats)
    for (int k = 0; k < n_alphas; ++k) { a.alphas[k] = alphas[k]; a.z_alphas[k] = z_alphas[k]; }
    const unsigned long long want = (n + kThreads - 1) / kThreads;
    const int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    k_uq_stats<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_loss_data_bwd(const float* mean, const float* log_var, const float* target, uint64_t n, const double* g,
                                 float* dmean, float* dlog_var, void* stream) {
    if (!mean || !target || !g || !dmean) PNO_FAIL(PNO_E_INVALID, "pno_loss_data_bwd: null pointer");
    if (n == 0) return PNO_OK;
    const unsigned long long want = (n + kThreads - 1) / kThreads;
    const int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    k_loss_data_bwd<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(mean, log_var, target, n, g, dmean, dlog_var);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Optimiser step of the trainer (training/trainer.py:283-291: clip_grad_norm_ then AdamW.step) without host syncs:
//   pno_sumsq_accumulate   acc (double, device) += sum g^2                         -> global gradient norm
//   pno_adamw_step         clip coefficient from the device-side norm, decoupled weight decay, Adam moments, update
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// analytic KL gradient of a spectral parameter (models.py:100-104), added to the data gradient on the fly:
//   kind 1 (mean, re / im parts): gs * p          kind 2 (log-variance): gs * (-0.5 (1 - exp(p)))
__device__ __forceinline__ float kl_extra(float p, int kind, float gs) {
    return kind == 1 ? gs * p : gs * (-0.5f * (1.0f - expf(p)));
}

__global__ void __launch_bounds__(kThreads) k_sumsq(const float* __restrict__ g, unsigned long long n, double* __restrict__ acc,
                                                   const float* __restrict__ p, int kl_kind, const float* __restrict__ kl_gscale) {
    __shared__ double s_red[kThreads / 32];
    double s = 0.0;
    const unsigned long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float gs = kl_kind ? *kl_gscale : 0.f;
    for (unsigned long long i = blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < n4; i += (unsigned long long)gridDim.x * kThreads) {
        float4 v = g4[i];
        if (kl_kind) {
            const float4 q = reinterpret_cast<const float4*>(p)[i];
            v.x += kl_extra(q.x, kl_kind, gs); v.y += kl_extra(q.y, kl_kind, gs);
            v.z += kl_extra(q.z, kl_kind, gs); v.w += kl_extra(q.w, kl_kind, gs);
        }
        s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    for (unsigned long long i = (n4 << 2) + blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * kThreads) {
        const float v = g[i] + (kl_kind ? kl_extra(p[i], kl_kind, gs) : 0.f);
        s += (double)v * v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < kThreads / 32 ? s_red[threadIdx.x] : 0.0;
        w = warp_sum(w);
        if (threadIdx.x == 0 && w != 0.0) atomicAdd(acc, w);
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt, max_norm;
    int kl_kind;
    float kl_gs;          // filled in the kernel from device memory
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float clip) {
    if (a.kl_kind) g += kl_extra(p, a.kl_kind, a.kl_gs);
    g *= clip;
    p *= 1.f - a.lr * a.weight_decay;                                   // decoupled weight decay (torch AdamW)
    m = fmaf(1.f - a.beta1, g - m, m);                                  // m.lerp_(g, 1 - beta1)
    v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);
    const float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
    p -= (a.lr / a.bias_c1) * (m / denom);
    return g;
}

__global__ void __launch_bounds__(kThreads) k_adamw(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, unsigned long long n, AdamArgs a,
                                                   const double* __restrict__ gsumsq, int write_grad,
                                                   const float* __restrict__ kl_gscale) {
    a.kl_gs = a.kl_kind ? *kl_gscale : 0.f;
    // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1 (torch/nn/utils/clip_grad.py)
    float clip = 1.f;
    if (gsumsq && a.max_norm > 0.f) clip = fminf(1.f, a.max_norm / ((float)sqrt(*gsumsq) + 1e-6f));
    const unsigned long long n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4*>(p), *g4 = reinterpret_cast<float4*>(g), *m4 = reinterpret_cast<float4*>(m),
           *v4 = reinterpret_cast<float4*>(v);
    for (unsigned long long i = blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < n4; i += (unsigned long long)gridDim.x * kThreads) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
        gg.x = adam_one(pp.x, gg.x, mm.x, vv.x, a, clip);
        gg.y = adam_one(pp.y, gg.y, mm.y, vv.y, a, clip);
        gg.z = adam_one(pp.z, gg.z, mm.z, vv.z, a, clip);
        gg.w = adam_one(pp.w, gg.w, mm.w, vv.w, a, clip);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
        if (write_grad) g4[i] = gg;
    }
    for (unsigned long long i = (n4 << 2) + blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * kThreads) {
        float pp = p[i], mm = m[i], vv = v[i];
        const float gg = adam_one(pp, g[i], mm, vv, a, clip);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (write_grad) g[i] = gg;
    }
}

// ---- multi-tensor forms: ONE launch over a table of tensors (the per-tensor forms cost 2 launches x ~40 tensors per step) ----
constexpr int kMtMax = 40;                 // tensors per launch (table travels as a kernel parameter: 40 x 48 B + header < 4 KB)
constexpr int kMtChunk = kThreads * 4 * 8; // floats per block
struct MtTensor {
    float *p, *g, *m, *v;
    unsigned long long n;
    int kl_kind;
    unsigned first_block;
};
struct MtTable {
    MtTensor t[kMtMax];
    int count;
};

__device__ __forceinline__ int mt_find(const MtTable& tab, unsigned block) {
    int k = 0;
    while (k + 1 < tab.count && tab.t[k + 1].first_block <= block) ++k;
    return k;
}

__global__ void __launch_bounds__(kThreads) k_sumsq_multi(const __grid_constant__ MtTable tab, double* __restrict__ acc,
                                                         const float* __restrict__ kl_gscale) {
    __shared__ double s_red[kThreads / 32];
    const int k = mt_find(tab, blockIdx.x);
    const MtTensor& t = tab.t[k];
    const float gs = t.kl_kind ? *kl_gscale : 0.f;
    const unsigned long long lo = (unsigned long long)(blockIdx.x - t.first_block) * kMtChunk;
    const unsigned long long hi = lo + kMtChunk < t.n ? lo + kMtChunk : t.n;
    double s = 0.0;
    const unsigned long long v_hi = lo + ((hi - lo) & ~3ull);
    for (unsigned long long i = lo + 4ull * threadIdx.x; i < v_hi; i += 4ull * kThreads) {
        float4 v = *reinterpret_cast<const float4*>(t.g + i);
        if (t.kl_kind) {
            const float4 q = *reinterpret_cast<const float4*>(t.p + i);
            v.x += kl_extra(q.x, t.kl_kind, gs); v.y += kl_extra(q.y, t.kl_kind, gs);
            v.z += kl_extra(q.z, t.kl_kind, gs); v.w += kl_extra(q.w, t.kl_kind, gs);
        }
        s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    for (unsigned long long i = v_hi + threadIdx.x; i < hi; i += kThreads) {
        const float v = t.g[i] + (t.kl_kind ? kl_extra(t.p[i], t.kl_kind, gs) : 0.f);
        s += (double)v * v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < kThreads / 32 ? s_red[threadIdx.x] : 0.0;
        w = warp_sum(w);
        if (threadIdx.x == 0 && w != 0.0) atomicAdd(acc, w);
    }
}

__global__ void __launch_bounds__(kThreads) k_adamw_multi(const __grid_constant__ MtTable tab, AdamArgs a,
                                                         const double* __restrict__ gsumsq, int write_grad,
                                                         const float* __restrict__ kl_gscale) {
    const int k = mt_find(tab, blockIdx.x);
    const MtTensor& t = tab.t[k];
    a.kl_kind = t.kl_kind;
    a.kl_gs = t.kl_kind ? *kl_gscale : 0.f;
    float clip = 1.f;
    if (gsumsq && a.max_norm > 0.f) clip = fminf(1.f, a.max_norm / ((float)sqrt(*gsumsq) + 1e-6f));
    const unsigned long long lo = (unsigned long long)(blockIdx.x - t.first_block) * kMtChunk;
    const unsigned long long hi = lo + kMtChunk < t.n ? lo + kMtChunk : t.n;
    const unsigned long long v_hi = lo + ((hi - lo) & ~3ull);
#pragma unroll 2
    for (unsigned long long i = lo + 4ull * threadIdx.x; i < v_hi; i += 4ull * kThreads) {
        float4 pp = *reinterpret_cast<float4*>(t.p + i), gg = *reinterpret_cast<float4*>(t.g + i),
               mm = *reinterpret_cast<float4*>(t.m + i), vv = *reinterpret_cast<float4*>(t.v + i);
        gg.x = adam_one(pp.x, gg.x, mm.x, vv.x, a, clip);
        gg.y = adam_one(pp.y, gg.y, mm.y, vv.y, a, clip);
        gg.z = adam_one(pp.z, gg.z, mm.z, vv.z, a, clip);
        gg.w = adam_one(pp.w, gg.w, mm.w, vv.w, a, clip);
        *reinterpret_cast<float4*>(t.p + i) = pp;
        *reinterpret_cast<float4*>(t.m + i) = mm;
        *reinterpret_cast<float4*>(t.v + i) = vv;
        if (write_grad) *reinterpret_cast<float4*>(t.g + i) = gg;
    }
    for (unsigned long long i = v_hi + threadIdx.x; i < hi; i += kThreads) {
        float pp = t.p[i], mm = t.m[i], vv = t.v[i];
        const float gg = adam_one(pp, t.g[i], mm, vv, a, clip);
        t.p[i] = pp; t.m[i] = mm; t.v[i] = vv;
        if (write_grad) t.g[i] = gg;
    }
}

// fills `tab` from host arrays [first, first + count); returns the number of blocks
static unsigned mt_fill(MtTable& tab, int first, int count, float* const* params, float* const* grads, float* const* ms,
                        float* const* vs, const uint64_t* ns, const int* kl_kinds) {
    unsigned blocks = 0;
    tab.count = count;
    for (int k = 0; k < count; ++k) {
        MtTensor& t = tab.t[k];
        t.p = params ? params[first + k] : nullptr;
        t.g = grads[first + k];
        t.m = ms ? ms[first + k] : nullptr;
        t.v = vs ? vs[first + k] : nullptr;
        t.n = ns[first + k];
        t.kl_kind = kl_kinds ? kl_kinds[first + k] : 0;
        t.first_block = blocks;
        blocks += (unsigned)((t.n + kMtChunk - 1) / kMtChunk);
    }
    return blocks;
}

}  // namespace

extern "C" int pno_sumsq_accumulate_multi(int count, float* const* grads, const uint64_t* ns, float* const* params,
                                          const int* kl_kinds, const float* kl_gscale, double* acc, void* stream) {
    if (count <= 0) return PNO_OK;
    if (!grads || !ns || !acc) PNO_FAIL(PNO_E_INVALID, "pno_sumsq_accumulate_multi: null pointer");
    for (int k = 0; k < count; ++k) {
        const int kind = kl_kinds ? kl_kinds[k] : 0;
        if (ns[k] == 0 || !grads[k] || (reinterpret_cast<uintptr_t>(grads[k]) & 15) || kind < 0 || kind > 2 ||
            (kind && (!params || !params[k] || !kl_gscale || (reinterpret_cast<uintptr_t>(params[k]) & 15))))
            PNO_FAIL(PNO_E_INVALID, "pno_sumsq_accumulate_multi: tensor %d: empty, null or misaligned, or bad kl_kind", k);
    }
    for (int first = 0; first < count; first += kMtMax) {
        MtTable tab;
        const int c = count - first < kMtMax ? count - first : kMtMax;
        const unsigned blocks = mt_fill(tab, first, c, params, grads, nullptr, nullptr, ns, kl_kinds);
        k_sumsq_multi<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(tab, acc, kl_gscale);
        PNO_LAUNCH_CHECK();
    }
    return PNO_OK;
}

extern "C" int pno_adamw_step_multi(int count, float* const* params, float* const* grads, float* const* exp_avgs,
                                    float* const* exp_avg_sqs, const uint64_t* ns, const int* kl_kinds, float lr, float beta1,
                                    float beta2, float eps, float weight_decay, uint32_t step, const double* grad_sumsq,
                                    float max_norm, int write_clipped_grad, const float* kl_gscale, void* stream) {
    if (count <= 0) return PNO_OK;
    if (!params || !grads || !exp_avgs || !exp_avg_sqs || !ns || step == 0) PNO_FAIL(PNO_E_INVALID, "pno_adamw_step_multi: bad arguments");
    for (int k = 0; k < count; ++k) {
        const int kind = kl_kinds ? kl_kinds[k] : 0;
        if (ns[k] == 0 || !params[k] || !grads[k] || !exp_avgs[k] || !exp_avg_sqs[k] || kind < 0 || kind > 2 || (kind && !kl_gscale) ||
            ((reinterpret_cast<uintptr_t>(params[k]) | reinterpret_cast<uintptr_t>(grads[k]) | reinterpret_cast<uintptr_t>(exp_avgs[k]) |
              reinterpret_cast<uintptr_t>(exp_avg_sqs[k])) & 15))
            PNO_FAIL(PNO_E_INVALID, "pno_adamw_step_multi: tensor %d: empty, null or misaligned, or bad kl_kind", k);
    }
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    a.kl_kind = 0; a.kl_gs = 0.f;
    a.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    for (int first = 0; first < count; first += kMtMax) {
        MtTable tab;
        const int c = count - first < kMtMax ? count - first : kMtMax;
        const unsigned blocks = mt_fill(tab, first, c, params, grads, exp_avgs, exp_avg_sqs, ns, kl_kinds);
        k_adamw_multi<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(tab, a, grad_sumsq, write_clipped_grad, kl_gscale);
        PNO_LAUNCH_CHECK();
    }
    return PNO_OK;
}

extern "C" int pno_sumsq_accumulate(const float* g, uint64_t n, double* acc, const float* param, int kl_kind, const float* kl_gscale,
                                    void* stream) {
    if (n == 0) return PNO_OK;
    if (!g || !acc) PNO_FAIL(PNO_E_INVALID, "pno_sumsq_accumulate: null pointer");
    if (kl_kind < 0 || kl_kind > 2 || (kl_kind && (!param || !kl_gscale || (reinterpret_cast<uintptr_t>(param) & 15))))
        PNO_FAIL(PNO_E_INVALID, "pno_sumsq_accumulate: kl_kind in {0, 1, 2}; kinds 1 / 2 need the (16-byte aligned) parameter and the scale");
    if (reinterpret_cast<uintptr_t>(g) & 15) PNO_FAIL(PNO_E_INVALID, "pno_sumsq_accumulate: gradient must be 16-byte aligned");
    const unsigned long long want = (n / 4 + kThreads - 1) / kThreads + 1;
    const int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    k_sumsq<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g, n, acc, param, kl_kind, kl_gscale);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, uint64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, uint32_t step, const double* grad_sumsq, float max_norm,
                              int write_clipped_grad, int kl_kind, const float* kl_gscale, void* stream) {
    if (n == 0) return PNO_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || step == 0) PNO_FAIL(PNO_E_INVALID, "pno_adamw_step: bad arguments");
    if (kl_kind < 0 || kl_kind > 2 || (kl_kind && !kl_gscale)) PNO_FAIL(PNO_E_INVALID, "pno_adamw_step: kl_kind in {0, 1, 2}, kinds 1 / 2 need the scale");
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        PNO_FAIL(PNO_E_INVALID, "pno_adamw_step: tensors must be 16-byte aligned");
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    a.kl_kind = kl_kind; a.kl_gs = 0.f;
    a.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    const unsigned long long want = (n / 4 + kThreads - 1) / kThreads + 1;
    const int grid = (int)(want < 148ull * 16 ? want : 148ull * 16);
    k_adamw<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, a, grad_sumsq,
                                                                      write_clipped_grad, kl_gscale);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
