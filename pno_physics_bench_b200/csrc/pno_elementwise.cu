// Noise materialisation (parity/debug), activation backward, KL reduction and MC moment kernels.
#include "pno_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void k_normal_activation(PhiloxKey key_in, uint64_t first, uint64_t n, float* __restrict__ out) {
    const PhiloxKey key = resolve_key(key_in);
    // one thread per quad overlapping [first, first+n)
    const uint64_t q0 = first >> 2;
    const uint64_t nq = ((first + n + 3) >> 2) - q0;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nq; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t q = q0 + t;
        const float4 z = philox_quad_normals(key, q);
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const uint64_t e = q * 4 + l;
            if (e >= first && e < first + n) out[e - first] = zz[l];
        }
    }
}

__global__ void k_normal_spectral(PhiloxKey key_in, int Ci, int Co, int m1, int m2, float* __restrict__ ere,
                                  float* __restrict__ eim) {
    const PhiloxKey key = resolve_key(key_in);
    const int Cop = (Co + 1) / 2, m1m2 = m1 * m2;
    const uint64_t nq = (uint64_t)2 * m1m2 * Ci * Cop;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < nq; q += (uint64_t)gridDim.x * blockDim.x) {
        const int op = (int)(q % Cop);
        uint64_t r = q / Cop;
        const int i = (int)(r % Ci);
        r /= Ci;
        const int mb = (int)(r % m1m2);
        const int blk = (int)(r / m1m2);
        const float4 z = philox_quad_normals(key, q);
        // logical layout [blk][i][o][kx][ky] == [blk][i][o][mb]
        const int o = 2 * op;
        const size_t base = (((size_t)blk * Ci + i) * Co + o) * m1m2 + mb;
        ere[base] = z.x;
        eim[base] = z.y;
        if (o + 1 < Co) {
            ere[base + m1m2] = z.z;
            eim[base + m1m2] = z.w;
        }
    }
}

__global__ void k_bits(PhiloxKey key_in, uint64_t q0, uint64_t nq, uint32_t* __restrict__ out) {
    const PhiloxKey key = resolve_key(key_in);
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nq; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = philox_quad_bits(key, q0 + t);
        reinterpret_cast<uint4*>(out)[t] = x;
    }
}

__global__ void k_act_bwd(const float* __restrict__ z, const float* __restrict__ g, uint64_t n, int act,
                          float* __restrict__ gz) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        gz[i] = g[i] * act_grad(z[i], act);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// KL of one block: -0.5 * sum(1 + lv - re^2 - im^2 - exp(lv)); fp32 per-thread partials over a short strip, fp64 across
// threads/blocks (sums of ~1e8 terms: SURVEY.md A.3 row P4).
__global__ void __launch_bounds__(kThreads) k_kl_fwd(const float* __restrict__ mu, const float* __restrict__ lv,
                                                     uint64_t n, double* __restrict__ accum) {
    double acc = 0.0;
    const float2* mu2 = reinterpret_cast<const float2*>(mu);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 m = mu2[i];
        const float l = lv[i];
        acc += (double)(1.0f + l - m.x * m.x - m.y * m.y - expf(l));
    }
    acc = warp_sum(acc);
    __shared__ double part[kThreads / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(accum, -0.5 * v);
    }
}

__global__ void k_kl_bwd(const float* __restrict__ mu, const float* __restrict__ lv, uint64_t n,
                         const float* __restrict__ gscale, int accumulate, float* __restrict__ dmu,
                         float* __restrict__ dlv) {
    const float gs = *gscale;
    const float2* mu2 = reinterpret_cast<const float2*>(mu);
    float2* dmu2 = reinterpret_cast<float2*>(dmu);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 m = mu2[i];
        float2 d = make_float2(gs * m.x, gs * m.y);
        float dl = gs * (-0.5f * (1.0f - expf(lv[i])));
        if (accumulate) {
            const float2 o = dmu2[i];
            d.x += o.x;
            d.y += o.y;
            dl += dlv[i];
        }
        dmu2[i] = d;
        dlv[i] = dl;
    }
}

__global__ void k_mc_accumulate(const float* __restrict__ y, uint64_t n, double* __restrict__ s1, double* __restrict__ s2) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double v = (double)y[i];
        s1[i] += v;
        s2[i] += v * v;
    }
}

// uncertainty.py:57-73: prediction = mean + sqrt(exp(lv) + 1e-8) * eps with eps from this repo's Philox stream (element = flat
// index), accumulated with its square and the aleatoric variance exp(lv); optionally also written out (return_samples)
__global__ void k_mc_accumulate_decompose(const float* __restrict__ mean, const float* __restrict__ lv, uint64_t n, PhiloxKey key_in,
                                          double* __restrict__ s1, double* __restrict__ s2, double* __restrict__ sa,
                                          float* __restrict__ pred) {
    const PhiloxKey key = resolve_key(key_in);
    const uint64_t nq = (n + 3) >> 2;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < nq; q += (uint64_t)gridDim.x * blockDim.x) {
        const float4 z4 = lv ? philox_quad_normals(key, q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float z[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const uint64_t i = q * 4 + l;
            if (i < n) {
                const float var = lv ? expf(lv[i]) : 0.f;
                const float p = lv ? mean[i] + sqrtf(var + 1e-8f) * z[l] : mean[i];
                s1[i] += (double)p;
                s2[i] += (double)p * (double)p;
                sa[i] += (double)var;
                if (pred) pred[i] = p;
            }
        }
    }
}

__global__ void k_mc_finalize(const double* __restrict__ s1, const double* __restrict__ s2, uint64_t n, uint32_t S,
                              float* __restrict__ mean, float* __restrict__ sd) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double m = s1[i] / (double)S;
        mean[i] = (float)m;
        if (S > 1) {
            const double var = (s2[i] - (double)S * m * m) / (double)(S - 1);
            sd[i] = (float)sqrt(var > 0.0 ? var : 0.0);
        } else {
            sd[i] = __int_as_float(0x7fc00000);   // torch.std of one sample is NaN (models.py:368)
        }
    }
}

// device-resident noise key (pno_set_dynamic_noise): key[0..1] = seed, key[2] = sample_idx (+= advance when only_advance)
__global__ void k_noise_key(uint32_t* key, uint32_t k0, uint32_t k1, uint32_t sample, int only_advance) {
    if (only_advance) { key[2] += sample; return; }
    key[0] = k0; key[1] = k1; key[2] = sample; key[3] = 0u;
}

inline int grid_for(uint64_t n) {
    uint64_t g = (n + kThreads - 1) / kThreads;
    const uint64_t cap = 148ull * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int pno_philox_normal_activation(uint64_t seed, uint32_t sample_idx, uint32_t site, uint64_t first_elem,
                                            uint64_t n, float* out, void* stream) {
    if (n == 0) return PNO_OK;
    if (!out) PNO_FAIL(PNO_E_INVALID, "pno_philox_normal_activation: null output");
    k_normal_activation<<<grid_for((n + 3) / 4 + 1), kThreads, 0, as_stream(stream)>>>(make_key(seed, sample_idx, site),
                                                                                     first_elem, n, out);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_philox_normal_spectral(uint64_t seed, uint32_t sample_idx, uint32_t site, int Ci, int Co, int m1,
                                          int m2, float* eps_re, float* eps_im, void* stream) {
    if (!eps_re || !eps_im || Ci <= 0 || Co <= 0 || m1 <= 0 || m2 <= 0)
        PNO_FAIL(PNO_E_INVALID, "pno_philox_normal_spectral: bad arguments");
    const uint64_t nq = (uint64_t)2 * m1 * m2 * Ci * ((Co + 1) / 2);
    k_normal_spectral<<<grid_for(nq), kThreads, 0, as_stream(stream)>>>(make_key(seed, sample_idx, site), Ci, Co, m1, m2,
                                                                        eps_re, eps_im);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_noise_key_write(uint32_t* device_key, uint64_t seed, uint32_t sample_idx, void* stream) {
    if (!device_key) PNO_FAIL(PNO_E_INVALID, "pno_noise_key_write: null key buffer");
    k_noise_key<<<1, 1, 0, as_stream(stream)>>>(device_key, (uint32_t)seed, (uint32_t)(seed >> 32), sample_idx, 0);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_noise_key_advance(uint32_t* device_key, uint32_t delta, void* stream) {
    if (!device_key) PNO_FAIL(PNO_E_INVALID, "pno_noise_key_advance: null key buffer");
    k_noise_key<<<1, 1, 0, as_stream(stream)>>>(device_key, 0u, 0u, delta, 1);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_philox_bits(uint64_t seed, uint32_t sample_idx, uint32_t site, uint64_t q0, uint64_t nq, uint32_t* out,
                               void* stream) {
    if (nq == 0) return PNO_OK;
    if (!out) PNO_FAIL(PNO_E_INVALID, "pno_philox_bits: null output");
    k_bits<<<grid_for(nq), kThreads, 0, as_stream(stream)>>>(make_key(seed, sample_idx, site), q0, nq, out);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_act_bwd(const float* pre_act, const float* g, uint64_t n, int act, float* gz, void* stream) {
    if (n == 0) return PNO_OK;
    if (!pre_act || !g || !gz) PNO_FAIL(PNO_E_INVALID, "pno_act_bwd: null pointer");
    k_act_bwd<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(pre_act, g, n, act, gz);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_kl_fwd(const float* mu, const float* lv, uint64_t n, double* kl_accum, void* stream) {
    if (n == 0) return PNO_OK;
    if (!mu || !lv || !kl_accum) PNO_FAIL(PNO_E_INVALID, "pno_kl_fwd: null pointer");
    k_kl_fwd<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(mu, lv, n, kl_accum);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_kl_bwd(const float* mu, const float* lv, uint64_t n, const float* gscale, int accumulate, float* dmu,
                          float* dlv, void* stream) {
    if (n == 0) return PNO_OK;
    if (!mu || !lv || !gscale || !dmu || !dlv) PNO_FAIL(PNO_E_INVALID, "pno_kl_bwd: null pointer");
    k_kl_bwd<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(mu, lv, n, gscale, accumulate, dmu, dlv);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_mc_accumulate(const float* y, uint64_t n, double* sum, double* sumsq, void* stream) {
    if (n == 0) return PNO_OK;
    if (!y || !sum || !sumsq) PNO_FAIL(PNO_E_INVALID, "pno_mc_accumulate: null pointer");
    k_mc_accumulate<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(y, n, sum, sumsq);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_mc_accumulate_decompose(const float* mean, const float* log_var, uint64_t n, uint64_t seed, uint32_t sample_idx,
                                          uint32_t site, double* sum, double* sumsq, double* sum_avar, float* pred_out,
                                          void* stream) {
    if (n == 0) return PNO_OK;
    if (!mean || !sum || !sumsq || !sum_avar) PNO_FAIL(PNO_E_INVALID, "pno_mc_accumulate_decompose: null pointer");
    k_mc_accumulate_decompose<<<grid_for((n + 3) / 4), kThreads, 0, as_stream(stream)>>>(mean, log_var, n, make_key(seed, sample_idx, site),
                                                                                       sum, sumsq, sum_avar, pred_out);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

extern "C" int pno_mc_finalize(const double* sum, const double* sumsq, uint64_t n, uint32_t S, float* mean, float* std_,
                               void* stream) {
    if (n == 0) return PNO_OK;
    if (!sum || !sumsq || !mean || !std_ || S == 0) PNO_FAIL(PNO_E_INVALID, "pno_mc_finalize: bad arguments");
    k_mc_finalize<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(sum, sumsq, n, S, mean, std_);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
