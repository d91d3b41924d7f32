// Per-mode complex channel mix with in-kernel weight sampling, on CUDA cores (engine PNO_ENGINE_FP32).
//
// Replaces sample_weights() + the two einsum("bixy,ioxy->boxy") calls of the reference (models.py:63-78, 33-35, 93-94)
// and their autograd.  The sampled weights W = mu + exp(lv/2) * eps are produced tile by tile in shared memory from the
// Philox stream and never exist in HBM; the backward kernels re-draw the same eps from the same counters.
//
// Layouts: xhat/yhat/ghat [mode][b][c] (re / im planes); mu [kx][ky][i][o] complex interleaved, lv [kx][ky][i][o],
// one array per block (mode < m1*m2 -> block 0 = weights1).
#include "pno_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int T = 64;    // output tile edge
constexpr int KC = 16;   // reduction chunk
constexpr int LD = T + 4;

struct MixArgs {
    int B, Ci, Co, m1m2, Cop;
    const float *mu1, *mu2, *lv1, *lv2;
    PhiloxKey key;
};

// eps (re, im) and sigma = exp(lv/2) of the 4 consecutive weights (i, o0..o0+3) of `mode`; o0 is a multiple of 4.
__device__ __forceinline__ void load_eps4(const MixArgs& a, int mode, int i, int o0, float er[4], float ei[4],
                                          float sg[4]) {
    const int blk = mode >= a.m1m2;
    const int mb = mode - blk * a.m1m2;
    const float* lv = blk ? a.lv2 : a.lv1;
    const size_t base = ((size_t)mb * a.Ci + i) * a.Co + o0;
    float l[4];
    if ((o0 + 3 < a.Co) && ((a.Co & 3) == 0)) {
        const float4 q = *reinterpret_cast<const float4*>(lv + base);
        l[0] = q.x; l[1] = q.y; l[2] = q.z; l[3] = q.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) l[j] = (o0 + j < a.Co) ? lv[base + j] : 0.f;
    }
    const uint64_t q = spectral_quad(blk, a.m1m2, mb, a.Ci, a.Cop, i, o0 >> 1);
    const float4 za = philox_quad_normals(a.key, q);
    const float4 zb = philox_quad_normals(a.key, q + 1);
    er[0] = za.x; ei[0] = za.y; er[1] = za.z; ei[1] = za.w;
    er[2] = zb.x; ei[2] = zb.y; er[3] = zb.z; ei[3] = zb.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) sg[j] = (o0 + j < a.Co) ? __expf(0.5f * l[j]) : 0.f;
}

// Sample (or just load, when lv == NULL) the 4 consecutive weights (i, o0..o0+3) of `mode`.
__device__ __forceinline__ void load_weights4(const MixArgs& a, int mode, int i, int o0, float wr[4], float wi[4]) {
    const int blk = mode >= a.m1m2;
    const int mb = mode - blk * a.m1m2;
    const float* mu = blk ? a.mu2 : a.mu1;
    const float* lv = blk ? a.lv2 : a.lv1;
    const size_t base = ((size_t)mb * a.Ci + i) * a.Co + o0;
    if ((o0 + 3 < a.Co) && ((a.Co & 3) == 0)) {
        const float4 p0 = *reinterpret_cast<const float4*>(mu + 2 * base);
        const float4 p1 = *reinterpret_cast<const float4*>(mu + 2 * base + 4);
        wr[0] = p0.x; wi[0] = p0.y; wr[1] = p0.z; wi[1] = p0.w;
        wr[2] = p1.x; wi[2] = p1.y; wr[3] = p1.z; wi[3] = p1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = o0 + j < a.Co;
            wr[j] = ok ? mu[2 * (base + j)] : 0.f;
            wi[j] = ok ? mu[2 * (base + j) + 1] : 0.f;
        }
    }
    if (lv) {
        float er[4], ei[4], sg[4];
        load_eps4(a, mode, i, o0, er, ei, sg);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            wr[j] = fmaf(sg[j], er[j], wr[j]);
            wi[j] = fmaf(sg[j], ei[j], wi[j]);
        }
    }
}

// yhat[mode][b][o] = sum_i xhat[mode][b][i] * W[i][o].   grid (M, ceil(Co/T)); b tiles looped inside.
__global__ void __launch_bounds__(kThreads) k_mix_fwd(const MixArgs a_in, const float* __restrict__ xr,
                                                      const float* __restrict__ xi, float* __restrict__ yr,
                                                      float* __restrict__ yi) {
    MixArgs a = a_in;
    a.key = resolve_key(a_in.key);
    __shared__ __align__(16) float sxr[KC][LD], sxi[KC][LD], swr[KC][LD], swi[KC][LD];
    const int mode = blockIdx.x, o_base = blockIdx.y * T;
    const int tid = threadIdx.x, tb = tid >> 4, to = tid & 15;
    for (int b_base = 0; b_base < a.B; b_base += T) {
        float accr[4][4] = {}, acci[4][4] = {};
        for (int i0 = 0; i0 < a.Ci; i0 += KC) {
            {   // xhat tile [T b][KC i] -> sx[i][b]
                const int bl = tid >> 2, i4 = (tid & 3) * 4;
                const int b = b_base + bl;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + i4 + j;
                    const bool ok = b < a.B && i < a.Ci;
                    const size_t off = ((size_t)mode * a.B + b) * a.Ci + i;
                    sxr[i4 + j][bl] = ok ? xr[off] : 0.f;
                    sxi[i4 + j][bl] = ok ? xi[off] : 0.f;
                }
            }
            {   // weight tile [KC i][T o]
                const int il = tid >> 4, o4 = (tid & 15) * 4;
                const int i = i0 + il, o = o_base + o4;
                float wr[4] = {0, 0, 0, 0}, wi[4] = {0, 0, 0, 0};
                if (i < a.Ci && o < a.Co) load_weights4(a, mode, i, o, wr, wi);
                *reinterpret_cast<float4*>(&swr[il][o4]) = make_float4(wr[0], wr[1], wr[2], wr[3]);
                *reinterpret_cast<float4*>(&swi[il][o4]) = make_float4(wi[0], wi[1], wi[2], wi[3]);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const float4 ar = *reinterpret_cast<const float4*>(&sxr[k][tb * 4]);
                const float4 ai = *reinterpret_cast<const float4*>(&sxi[k][tb * 4]);
                const float4 wr = *reinterpret_cast<const float4*>(&swr[k][to * 4]);
                const float4 wi = *reinterpret_cast<const float4*>(&swi[k][to * 4]);
                const float xa[4] = {ar.x, ar.y, ar.z, ar.w}, xb[4] = {ai.x, ai.y, ai.z, ai.w};
                const float wa[4] = {wr.x, wr.y, wr.z, wr.w}, wb[4] = {wi.x, wi.y, wi.z, wi.w};
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        accr[p][q] = fmaf(xa[p], wa[q], accr[p][q]);
                        accr[p][q] = fmaf(-xb[p], wb[q], accr[p][q]);
                        acci[p][q] = fmaf(xa[p], wb[q], acci[p][q]);
                        acci[p][q] = fmaf(xb[p], wa[q], acci[p][q]);
                    }
            }
            __syncthreads();
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int b = b_base + tb * 4 + p;
            if (b >= a.B) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = o_base + to * 4 + q;
                if (o < a.Co) {
                    const size_t off = ((size_t)mode * a.B + b) * a.Co + o;
                    yr[off] = accr[p][q];
                    yi[off] = acci[p][q];
                }
            }
        }
    }
}

// dxhat[mode][b][i] = sum_o ghat[mode][b][o] * conj(W[i][o]).   grid (M, ceil(Ci/T))
__global__ void __launch_bounds__(kThreads) k_mix_bwd_dx(const MixArgs a_in, const float* __restrict__ gr,
                                                         const float* __restrict__ gi, float* __restrict__ dxr,
                                                         float* __restrict__ dxi) {
    MixArgs a = a_in;
    a.key = resolve_key(a_in.key);
    __shared__ __align__(16) float sgr[KC][LD], sgi[KC][LD], swr[KC][LD], swi[KC][LD];
    const int mode = blockIdx.x, i_base = blockIdx.y * T;
    const int tid = threadIdx.x, tb = tid >> 4, ti = tid & 15;
    for (int b_base = 0; b_base < a.B; b_base += T) {
        float accr[4][4] = {}, acci[4][4] = {};
        for (int o0 = 0; o0 < a.Co; o0 += KC) {
            {   // ghat tile [T b][KC o] -> sg[o][b]
                const int bl = tid >> 2, o4 = (tid & 3) * 4;
                const int b = b_base + bl;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = o0 + o4 + j;
                    const bool ok = b < a.B && o < a.Co;
                    const size_t off = ((size_t)mode * a.B + b) * a.Co + o;
                    sgr[o4 + j][bl] = ok ? gr[off] : 0.f;
                    sgi[o4 + j][bl] = ok ? gi[off] : 0.f;
                }
            }
            {   // weight tile [T i][KC o] -> sw[o][i]
                const int il = tid >> 2, o4 = (tid & 3) * 4;
                const int i = i_base + il, o = o0 + o4;
                float wr[4] = {0, 0, 0, 0}, wi[4] = {0, 0, 0, 0};
                if (i < a.Ci && o < a.Co) load_weights4(a, mode, i, o, wr, wi);
#pragma unroll
                for (int j = 0; j < 4; ++j) { swr[o4 + j][il] = wr[j]; swi[o4 + j][il] = wi[j]; }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const float4 ar = *reinterpret_cast<const float4*>(&sgr[k][tb * 4]);
                const float4 ai = *reinterpret_cast<const float4*>(&sgi[k][tb * 4]);
                const float4 wr = *reinterpret_cast<const float4*>(&swr[k][ti * 4]);
                const float4 wi = *reinterpret_cast<const float4*>(&swi[k][ti * 4]);
                const float ga[4] = {ar.x, ar.y, ar.z, ar.w}, gb[4] = {ai.x, ai.y, ai.z, ai.w};
                const float wa[4] = {wr.x, wr.y, wr.z, wr.w}, wb[4] = {wi.x, wi.y, wi.z, wi.w};
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {   // (ga + i gb)(wa - i wb)
                        accr[p][q] = fmaf(ga[p], wa[q], accr[p][q]);
                        accr[p][q] = fmaf(gb[p], wb[q], accr[p][q]);
                        acci[p][q] = fmaf(gb[p], wa[q], acci[p][q]);
                        acci[p][q] = fmaf(-ga[p], wb[q], acci[p][q]);
                    }
            }
            __syncthreads();
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int b = b_base + tb * 4 + p;
            if (b >= a.B) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i_base + ti * 4 + q;
                if (i < a.Ci) {
                    const size_t off = ((size_t)mode * a.B + b) * a.Ci + i;
                    dxr[off] = accr[p][q];
                    dxi[off] = acci[p][q];
                }
            }
        }
    }
}

// dmu[mode][i][o] = sum_b conj(xhat[mode][b][i]) ghat[mode][b][o];  dlv = 0.5 s (eps_re Re + eps_im Im)
// grid (M, ceil(Ci/T) * ceil(Co/T))
__global__ void __launch_bounds__(kThreads) k_mix_bwd_dw(const MixArgs a_in, const float* __restrict__ xr,
                                                         const float* __restrict__ xi, const float* __restrict__ gr,
                                                         const float* __restrict__ gi, float* __restrict__ dmu1,
                                                         float* __restrict__ dmu2, float* __restrict__ dlv1,
                                                         float* __restrict__ dlv2) {
    MixArgs a = a_in;
    a.key = resolve_key(a_in.key);
    __shared__ __align__(16) float sxr[KC][LD], sxi[KC][LD], sgr[KC][LD], sgi[KC][LD];
    const int mode = blockIdx.x;
    const int ntile_o = (a.Co + T - 1) / T;
    const int i_base = (blockIdx.y / ntile_o) * T, o_base = (blockIdx.y % ntile_o) * T;
    const int tid = threadIdx.x, ti = tid >> 4, to = tid & 15;
    float accr[4][4] = {}, acci[4][4] = {};
    for (int b0 = 0; b0 < a.B; b0 += KC) {
        {   // [KC b][T c] tiles, rows contiguous in c
            const int bl = tid >> 4, c4 = (tid & 15) * 4;
            const int b = b0 + bl;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i_base + c4 + j, o = o_base + c4 + j;
                const bool oki = b < a.B && i < a.Ci, oko = b < a.B && o < a.Co;
                const size_t offi = ((size_t)mode * a.B + b) * a.Ci + i;
                const size_t offo = ((size_t)mode * a.B + b) * a.Co + o;
                sxr[bl][c4 + j] = oki ? xr[offi] : 0.f;
                sxi[bl][c4 + j] = oki ? xi[offi] : 0.f;
                sgr[bl][c4 + j] = oko ? gr[offo] : 0.f;
                sgi[bl][c4 + j] = oko ? gi[offo] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float4 ar = *reinterpret_cast<const float4*>(&sxr[k][ti * 4]);
            const float4 ai = *reinterpret_cast<const float4*>(&sxi[k][ti * 4]);
            const float4 br = *reinterpret_cast<const float4*>(&sgr[k][to * 4]);
            const float4 bi = *reinterpret_cast<const float4*>(&sgi[k][to * 4]);
            const float xa[4] = {ar.x, ar.y, ar.z, ar.w}, xb[4] = {ai.x, ai.y, ai.z, ai.w};
            const float ga[4] = {br.x, br.y, br.z, br.w}, gb[4] = {bi.x, bi.y, bi.z, bi.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // (xa - i xb)(ga + i gb)
                    accr[p][q] = fmaf(xa[p], ga[q], accr[p][q]);
                    accr[p][q] = fmaf(xb[p], gb[q], accr[p][q]);
                    acci[p][q] = fmaf(xa[p], gb[q], acci[p][q]);
                    acci[p][q] = fmaf(-xb[p], ga[q], acci[p][q]);
                }
        }
        __syncthreads();
    }
    const int blk = mode >= a.m1m2;
    const int mb = mode - blk * a.m1m2;
    float* dmu = blk ? dmu2 : dmu1;
    float* dlv = blk ? dlv2 : dlv1;
    const bool sampled = (blk ? a.lv2 : a.lv1) != nullptr;
    const int o = o_base + to * 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = i_base + ti * 4 + p;
        if (i >= a.Ci || o >= a.Co) continue;
        const size_t base = ((size_t)mb * a.Ci + i) * a.Co + o;
        float er[4], ei[4], sg[4];
        if (sampled && dlv) load_eps4(a, mode, i, o, er, ei, sg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (o + q >= a.Co) break;
            if (dmu) {
                dmu[2 * (base + q)] = accr[p][q];
                dmu[2 * (base + q) + 1] = acci[p][q];
            }
            if (sampled && dlv) dlv[base + q] = 0.5f * sg[q] * fmaf(er[q], accr[p][q], ei[q] * acci[p][q]);
        }
    }
}

MixArgs make_args(const pno_plan* p, int B, int Ci, int Co, const float* mu1, const float* mu2, const float* lv1,
                  const float* lv2, PhiloxKey key) {
    MixArgs a;
    a.B = B; a.Ci = Ci; a.Co = Co; a.m1m2 = p->m1 * p->m2; a.Cop = (Co + 1) / 2;
    a.mu1 = mu1; a.mu2 = mu2; a.lv1 = lv1; a.lv2 = lv2; a.key = key;
    return a;
}

}  // namespace

int simt_mix_fwd(const pno_plan* p, const float* xr, const float* xi, int B, int Ci, int Co, const float* mu1,
                 const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* yr, float* yi,
                 cudaStream_t st) {
    const MixArgs a = make_args(p, B, Ci, Co, mu1, mu2, lv1, lv2, key);
    dim3 grid(p->M, (Co + T - 1) / T);
    k_mix_fwd<<<grid, kThreads, 0, st>>>(a, xr, xi, yr, yi);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

int simt_mix_bwd_dx(const pno_plan* p, const float* gr, const float* gi, int B, int Ci, int Co, const float* mu1,
                    const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* dxr, float* dxi,
                    cudaStream_t st) {
    const MixArgs a = make_args(p, B, Ci, Co, mu1, mu2, lv1, lv2, key);
    dim3 grid(p->M, (Ci + T - 1) / T);
    k_mix_bwd_dx<<<grid, kThreads, 0, st>>>(a, gr, gi, dxr, dxi);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

int simt_mix_bwd_dw(const pno_plan* p, const float* xr, const float* xi, const float* gr, const float* gi, int B, int Ci,
                    int Co, const float* lv1, const float* lv2, PhiloxKey key, float* dmu1, float* dmu2, float* dlv1,
                    float* dlv2, cudaStream_t st) {
    // mu is not needed for dW: the epilogue only re-draws eps and sigma (load_eps4)
    const MixArgs a = make_args(p, B, Ci, Co, nullptr, nullptr, lv1, lv2, key);
    dim3 grid(p->M, ((Ci + T - 1) / T) * ((Co + T - 1) / T));
    k_mix_bwd_dw<<<grid, kThreads, 0, st>>>(a, xr, xi, gr, gi, dmu1, dmu2, dlv1, dlv2);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
