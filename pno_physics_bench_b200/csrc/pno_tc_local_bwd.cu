// Backward of the local branch (the two 1x1 convs + reparameterisation, reference models.py:298-303 through autograd) on the
// tensor cores.  With  dYm = g  and  dYl = g * d out/d lv  (saved by the forward kernel):
//   dx[b]   = Wm^T dYm[b] + Wl^T dYl[b]                     -> ONE GEMM over the stacked gradient G2 = [dYm; dYl] ([B][2Co][HW])
//                                                              with the stacked transposed weights: the forward kernel
//                                                              (k_tc_local, no noise) runs it unchanged
//   dWm, dWl = sum_b dY{m,l}[b] x[b]^T                       -> k_tc_local_dw below: D[2Co][Ci] = G2 . x^T, reduction over the
//                                                              B*HW pixels split across the CTAs of the grid
//   dbm, dbl = row sums of G2                                -> k_local_bwd_prep (which also builds G2)
//
// k_tc_local_dw: work item = (128-row tile of G2, slice of the pixel range); both operands are K-major (pixels contiguous)
// straight from the NCHW tensors by TMA ([32 px][128 | Ci rows] SWIZZLE_128B boxes), accumulated in one [128 x Ci] TMEM tile
// over the whole slice, then added to the output with float atomics (at most ~74 contributions per element).
//   warp 0 TMA producer; warp 1 MMA issuer; warps 2-9 (3xTF32) hi/lo splitter; warps 10-13 epilogue.
#include <stdlib.h>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int TM = 128, TK = 32;
constexpr int A_BYTES = TM * TK * 4;          // 16 KB
constexpr int kSplitWarps = 8;
constexpr int kThreads = (2 + kSplitWarps + 4) * 32;

struct DwArgs {
    int Ci, rows, HW, B;          // rows = rows of G2 per batch item (Co or 2 Co)
    int m_tiles, k_splits;        // work items = m_tiles * k_splits
    int kb_total;                 // B * HW / 32 pixel blocks
    int stages, stage_bytes, b_bytes;
    float* dw;                    // [rows][Ci], zero-initialised by the caller
};

template <int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_local_dw(const DwArgs a, const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_raw[4], bar_split[4], bar_empty[4], bar_done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_split[s], kSplitWarps * 32);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_done, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const bool has_work = (int)blockIdx.x < a.m_tiles * a.k_splits;
    const int mt = blockIdx.x % a.m_tiles, ks = blockIdx.x / a.m_tiles;
    // pixel blocks [kb0, kb1) of this slice (contiguous in (b, p))
    const int kb0 = (int)((long long)a.kb_total * ks / a.k_splits), kb1 = (int)((long long)a.kb_total * (ks + 1) / a.k_splits);
    const int kb_per_img = a.HW / TK;

    if (has_work && warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&bar_empty[stage], phase ^ 1);
                unsigned char* st = smem + (size_t)stage * a.stage_bytes;
                const int b = kb / kb_per_img, p0 = (kb - b * kb_per_img) * TK;
                mbar_arrive_expect_tx(&bar_raw[stage], A_BYTES + a.b_bytes);
                tma_load_2d(st, &tmG, &bar_raw[stage], p0, b * a.rows + mt * TM);
                tma_load_2d(st + A_BYTES, &tmX, &bar_raw[stage], p0, b * a.Ci);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (has_work && warp == 1) {
        const uint32_t el = lane == 0;
        const uint32_t idesc = make_idesc(FMT_TF32, TM, a.Ci, MAJOR_K, MAJOR_K);
        const uint32_t dhi = sdesc_base(0, 16, 1024, SWIZZLE_128B).hi;
        const uint32_t lo = (A_BYTES + a.b_bytes) >> 4;          // hi tile -> its lo copy (3xTF32)
        int stage = 0, phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(NSPLIT == 3 ? &bar_split[stage] : &bar_raw[stage], phase);
            tc_fence_after_sync();
            const uint32_t s0 = smem_u32(smem + (size_t)stage * a.stage_bytes);
            const uint32_t ad = sdesc_base(s0, 16, 1024, SWIZZLE_128B).lo;
            const uint32_t bd = sdesc_base(s0 + A_BYTES, 16, 1024, SWIZZLE_128B).lo;
#pragma unroll
            for (int k = 0; k < TK / 8; ++k) {
                const uint32_t first = (kb != kb0 || k) ? 1u : 0u;
                mma_tf32(tmem, ad + 2 * k, dhi, bd + 2 * k, dhi, idesc, first, el);
                if (NSPLIT == 3) {
                    mma_tf32(tmem, ad + lo + 2 * k, dhi, bd + 2 * k, dhi, idesc, 1u, el);
                    mma_tf32(tmem, ad + 2 * k, dhi, bd + lo + 2 * k, dhi, idesc, 1u, el);
                }
            }
            mma_commit(&bar_empty[stage], el);
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&bar_done, el);
    } else if (has_work && warp < 2 + kSplitWarps) {
        if (NSPLIT == 3) {
            const int t = threadIdx.x - 64;
            const int n16 = (A_BYTES + a.b_bytes) / 16;
            int stage = 0, phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&bar_raw[stage], phase);
                const uint32_t hi = smem_u32(smem + (size_t)stage * a.stage_bytes), lw = hi + A_BYTES + a.b_bytes;
#pragma unroll 4
                for (int i = t; i < n16; i += kSplitWarps * 32) {
                    const float4 v = lds128(hi + 16 * i);
                    // the tensor core drops the low 13 mantissa bits of the raw tile (the hi operand): only lo = v - trunc(v) is stored
                    sts128(lw + 16 * i, tf32_lo_rn4(v));
                }
                fence_proxy_async_smem();
                mbar_arrive(&bar_split[stage]);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (has_work && warp >= 2 + kSplitWarps) {
        const int q = warp & 3;
        mbar_wait(&bar_done, 0);
        tc_fence_after_sync();
        float* dst = a.dw + (size_t)(mt * TM + q * 32 + lane) * a.Ci;
        for (int c = 0; c < a.Ci; c += 16) {
            float v[16];
            tmem_ld16(tmem_addr(tmem, q * 32, c), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + c + j, v[j]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

// G2[b][o][p] = g[b][o][p], G2[b][Co+o][p] = g[b][o][p] * dlv[b][o][p]; db[o] += row sums (dbm), db[Co+o] (dbl).
// One block per (b, o) row.  noisy == 0: only the bias sums (G2 is g itself).
__global__ void __launch_bounds__(256) k_local_bwd_prep(const float* __restrict__ g, const float* __restrict__ dlv, int Co, int HW,
                                                        int noisy, float* __restrict__ g2, float* __restrict__ db) {
    __shared__ float s_red[2][8];
    const int row = blockIdx.x;                 // b * Co + o
    const int b = row / Co, o = row - b * Co;
    const float4* gs = reinterpret_cast<const float4*>(g + (size_t)row * HW);
    const float4* ls = noisy ? reinterpret_cast<const float4*>(dlv + (size_t)row * HW) : nullptr;
    float4* d0 = noisy ? reinterpret_cast<float4*>(g2 + ((size_t)b * 2 * Co + o) * HW) : nullptr;
    float4* d1 = noisy ? reinterpret_cast<float4*>(g2 + ((size_t)b * 2 * Co + Co + o) * HW) : nullptr;
    float s0 = 0.f, s1 = 0.f;
    for (int i = threadIdx.x; i < HW / 4; i += 256) {
        const float4 v = gs[i];
        s0 += (v.x + v.y) + (v.z + v.w);
        if (noisy) {
            const float4 l = ls[i];
            const float4 w = make_float4(v.x * l.x, v.y * l.y, v.z * l.z, v.w * l.w);
            s1 += (w.x + w.y) + (w.z + w.w);
            d0[i] = v;
            d1[i] = w;
        }
    }
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, of);
        s1 += __shfl_xor_sync(0xffffffffu, s1, of);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = s0; s_red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x < 2 && db) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_red[threadIdx.x][w];
        if (threadIdx.x == 0) atomicAdd(db + o, s);
        else if (noisy) atomicAdd(db + Co + o, s);
    }
}

// Lift-shaped layers (Ci <= 4, e.g. the 3 -> C lift of models.py:288-290): the weight gradient is a handful of dot products per
// (b, o) row and the pass is a pure streaming read of g and d out / d lv.  One block per (b, o) row:
//   dw[o][i] += sum_p g x_i ; dw[Co + o][i] += sum_p g dlv x_i ; db[o] += sum_p g ; db[Co + o] += sum_p g dlv
__global__ void __launch_bounds__(256) k_local_thin_in_bwd(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ dlv, int Ci, int Co, int HW, int noisy,
                                                           float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float s_red[10][8];
    const int row = blockIdx.x;
    const int b = row / Co, o = row - b * Co;
    const float4* gs = reinterpret_cast<const float4*>(g + (size_t)row * HW);
    const float4* ls = noisy ? reinterpret_cast<const float4*>(dlv + (size_t)row * HW) : nullptr;
    float acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.f;
    for (int i4 = threadIdx.x; i4 < HW / 4; i4 += 256) {
        const float4 gv = gs[i4];
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (noisy) {
            const float4 l = ls[i4];
            wv = make_float4(gv.x * l.x, gv.y * l.y, gv.z * l.z, gv.w * l.w);
        }
        acc[8] += (gv.x + gv.y) + (gv.z + gv.w);
        acc[9] += (wv.x + wv.y) + (wv.z + wv.w);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < Ci) {
                const float4 xv = reinterpret_cast<const float4*>(x + ((size_t)b * Ci + i) * HW)[i4];
                acc[i] += gv.x * xv.x + gv.y * xv.y + gv.z * xv.z + gv.w * xv.w;
                acc[4 + i] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
            }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], of);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_red[threadIdx.x][w];
        const int k = threadIdx.x;
        if (k < 4) { if (k < Ci && dw) atomicAdd(dw + (size_t)o * Ci + k, s); }
        else if (k < 8) { if (k - 4 < Ci && dw && noisy) atomicAdd(dw + (size_t)(Co + o) * Ci + (k - 4), s); }
        else if (k == 8) { if (db) atomicAdd(db + o, s); }
        else if (db && noisy) atomicAdd(db + Co + o, s);
    }
}

// Projection-shaped layers (Co <= 4, e.g. the C -> 1 heads of models.py:309-314): one block per (b, i) row of x computes
//   dx[b][i][p] = sum_o Wm[o][i] g[b][o][p] + Wl[o][i] g[b][o][p] dlv[b][o][p]      (streaming write)
//   dw[o][i] += sum_p g x ;  dw[Co + o][i] += sum_p g dlv x                         (one pass over x);  db from the i == 0 rows
template <int CO>
__global__ void __launch_bounds__(256) k_local_thin_out_bwd(const float* __restrict__ x, const float* __restrict__ g,
                                                            const float* __restrict__ dlv, const float* __restrict__ wm,
                                                            const float* __restrict__ wl, int Ci, int HW, int noisy,
                                                            float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float s_red[4 * CO][8];
    const int row = blockIdx.x;
    const int b = row / Ci, i = row - b * Ci;
    float wmi[CO], wli[CO], acc[4 * CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        wmi[o] = wm[(size_t)o * Ci + i];
        wli[o] = noisy ? wl[(size_t)o * Ci + i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4 * CO; ++k) acc[k] = 0.f;
    const float4* xs = reinterpret_cast<const float4*>(x + (size_t)row * HW);
    float4* dxs = dx ? reinterpret_cast<float4*>(dx + (size_t)row * HW) : nullptr;
    for (int p4 = threadIdx.x; p4 < HW / 4; p4 += 256) {
        const float4 xv = xs[p4];
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < CO; ++o) {
            const float4 gv = reinterpret_cast<const float4*>(g + ((size_t)b * CO + o) * HW)[p4];
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (noisy) {
                const float4 l = reinterpret_cast<const float4*>(dlv + ((size_t)b * CO + o) * HW)[p4];
                wv = make_float4(gv.x * l.x, gv.y * l.y, gv.z * l.z, gv.w * l.w);
            }
            d.x += wmi[o] * gv.x + wli[o] * wv.x; d.y += wmi[o] * gv.y + wli[o] * wv.y;
            d.z += wmi[o] * gv.z + wli[o] * wv.z; d.w += wmi[o] * gv.w + wli[o] * wv.w;
            acc[o] += gv.x * xv.x + gv.y * xv.y + gv.z * xv.z + gv.w * xv.w;
            acc[CO + o] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
            acc[2 * CO + o] += (gv.x + gv.y) + (gv.z + gv.w);
            acc[3 * CO + o] += (wv.x + wv.y) + (wv.z + wv.w);
        }
        if (dxs) dxs[p4] = d;
    }
#pragma unroll
    for (int k = 0; k < 4 * CO; ++k) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], of);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 4 * CO) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_red[threadIdx.x][w];
        const int kind = threadIdx.x / CO, o = threadIdx.x % CO;
        if (kind == 0) { if (dw) atomicAdd(dw + (size_t)o * Ci + i, s); }
        else if (kind == 1) { if (dw && noisy) atomicAdd(dw + (size_t)(CO + o) * Ci + i, s); }
        else if (i == 0 && db) {                       // bias sums once per batch item
            if (kind == 2) atomicAdd(db + o, s);
            else if (noisy) atomicAdd(db + CO + o, s);
        }
    }
}

// wcat[i][o] = Wm[o][i], wcat[i][Co + o] = Wl[o][i]   ([Ci][2 Co]; [Ci][Co] when Wl == NULL)
__global__ void k_wcat(const float* __restrict__ wm, const float* __restrict__ wl, int Ci, int Co, float* __restrict__ wcat) {
    const int n = Ci * Co, ld = wl ? 2 * Co : Co;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int o = e / Ci, i = e - o * Ci;
        wcat[(size_t)i * ld + o] = wm[e];
        if (wl) wcat[(size_t)i * ld + Co + o] = wl[e];
    }
}

}  // namespace

size_t tc_local_bwd_workspace_bytes(int B, int Ci, int Co, int HW, int noisy) {
    if (Ci <= 4 || Co <= 4) return 1024;                   // lift-shaped layers stream g / dlv directly (no stacked gradient)
    const size_t g2 = noisy ? (size_t)B * 2 * Co * HW * 4 : 0;
    return g2 + (size_t)Ci * (noisy ? 2 : 1) * Co * 4 + 1024;
}

int tc_local_bwd(int engine, const float* x, const float* g, const float* dout_dlv, int B, int Ci, int Co, int HW, const float* Wm,
                 const float* Wl, float* dx, float* dw, float* db, void* workspace, cudaStream_t st) {
    const int noisy = Wl != nullptr;
    const int rows = noisy ? 2 * Co : Co;
    if (Ci <= 4 && dx == nullptr && HW % 4 == 0 && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) |
                                                      reinterpret_cast<uintptr_t>(dout_dlv)) & 15)) {
        // lift-shaped layer: streaming dot products (no input gradient: x is data)
        if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched
        if (noisy && !dout_dlv) PNO_FAIL(PNO_E_INVALID, "tc_local_bwd: d out / d log_var is needed when Wl is given");
        if (dw) PNO_CUDA(cudaMemsetAsync(dw, 0, (size_t)rows * Ci * 4, st));
        if (db) PNO_CUDA(cudaMemsetAsync(db, 0, (size_t)rows * 4, st));
        k_local_thin_in_bwd<<<B * Co, 256, 0, st>>>(x, g, dout_dlv, Ci, Co, HW, noisy, dw, db);
        PNO_LAUNCH_CHECK();
        return PNO_OK;
    }
    if (Co <= 4 && HW % 4 == 0 && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dout_dlv) |
                                     reinterpret_cast<uintptr_t>(dx)) & 15)) {
        // projection-shaped layer: one pass over x
        if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched
        if (noisy && !dout_dlv) PNO_FAIL(PNO_E_INVALID, "tc_local_bwd: d out / d log_var is needed when Wl is given");
        if (dw) PNO_CUDA(cudaMemsetAsync(dw, 0, (size_t)rows * Ci * 4, st));
        if (db) PNO_CUDA(cudaMemsetAsync(db, 0, (size_t)rows * 4, st));
#define PNO_PROJ_BWD(CO_) k_local_thin_out_bwd<CO_><<<B * Ci, 256, 0, st>>>(x, g, dout_dlv, Wm, Wl, Ci, HW, noisy, dx, dw, db)
        if (Co == 1) PNO_PROJ_BWD(1); else if (Co == 2) PNO_PROJ_BWD(2); else if (Co == 3) PNO_PROJ_BWD(3); else PNO_PROJ_BWD(4);
#undef PNO_PROJ_BWD
        PNO_LAUNCH_CHECK();
        return PNO_OK;
    }
    if (Co % TM || Ci % 128 || Ci > 256 || HW % 128 || rows % 32)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_local_bwd tiles Co%%128 == 0, Ci in {128, 256}, HW%%128 == 0 (got %d, %d, %d)", Co, Ci, HW);
    if (noisy && !dout_dlv) PNO_FAIL(PNO_E_INVALID, "tc_local_bwd: d out / d log_var is needed when Wl is given");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dout_dlv) |
         reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(workspace)) & 15)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_local_bwd needs 16-byte aligned tensors");
    if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched
    float* g2 = noisy ? static_cast<float*>(workspace) : const_cast<float*>(g);
    float* wcat = static_cast<float*>(workspace) + (noisy ? (size_t)B * 2 * Co * HW : 0);
    if (db) PNO_CUDA(cudaMemsetAsync(db, 0, (size_t)rows * 4, st));
    if (noisy || db) {
        k_local_bwd_prep<<<B * Co, 256, 0, st>>>(g, dout_dlv, Co, HW, noisy, g2, db);
        PNO_LAUNCH_CHECK();
    }
    int rc;
    if (dx) {        // dx = [Wm^T | Wl^T] G2: the forward kernel without noise, "input channels" = rows of G2
        k_wcat<<<(Ci * Co + 255) / 256, 256, 0, st>>>(Wm, Wl, Ci, Co, wcat);
        PNO_LAUNCH_CHECK();
        PhiloxKey key = make_key(0, 0, 0);
        if ((rc = tc_local_fwd(engine, g2, B, rows, Ci, HW, wcat, nullptr, nullptr, nullptr, key, 0, dx, nullptr, st))) return rc;
    }
    if (dw) {
        const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
        DwArgs a;
        memset(&a, 0, sizeof(a));
        a.Ci = Ci; a.rows = rows; a.HW = HW; a.B = B; a.dw = dw;
        a.m_tiles = rows / TM;
        int dev = 0, sms = 148;
        PNO_CUDA(cudaGetDevice(&dev));
        PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        a.kb_total = (int)((long long)B * HW / TK);
        a.k_splits = sms / a.m_tiles;
        // 3xTF32: eight times as many (shorter) pixel slices.  The tensor core's fp32 accumulator truncates on every MMA, so the
        // error of the accumulated sum grows with the number of MMAs chained into it (1 329 per slice of the cfg-2 reduction over
        // 131 072 pixels otherwise); the slices are combined by fp32 atomics (round to nearest).
        if (x3) a.k_splits *= 8;
        if (a.k_splits > a.kb_total) a.k_splits = a.kb_total;
        if (a.k_splits < 1) a.k_splits = 1;
        a.b_bytes = Ci * TK * 4;
        a.stage_bytes = (A_BYTES + a.b_bytes) * (x3 ? 2 : 1);
        a.stages = (200 * 1024) / a.stage_bytes;
        if (a.stages > 4) a.stages = 4;
        const int smem = a.stages * a.stage_bytes + 1024;
        CUtensorMap tmG, tmX;
        {
            const uint64_t dims[2] = {(uint64_t)HW, (uint64_t)B * rows};
            const uint64_t str[1] = {(uint64_t)HW * 4};
            const uint32_t box[2] = {TK, TM};
            if ((rc = pno_encode_tmap(&tmG, g2, 2, dims, str, box, 1))) return rc;
        }
        {
            const uint64_t dims[2] = {(uint64_t)HW, (uint64_t)B * Ci};
            const uint64_t str[1] = {(uint64_t)HW * 4};
            const uint32_t box[2] = {TK, (uint32_t)Ci};
            if ((rc = pno_encode_tmap(&tmX, x, 2, dims, str, box, 1))) return rc;
        }
        PNO_CUDA(cudaMemsetAsync(dw, 0, (size_t)rows * Ci * 4, st));
        const int grid = a.m_tiles * a.k_splits;
        if (x3) {
            PNO_CUDA(cudaFuncSetAttribute(k_tc_local_dw<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            PNO_CUDA(pno_launch(k_tc_local_dw<3>, grid, kThreads, smem, st, a, tmG, tmX));
        } else {
            PNO_CUDA(cudaFuncSetAttribute(k_tc_local_dw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            PNO_CUDA(pno_launch(k_tc_local_dw<1>, grid, kThreads, smem, st, a, tmG, tmX));
        }
        PNO_LAUNCH_CHECK();
    }
    return PNO_OK;
}
