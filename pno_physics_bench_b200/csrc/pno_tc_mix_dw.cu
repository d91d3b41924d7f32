// Weight gradient of the spectral channel mix on the tensor cores (autograd of reference models.py:33-35, 63-80, 93-94):
//   dmu[mode][i][o] = sum_b conj(xhat[mode][b][i]) ghat[mode][b][o]
//   dlv[mode][i][o] = 0.5 exp(lv/2) (eps_re Re(dmu) + eps_im Im(dmu))        (eps re-drawn from the forward's Philox counters)
// Work item = (mode, 128 input channels i, 128 output channels o).  The reduction runs over the batch (K = B): a few K steps, fed
// through the stage ring in chunks of KC batch rows (one chunk for the training batches of BASELINE config 2; several for the global
// batches of mode-parallel training, any B), so the kernel is all epilogue: both operands are MN-major tiles straight from the [mode][b][c] spectra by TMA (3-D maps: rows
// b >= B are zero-filled), the issuer needs 4 real MMAs per K step (a negated-A instruction descriptor gives the minus sign),
// accumulators (Re | Im, 2 x 128 columns) are double buffered in tensor memory, and 16 epilogue warps re-draw the noise, load
// log_var and write dmu (interleaved complex) and dlv as whole 32-byte sectors per lane.
//   warp 0 TMA producer; warp 1 MMA issuer; warps 2-5 (3xTF32) hi/lo splitter; warps 6-21 epilogue.
#include <stdlib.h>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int TM = 128;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (2 + 4 + kEpiWarps) * 32;

struct DwGeom {
    int B, Bp, Ci, Co, m1m2, Cop, n_modes;
    int mode_first, w_sub;          // mode window (see MixGeom in pno_tc_mix.cu)
    int i_tiles, o_tiles, n_items;
    int KC, n_kc;                   // batch rows per ring stage (multiple of 8), chunks per work item
    int tile_bytes, stage_bytes, stages;
    const float *lv1, *lv2;
    float *dmu1, *dmu2, *dlv1, *dlv2;
    PhiloxKey key;
};

__device__ __forceinline__ void ldg256(const float* p, float v[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float v[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

template <int NSPLIT, bool SAMPLED>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_mix_dw(const DwGeom g, const __grid_constant__ CUtensorMap tmXr, const __grid_constant__ CUtensorMap tmXi,
            const __grid_constant__ CUtensorMap tmGr, const __grid_constant__ CUtensorMap tmGi) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_full[4], bar_split[4], bar_empty[4], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_split[s], 128);
            mbar_init(&bar_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_tfull[s], 1);
            mbar_init(&bar_tempty[s], kEpiWarps * 32);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmXr);
        tma_prefetch_desc(&tmXi);
        tma_prefetch_desc(&tmGr);
        tma_prefetch_desc(&tmGi);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int T = g.tile_bytes;                      // one operand tile: [4 blocks of 32 channels][KC batch rows][128 B]
    const int per_mode = g.i_tiles * g.o_tiles;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
                const int mode = item / per_mode, r = item - mode * per_mode;
                const int it = r / g.o_tiles, ot = r - it * g.o_tiles;
                for (int kc = 0; kc < g.n_kc; ++kc) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    unsigned char* st = smem + (size_t)stage * g.stage_bytes;
                    mbar_arrive_expect_tx(&bar_full[stage], 4 * T);
                    const int b0 = kc * g.KC;                      // rows b >= B of the box are out of bounds: zero fill
                    for (int cb = 0; cb < 4; ++cb) {
                        tma_load_3d(st + 0 * T + cb * g.KC * 128, &tmXr, &bar_full[stage], it * TM + cb * 32, b0, mode);
                        tma_load_3d(st + 1 * T + cb * g.KC * 128, &tmXi, &bar_full[stage], it * TM + cb * 32, b0, mode);
                        tma_load_3d(st + 2 * T + cb * g.KC * 128, &tmGr, &bar_full[stage], ot * TM + cb * 32, b0, mode);
                        tma_load_3d(st + 3 * T + cb * g.KC * 128, &tmGi, &bar_full[stage], ot * TM + cb * 32, b0, mode);
                    }
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t el = lane == 0;
        const uint32_t id = make_idesc(FMT_TF32, TM, TM, MAJOR_MN, MAJOR_MN);
        const uint32_t idn = make_idesc(FMT_TF32, TM, TM, MAJOR_MN, MAJOR_MN, 1, 0);        // -A
        const uint32_t dhi = sdesc_base(0, (uint32_t)g.KC * 128, 512, SWIZZLE_128B_BASE32B).hi;
        const uint32_t t16 = T >> 4, lo = (4 * T) >> 4;
        int stage = 0, phase = 0, n = 0;
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++n) {
            const int acc = n & 1;
            mbar_wait(&bar_tempty[acc], ((n >> 1) & 1) ^ 1);
            const uint32_t d_re = tmem + acc * 256, d_im = d_re + 128;
          for (int kc = 0; kc < g.n_kc; ++kc) {
            mbar_wait(NSPLIT == 3 ? &bar_split[stage] : &bar_full[stage], phase);
            tc_fence_after_sync();
            const uint32_t xr = sdesc_base(smem_u32(smem + (size_t)stage * g.stage_bytes), (uint32_t)g.KC * 128, 512, SWIZZLE_128B_BASE32B).lo;
            const uint32_t xi = xr + t16, gr = xr + 2 * t16, gi = xr + 3 * t16;
            for (int ks = 0; ks < g.KC / 8; ++ks) {
                const uint32_t o = ks * 64;                 // 8 K rows x 128 B
                const uint32_t first = (kc | ks) ? 1u : 0u;
                // Re += Xre^T Gre + Xim^T Gim ; Im += Xre^T Gim - Xim^T Gre
                mma_tf32(d_re, xr + o, dhi, gr + o, dhi, id, first, el);
                mma_tf32(d_re, xi + o, dhi, gi + o, dhi, id, 1u, el);
                mma_tf32(d_im, xr + o, dhi, gi + o, dhi, id, first, el);
                mma_tf32(d_im, xi + o, dhi, gr + o, dhi, idn, 1u, el);
                if (NSPLIT == 3) {
                    mma_tf32(d_re, xr + lo + o, dhi, gr + o, dhi, id, 1u, el);
                    mma_tf32(d_re, xr + o, dhi, gr + lo + o, dhi, id, 1u, el);
                    mma_tf32(d_re, xi + lo + o, dhi, gi + o, dhi, id, 1u, el);
                    mma_tf32(d_re, xi + o, dhi, gi + lo + o, dhi, id, 1u, el);
                    mma_tf32(d_im, xr + lo + o, dhi, gi + o, dhi, id, 1u, el);
                    mma_tf32(d_im, xr + o, dhi, gi + lo + o, dhi, id, 1u, el);
                    mma_tf32(d_im, xi + lo + o, dhi, gr + o, dhi, idn, 1u, el);
                    mma_tf32(d_im, xi + o, dhi, gr + lo + o, dhi, idn, 1u, el);
                }
            }
            mma_commit(&bar_empty[stage], el);
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
          }
            mma_commit(&bar_tfull[acc], el);
        }
    } else if (warp < 6) {
        if (NSPLIT == 3) {
            const int t = threadIdx.x - 64;
            int stage = 0, phase = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x)
            for (int kc = 0; kc < g.n_kc; ++kc) {
                mbar_wait(&bar_full[stage], phase);
                const uint32_t hi = smem_u32(smem + (size_t)stage * g.stage_bytes), lw = hi + 4 * T;
#pragma unroll 4
                for (int i = t; i < 4 * T / 16; i += 128) {
                    const float4 v = lds128(hi + 16 * i);
                    // the tensor core drops the low 13 mantissa bits of the raw tile (the hi operand): only lo = v - trunc(v) is stored
                    sts128(lw + 16 * i, tf32_lo_rn4(v));
                }
                fence_proxy_async_smem();
                mbar_arrive(&bar_split[stage]);
                if (++stage == g.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue: lane = input channel i, 32 of the 128 output channels per warp ================
        const int q = warp & 3, cw = (warp - 6) >> 2;
        const PhiloxKey key = resolve_key(g.key);
        int n = 0;
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++n) {
            const int mode = item / per_mode, r = item - mode * per_mode;
            const int it = r / g.o_tiles, ot = r - it * g.o_tiles;
            const int gm = mode + g.mode_first;
            const int blk = gm >= g.m1m2, mb = gm - blk * g.m1m2;
            const int i = it * TM + q * 32 + lane;
            const int o0 = ot * TM + cw * 32;
            const size_t base = ((size_t)(mb - g.w_sub) * g.Ci + i) * g.Co + o0;
            float* dmu = (blk ? g.dmu2 : g.dmu1);
            float* dlv = (blk ? g.dlv2 : g.dlv1);
            const float* lv = blk ? g.lv2 : g.lv1;
            const uint64_t quad0 = spectral_quad(blk, g.m1m2, mb, g.Ci, g.Cop, i, o0 >> 1);
            const int acc = n & 1;
            mbar_wait(&bar_tfull[acc], (n >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t t_re = tmem_addr(tmem, q * 32, acc * 256 + cw * 32), t_im = t_re + 128;
#pragma unroll 1
            for (int c = 0; c < 32; c += 16) {
                float re[16], im[16];
                tmem_ld16(t_re + c, re);
                tmem_ld16(t_im + c, im);
                float l[16];
                if (SAMPLED) {
                    ldg256(lv + base + c, l);
                    ldg256(lv + base + c + 8, l + 8);
                }
                tmem_ld_wait();
                if (dmu) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float w[8] = {re[4 * k], im[4 * k], re[4 * k + 1], im[4 * k + 1], re[4 * k + 2], im[4 * k + 2], re[4 * k + 3],
                                            im[4 * k + 3]};
                        stg256(dmu + 2 * (base + c + 4 * k), w);
                    }
                }
                if (SAMPLED) {
                    float d[16];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {        // one Philox call: (eps_re, eps_im) of o and o + 1
                        const float4 z = philox_quad_normals(key, quad0 + (c >> 1) + k);
                        const float s0 = 0.5f * fast_ex2(0.72134752044448170f * l[2 * k]);
                        const float s1 = 0.5f * fast_ex2(0.72134752044448170f * l[2 * k + 1]);
                        d[2 * k] = s0 * fmaf(z.x, re[2 * k], z.y * im[2 * k]);
                        d[2 * k + 1] = s1 * fmaf(z.z, re[2 * k + 1], z.w * im[2 * k + 1]);
                    }
                    stg256(dlv + base + c, d);
                    stg256(dlv + base + c + 8, d + 8);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bar_tempty[acc]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

int tc_mix_bwd_dw(const pno_plan* p, int engine, const float* xr, const float* xi, const float* gr, const float* gi, int B, int Ci,
                  int Co, const float* lv1, const float* lv2, PhiloxKey key, float* dmu1, float* dmu2, float* dlv1, float* dlv2,
                  cudaStream_t st, int mode_first, int mode_count) {
    const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
    const bool sampled = lv1 != nullptr && dlv1 != nullptr;
    if (B < 1 || Ci % TM || Co % TM || !dmu1 || !dmu2)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_mix_bwd_dw tiles B >= 1, Ci%%128 == 0, Co%%128 == 0 and writes both dmu blocks (got %d, %d, %d)",
                 B, Ci, Co);
    if ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(xi) | reinterpret_cast<uintptr_t>(gr) |
         reinterpret_cast<uintptr_t>(gi) | reinterpret_cast<uintptr_t>(lv1) | reinterpret_cast<uintptr_t>(lv2) |
         reinterpret_cast<uintptr_t>(dmu1) | reinterpret_cast<uintptr_t>(dmu2) | reinterpret_cast<uintptr_t>(dlv1) |
         reinterpret_cast<uintptr_t>(dlv2)) & 31)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_mix_bwd_dw needs 32-byte aligned tensors");
    DwGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Bp = (B + 7) / 8 * 8; g.Ci = Ci; g.Co = Co; g.m1m2 = p->m1 * p->m2; g.Cop = (Co + 1) / 2; g.n_modes = p->M;
    if (mode_count >= 0) {
        const int blk0 = mode_first >= g.m1m2;
        if (mode_first < 0 || mode_count < 1 || mode_first + mode_count > (blk0 + 1) * g.m1m2)
            PNO_FAIL(PNO_E_INVALID, "tc_mix_bwd_dw: mode window [%d, %d) must lie inside one weight block of %d modes", mode_first,
                     mode_first + mode_count, g.m1m2);
        g.n_modes = mode_count;
        g.mode_first = mode_first;
        g.w_sub = mode_first - blk0 * g.m1m2;
    }
    g.i_tiles = Ci / TM; g.o_tiles = Co / TM; g.n_items = g.n_modes * g.i_tiles * g.o_tiles;
    // batch rows per ring stage: the whole batch while one stage holds it (B <= 32 with hi / lo copies, <= 48 single pass: the
    // training batches), else chunks small enough for a three-stage ring
    const int kc_one = x3 ? 32 : 48, kc_many = x3 ? 16 : 32;
    g.KC = g.Bp <= kc_one ? g.Bp : kc_many;
    g.n_kc = (g.Bp + g.KC - 1) / g.KC;
    g.tile_bytes = TM * g.KC * 4;
    g.stage_bytes = 4 * g.tile_bytes * (x3 ? 2 : 1);
    g.stages = (200 * 1024) / g.stage_bytes;
    if (g.stages > 3) g.stages = 3;
    if (g.stages < 1) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_mix_bwd_dw: %d bytes of shared memory per stage", g.stage_bytes);
    if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched
    g.lv1 = lv1; g.lv2 = lv2; g.dmu1 = dmu1; g.dmu2 = dmu2; g.dlv1 = dlv1; g.dlv2 = dlv2; g.key = key;
    const int smem = g.stages * g.stage_bytes + 1024;
    CUtensorMap tmXr, tmXi, tmGr, tmGi;
    int rc;
    {   // [mode][b][c]: box = 32 channels x Bp batch rows of one mode; rows b >= B are out of bounds -> zero fill
        const uint64_t dx[3] = {(uint64_t)Ci, (uint64_t)B, (uint64_t)g.n_modes}, dg[3] = {(uint64_t)Co, (uint64_t)B, (uint64_t)g.n_modes};
        const uint64_t sx[2] = {(uint64_t)Ci * 4, (uint64_t)B * Ci * 4}, sg[2] = {(uint64_t)Co * 4, (uint64_t)B * Co * 4};
        const uint32_t box[3] = {32, (uint32_t)g.KC, 1};
        if ((rc = pno_encode_tmap(&tmXr, xr, 3, dx, sx, box, 2))) return rc;
        if ((rc = pno_encode_tmap(&tmXi, xi, 3, dx, sx, box, 2))) return rc;
        if ((rc = pno_encode_tmap(&tmGr, gr, 3, dg, sg, box, 2))) return rc;
        if ((rc = pno_encode_tmap(&tmGi, gi, 3, dg, sg, box, 2))) return rc;
    }
    int dev = 0, sms = 148;
    PNO_CUDA(cudaGetDevice(&dev));
    PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms = pno_usable_sms(sms);
    const int grid = g.n_items < sms ? g.n_items : sms;
#define PNO_LAUNCH_DW(NS, SM)                                                                                      \
    do {                                                                                                           \
        PNO_CUDA(cudaFuncSetAttribute(k_tc_mix_dw<NS, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
        PNO_CUDA(pno_launch(k_tc_mix_dw<NS, SM>, grid, kThreads, smem, st, g, tmXr, tmXi, tmGr, tmGi));            \
    } while (0)
    if (x3 && sampled) PNO_LAUNCH_DW(3, true);
    else if (x3) PNO_LAUNCH_DW(3, false);
    else if (sampled) PNO_LAUNCH_DW(1, true);
    else PNO_LAUNCH_DW(1, false);
#undef PNO_LAUNCH_DW
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
