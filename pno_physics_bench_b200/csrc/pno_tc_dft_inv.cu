// Pruned inverse 2-D transform + block epilogue on the tensor cores (engines PNO_ENGINE_TC_*):
//   y[b][c][h][w] = act( sum_{kx2,ky} coef * Re(yhat[kx2*m2+ky][b][c] e^{+2 pi i (kx h/H + ky w/W)}) + addend )
// (reference models.py:92-97 + 306; adjoint != 0 gives the backward of the forward transform) as two chained GEMMs per
// group of G consecutive channels, all operands in K-major SWIZZLE_32B "slab" form so that K = 2*m1 and K = 2*m2 need no
// padding:
//   loaders     yhat[mode][b][c0..c0+G) (16/32-byte runs) -> [tf32 hi/lo] -> B1 slabs  Yre/Yim[(img,ky)][kx2]
//   stage 1     Zre[h][(img,ky)] = C Yre^T - S Yim^T ; Zim = S Yre^T + C Yim^T          (A = column twiddles, M = 128 lanes = h)
//   stage 2     y[(img,h)][w] = Z * TwB^T             TwB[w][(re|im,ky)] = c_ky/(HW) * (cos | -sin)(2 pi ky w / W)
//   epilogue    TMEM -> + addend -> pre_act -> activation -> y
// Two forms of stage 2:
//   * single pass (TS = true): the MMA reads Z straight from the stage-1 accumulators in tensor memory (tcgen05.mma with the A
//     operand in TMEM) -- no transposer, no Zt buffers, no barrier between the stages (the MMA pipe runs in issue order).  The
//     stage-1 table repeats its H rows over all 128 lanes, so each lane quarter holds a copy of Z; the MMA of image il writes its
//     own W accumulator columns and the epilogue warps of lane quarter q read the columns of image (32 q) / H.
//   * 3xTF32 (TS = false): Z must be split into hi + lo, so transposer warps move TMEM Z -> registers -> hi/lo -> Zt slabs
//     [(img,h)][(re|im, ky)] in shared memory and stage 2 is a shared x shared MMA.
// Output: when there is an addend (always on the block path) it is staged by TMA as swizzled [128 rows][128 B] boxes; the
// epilogue replaces it in place with y and warp 0 sends the finished 128-row tile (one contiguous block of y) out with a TMA
// store -- no global store instructions.  Otherwise each thread owns one image row and stores through a transpose scratch.
//   warp 0 addend loads + tile stores (TMA); warp 1 MMA issuer; warps 2-9 loaders (+ transposers); warps 10.. epilogue (8 or 16
//   warps: exact GELU makes it the issue-heavy role).  A warp can only read the TMEM lane quarter (warp % 4): with the
//   transposer, for H <= 64 the stage-1 accumulators live in quarters 0-1, so the four front warps of those quarters transpose
//   (one accumulator half each) and the other four load; for H = 128 warps 2-5 load and warps 6-9 transpose both halves.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int kFront = 64 + 256;   // idle + issuer, loaders + transposers
constexpr int kMaxThreads = kFront + 16 * 32;

struct InvGeom {
    int H, W, m1, m2, B, C;
    int K1, K2;      // 2*m1, 2*m2
    int S1, S2;      // slabs: K1/8, K2/8
    int G, IPT, TPG; // images per group / per 128-row tile, tiles per group
    int N1;          // G*m2
    int n_groups;
    int adjoint, act;
    int epi_warps, off_scratch;   // epilogue warps (8 or 16) and their transpose scratch (2 KB per warp)
    int ts;                       // stage 2 reads Z from tensor memory (single-pass mode)
    int n_lw;                     // loader warps (4, or all 8 front warps when stage 2 reads Z from tensor memory)
    int zs;                       // rows of B1 / columns of Z per image: m2, or m2 rounded up to 8 (tensor-memory operand mode)
    int rowsA;                    // rows of one slab of the stage-1 table image (H, or 128 with the H rows repeated)
    int d2_stride;                // tensor-memory columns of one stage-2 accumulator set (W, or IPT * W)
    int nb1, nzt, nd2;            // buffer counts: B1 operand sets, Zt operand sets, stage-2 accumulators
    int b1_buf;                   // bytes of one B1 operand set
    int add_stages, off_add, add_bytes;   // TMA-staged addend tiles (0 stages: the epilogue loads the addend with LDG)
    int ldg_t;                            // LDG addend: row-major loads transposed through the store scratch (PNO_INV_LDGT=0: lane-per-row loads)
    int tma_out;                          // the epilogue overwrites the staged addend tile with y and the tile leaves by TMA store
    int off_tabA, off_tabB, off_b1, off_zt, tabA_one, tabB_one, b1_one, b1_im, zt_one, zt_buf, smem_bytes;
    long long* dbg;   // optional per-CTA wait-cycle counters (pno_debug_set_buffer)
    int dbg_mma;      // diagnostics: the issuer also waits for its own batches (serialises the pipeline)
};

__device__ __forceinline__ void wait_t(uint64_t* bar, uint32_t parity, long long& acc) {
#ifdef PNO_TRACE
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
#else
    mbar_wait(bar, parity);
#endif
}

#define TRACE(role, ev) PNO_TRACE_EV(g.dbg, role, ev)

__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

// TS: stage 2 takes its A operand (Z) straight from the stage-1 accumulators in tensor memory (single-pass mode only): no
// transposer, no Zt buffers.  The stage-1 table repeats its H rows over all 128 lanes, so every lane quarter holds a copy of
// Z; the stage-2 MMA of image il of a tile writes its own W accumulator columns and the epilogue warps of lane quarter q read
// the columns of image (32 q) / H.
template <int NSPLIT, int G, int ACT, bool TS>
__global__ void __launch_bounds__(kMaxThreads, 1)
k_tc_dft_inv(const InvGeom g, const float* __restrict__ tabA_g, const float* __restrict__ tabB_g,
             const float* __restrict__ coef, const float* __restrict__ yr, const float* __restrict__ yi,
             const float* __restrict__ addend, float* __restrict__ y, float* __restrict__ pre_act,
             const __grid_constant__ CUtensorMap tmAdd, const __grid_constant__ CUtensorMap tmOut) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_b1_full[2], bar_b1_empty[2], bar_d1_full, bar_d1_empty;
    __shared__ __align__(8) uint64_t bar_zt_full[3], bar_zt_empty[3], bar_d2_full[3], bar_d2_empty[3];
    __shared__ __align__(8) uint64_t bar_add_full[4], bar_add_empty[4];
    __shared__ uint32_t tmem_slot;
    __shared__ uint32_t s_witem[512];     // loader work items: part << 16 | first kx2 << 8 | first ky (no divisions in the hot loop)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0;
    (void)w4; (void)w5;
    PNO_TRACE_DECL();
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_b1_full[s], g.n_lw * 32);
            mbar_init(&bar_b1_empty[s], 1);
        }
        mbar_init(&bar_d1_full, 1);
        mbar_init(&bar_d1_empty, 128);
        for (int s = 0; s < 3; ++s) {
            mbar_init(&bar_zt_full[s], 128);
            mbar_init(&bar_zt_empty[s], 1);
            mbar_init(&bar_d2_full[s], 1);
            mbar_init(&bar_d2_empty[s], g.epi_warps * 32);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&bar_add_full[s], 1);
            mbar_init(&bar_add_empty[s], g.epi_warps * 32);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    {
        const int kyb_n = g.m2 >> 2, kxb_n = g.K1 >> 3;
        for (int w = threadIdx.x; w < 2 * kxb_n * kyb_n; w += blockDim.x) {
            const int part = w / (kxb_n * kyb_n), r = w - part * kxb_n * kyb_n;
            const int kxb = r / kyb_n, kyb = r - kxb * kyb_n;
            s_witem[w] = (uint32_t)part << 16 | (uint32_t)(kxb * 8) << 8 | (uint32_t)(kyb * 4);
        }
    }
    // twiddle tables: the plan holds them as the exact shared-memory image (slab-swizzled), so this is a straight copy
    {
        // global images: tabA = (C hi, C lo, S hi, S lo), tabB = (hi, lo); shared: (C hi, [C lo], S hi, [S lo]) and (hi, [lo])
        constexpr int NTc = NSPLIT == 3 ? 2 : 1;
        const int nA = g.tabA_one / 16, nB = g.tabB_one / 16;               // float4 per table
        const float4* srcA = reinterpret_cast<const float4*>(tabA_g);
        float4* dstA = reinterpret_cast<float4*>(smem + g.off_tabA);
        for (int e = threadIdx.x; e < 2 * NTc * nA; e += blockDim.x) {
            const int tbl = e / nA, i = e - tbl * nA;                        // destination table index
            const int src_tbl = NTc == 2 ? tbl : 2 * tbl;                    // 1x: C hi, S hi = global tables 0 and 2
            dstA[e] = srcA[(size_t)src_tbl * nA + i];
        }
        const float4* srcB = reinterpret_cast<const float4*>(tabB_g);
        float4* dstB = reinterpret_cast<float4*>(smem + g.off_tabB);
        for (int e = threadIdx.x; e < NTc * nB; e += blockDim.x) dstB[e] = srcB[e];
        if (TS) {                                  // rows ky >= m2 of an image's B1 block are never written: zero them once
            float4* b1z = reinterpret_cast<float4*>(smem + g.off_b1);
            for (int e = threadIdx.x; e < g.nb1 * g.b1_buf / 16; e += blockDim.x) b1z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                    // spectrum and addend (previous kernels' outputs) are complete from here on
    const int d2_col = 2 * g.N1;
    // role of the front warps 2..9
    const bool small_h = g.H <= 64;
    const bool front = warp >= 2 && warp < 10;
    const bool is_trans = !TS && front && (small_h ? (warp & 3) < 2 : warp >= 6);
    const bool is_loader = front && !is_trans;
    const int lw = (small_h && !TS) ? (((warp - 2) >> 2) * 2 + ((warp & 3) - 2)) : warp - 2;     // loader warp index 0 .. n_lw-1
    const int part_lo = small_h ? (warp - 4) >> 2 : 0, part_hi = small_h ? part_lo + 1 : 2;   // accumulator halves of a transposer warp

    if (warp == 0) {
        // ================================ addend producer (TMA) ================================
        // The local-branch tile of a 128-row output tile is W/32 swizzled [128 rows x 128 B] boxes; staging it with TMA two or
        // three tiles ahead takes the addend off the epilogue's critical path (as LDG it was 26% of the epilogue's stall
        // samples) and off the L1 miss-tracking budget.
        // With tma_out the epilogue writes y over the addend tile in place and this thread sends the finished tile out with a
        // TMA store (one contiguous 128-row block of y) before it refills the buffer: no global store instructions at all.
        if (lane == 0 && g.add_stages > 0 && addend != nullptr) {
            int n = 0;
            auto tile_row0 = [&](int t) {                 // first row of this CTA's t-th tile in the [B*C*H][W] view
                const int grp = blockIdx.x + (t / g.TPG) * gridDim.x, j = t % g.TPG;
                return (grp * G + j * g.IPT) * g.H;
            };
            auto store_tile = [&](int t) {                // tile t is complete in its buffer: send it out
                const int sb = t % g.add_stages;
                mbar_wait(&bar_add_empty[sb], (t / g.add_stages) & 1);
                TRACE(5, 1);
                const unsigned char* src = smem + g.off_add + sb * g.add_bytes;
                const int r0 = tile_row0(t);
                for (int cb = 0; cb < g.W / 32; ++cb) tma_store_2d(&tmOut, src + cb * 128 * 128, cb * 32, r0);
                bulk_commit_group();
            };
            for (int grp = blockIdx.x; grp < g.n_groups; grp += gridDim.x) {
                for (int j = 0; j < g.TPG; ++j, ++n) {
                    const int sa = n % g.add_stages;
                    if (g.tma_out) {
                        if (n >= g.add_stages) {
                            store_tile(n - g.add_stages);
                            bulk_wait_read0();            // the store has read the buffer: it may be refilled
                            TRACE(5, 2);
                        }
                    } else {
                        mbar_wait(&bar_add_empty[sa], ((n / g.add_stages) & 1) ^ 1);
                    }
                    unsigned char* dst = smem + g.off_add + sa * g.add_bytes;
                    const int row0 = (grp * G + j * g.IPT) * g.H;
                    mbar_arrive_expect_tx(&bar_add_full[sa], g.add_bytes);
                    for (int cb = 0; cb < g.W / 32; ++cb) tma_load_2d(dst + cb * 128 * 128, &tmAdd, &bar_add_full[sa], cb * 32, row0);
                    TRACE(5, 3);
                }
            }
            if (g.tma_out) {
                for (int t = n > g.add_stages ? n - g.add_stages : 0; t < n; ++t) store_tile(t);
                bulk_wait0();
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        // (a second issuer warp for stage 2 was measured: no gain in single-pass mode, slower in 3xTF32 mode)
        {   // all 32 lanes run the loop (uniform registers); `el` predicates the tcgen05 instructions
            const uint32_t el = lane == 0;
            const uint32_t id1 = make_idesc(FMT_TF32, 128, g.N1, MAJOR_K, MAJOR_K);
            const uint32_t id1n = make_idesc(FMT_TF32, 128, g.N1, MAJOR_K, MAJOR_K, 1, 0);
            const uint32_t id2 = make_idesc(FMT_TF32, 128, g.W, MAJOR_K, MAJOR_K);
            // descriptor words: every operand is K-major SW32 slabs (LBO 16, SBO 256); only the low word changes per MMA
            const uint32_t dhi = sdesc_base(0, 16, 256, SWIZZLE_32B).hi;
            const uint32_t tA = sdesc_base(smem_u32(smem + g.off_tabA), 16, 256, SWIZZLE_32B).lo;
            const uint32_t tB = sdesc_base(smem_u32(smem + g.off_tabB), 16, 256, SWIZZLE_32B).lo;
            const uint32_t b1 = sdesc_base(smem_u32(smem + g.off_b1), 16, 256, SWIZZLE_32B).lo;
            const uint32_t zt = sdesc_base(smem_u32(smem + g.off_zt), 16, 256, SWIZZLE_32B).lo;
            const uint32_t slabA = g.rowsA * 2, slabB1 = g.N1 * 2, slabZ = 128 * 2, slabTB = g.W * 2;     // slab strides in 16-byte units
            const uint32_t tA1 = g.tabA_one >> 4, b1o = g.b1_one >> 4, b1i = g.b1_im >> 4;
            constexpr uint32_t NTc = NSPLIT == 3 ? 2 : 1;      // table copies: S hi sits NTc tables after C hi
            int ng = 0, nt = 0;
            for (int grp = blockIdx.x; grp < g.n_groups; grp += gridDim.x, ++ng) {
                {
                const int bb = ng % g.nb1;
                wait_t(&bar_b1_full[bb], (ng / g.nb1) & 1, w0);
                TRACE(0, 1);
                if (!TS) wait_t(&bar_d1_empty, (ng & 1) ^ 1, w1);   // TS: stage 2 of the previous group is ahead of us in the MMA pipe
                TRACE(0, 2);
                tc_fence_after_sync();
                const uint32_t d_re = tmem, d_im = tmem + g.N1;
                uint32_t a = tA, b = b1 + bb * (g.b1_buf >> 4);
                for (int s = 0; s < g.S1; ++s, a += slabA, b += slabB1) {
                    const uint32_t first = s ? 1u : 0u;
                    // (C hi, [C lo], S hi, [S lo]) at a + {0, tA1, NTc*tA1, (NTc+1)*tA1};  (Yre hi, [Yre lo]) at b + {0, b1o}, (Yim ..) at b + b1i + ..
                    const uint32_t sa = a + NTc * tA1;
                    mma_tf32(d_re, a, dhi, b, dhi, id1, first, el);
                    mma_tf32(d_re, sa, dhi, b + b1i, dhi, id1n, 1u, el);
                    mma_tf32(d_im, sa, dhi, b, dhi, id1, first, el);
                    mma_tf32(d_im, a, dhi, b + b1i, dhi, id1, 1u, el);
                    if (NSPLIT == 3) {
                        mma_tf32(d_re, a + tA1, dhi, b, dhi, id1, 1u, el);
                        mma_tf32(d_re, a, dhi, b + b1o, dhi, id1, 1u, el);
                        mma_tf32(d_re, sa + tA1, dhi, b + b1i, dhi, id1n, 1u, el);
                        mma_tf32(d_re, sa, dhi, b + b1i + b1o, dhi, id1n, 1u, el);
                        mma_tf32(d_im, sa + tA1, dhi, b, dhi, id1, 1u, el);
                        mma_tf32(d_im, sa, dhi, b + b1o, dhi, id1, 1u, el);
                        mma_tf32(d_im, a + tA1, dhi, b + b1i, dhi, id1, 1u, el);
                        mma_tf32(d_im, a, dhi, b + b1i + b1o, dhi, id1, 1u, el);
                    }
                }
                mma_commit(&bar_b1_empty[bb], el);
                if (!TS) mma_commit(&bar_d1_full, el);
                TRACE(0, 3);
#ifdef PNO_TRACE
                if (g.dbg_mma) {                         // diagnostics: issue -> completion latency of the stage-1 batch
                    const long long tq = clock64();
                    mbar_wait(&bar_d1_full, ng & 1);
                    w4 += clock64() - tq;
                }
#endif
                }
                if (TS) {
                    // stage 2 from tensor memory: the MMA pipe executes in issue order, so these follow stage 1 without a
                    // round trip through a barrier, and stage 1 of the next group cannot overtake them
                    const int zs8 = g.zs >> 3;
                    for (int j = 0; j < g.TPG; ++j, ++nt) {
                        const int db = nt % g.nd2;
                        wait_t(&bar_d2_empty[db], ((nt / g.nd2) & 1) ^ 1, w3);
                        TRACE(0, 5);
                        tc_fence_after_sync();
                        for (int il = 0; il < g.IPT; ++il) {
                            const uint32_t d2 = tmem + d2_col + db * g.d2_stride + il * g.W;
                            uint32_t zc = tmem + (j * g.IPT + il) * g.zs, w = tB;
                            for (int s = 0; s < zs8; ++s, zc += 8, w += slabTB)
                                mma_tf32_ts(d2, zc, w, dhi, id2, s ? 1u : 0u, el);
                            zc = tmem + g.N1 + (j * g.IPT + il) * g.zs;
                            for (int s = 0; s < zs8; ++s, zc += 8, w += slabTB) mma_tf32_ts(d2, zc, w, dhi, id2, 1u, el);
                        }
                        mma_commit(&bar_d2_full[db], el);
                        TRACE(0, 6);
                    }
                } else
                for (int j = 0; j < g.TPG; ++j, ++nt) {
                    const int zb = nt % g.nzt, db = nt % g.nd2;
                    wait_t(&bar_zt_full[zb], (nt / g.nzt) & 1, w2);
                    TRACE(0, 4);
                    wait_t(&bar_d2_empty[db], ((nt / g.nd2) & 1) ^ 1, w3);
                    TRACE(0, 5);
                    tc_fence_after_sync();
                    const uint32_t d2 = tmem + d2_col + db * g.d2_stride;
                    uint32_t z = zt + zb * (g.zt_buf >> 4), w = tB;
                    for (int s = 0; s < g.S2; ++s, z += slabZ, w += slabTB) {
                        mma_tf32(d2, z, dhi, w, dhi, id2, s ? 1u : 0u, el);
                        if (NSPLIT == 3) {
                            mma_tf32(d2, z + (g.zt_one >> 4), dhi, w, dhi, id2, 1u, el);
                            mma_tf32(d2, z, dhi, w + (g.tabB_one >> 4), dhi, id2, 1u, el);
                        }
                    }
                    mma_commit(&bar_zt_empty[zb], el);
                    mma_commit(&bar_d2_full[db], el);
                    TRACE(0, 6);
#ifdef PNO_TRACE
                    if (g.dbg_mma) {
                        const long long tq = clock64();
                        mbar_wait(&bar_d2_full[db], (nt / g.nd2) & 1);
                        w5 += clock64() - tq;
                    }
#endif
                }
            }
        }
    } else if (is_loader) {
        // ================================ loaders: yhat -> B1 slabs ================================
        // Warp item = (re/im plane, block of 8 kx2, block of 4 ky); lane = (ky & 3, kx2 & 7).  A lane gathers the G channels
        // of its mode (one 16/32-byte run) and the warp writes, per channel, 4 slab rows x 32 B = 128 contiguous bytes of B1:
        // conflict-free, and the index arithmetic is warp-uniform except for two lane terms.
        constexpr int NLW = TS ? 8 : 4;                        // loader warps (= g.n_lw)
        const int kyb_n = g.m2 >> 2, kxb_n = g.K1 >> 3;
        const int witems = 2 * kxb_n * kyb_n;
        const int lkx = lane & 7, lky = lane >> 3;
        const int q4 = g.zs >> 2;                              // rows i*zs + ky: ((row >> 2) & 1) = (i*q4 + (ky >> 2)) & 1
        unsigned char* b1 = smem + g.off_b1;
        int ng = 0;
        auto decode = [&](int w, int& part, int& kx2, int& ky) {
            const uint32_t e = s_witem[w];
            part = e >> 16;
            kx2 = ((e >> 8) & 0xFF) + lkx;
            ky = (e & 0xFF) + lky;
        };
        for (int grp = blockIdx.x; grp < g.n_groups; grp += gridDim.x, ++ng) {
            const int img0 = grp * g.G;
            const int b = img0 / g.C, c0 = img0 % g.C;
            // the spectrum rows of the group AFTER this one are pulled into L2 while this loader would otherwise idle, so
            // that its 32-byte gathers are L2 hits instead of DRAM round trips
            if (grp + (int)gridDim.x < g.n_groups) {
                const int imgn = (grp + gridDim.x) * g.G;
                const size_t on = (size_t)(imgn / g.C) * g.C + imgn % g.C;
                for (int w = lw; w < witems; w += NLW) {
                    int part, kx2, ky;
                    decode(w, part, kx2, ky);
                    prefetch_l2((part ? yi : yr) + (size_t)(kx2 * g.m2 + ky) * g.B * g.C + on);
                }
            }
            const int bb = ng % g.nb1;
            bool waited = false;          // the first round of gathers is in flight while the B1 set is still being consumed
            const size_t o_grp = (size_t)b * g.C + c0;
            const uint32_t dbase = smem_u32(b1) + bb * g.b1_buf;
            constexpr int U = G == 8 ? 4 : 7;          // warp items in flight (register budget): loads first, then the scatter
            for (int w0i = lw; w0i < witems; w0i += NLW * U) {
                float4 va[U], vb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int w = w0i + NLW * u;
                    if (w < witems) {
                        int part, kx2, ky;
                        decode(w, part, kx2, ky);
                        const float* src = (part ? yi : yr) + (size_t)(kx2 * g.m2 + ky) * g.B * g.C + o_grp;
                        if (G == 8) {          // one 256-bit load per 32-byte run: every lane hits its own line, so each load
                                               // instruction costs the load / store unit 32 passes whatever its width
                            asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=f"(va[u].x), "=f"(va[u].y), "=f"(va[u].z), "=f"(va[u].w), "=f"(vb[u].x), "=f"(vb[u].y),
                                           "=f"(vb[u].z), "=f"(vb[u].w)
                                         : "l"(src));
                        } else {
                            va[u] = __ldg(reinterpret_cast<const float4*>(src));
                        }
                    }
                }
                if (!waited) {
                    wait_t(&bar_b1_empty[bb], ((ng / g.nb1) & 1) ^ 1, w0);
                    if (warp == 2) TRACE(3, 1);
                    waited = true;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int w = w0i + NLW * u;
                    if (w < witems) {
                        int part, kx2, ky;
                        decode(w, part, kx2, ky);
                        const float v[8] = {va[u].x, va[u].y, va[u].z, va[u].w, vb[u].x, vb[u].y, vb[u].z, vb[u].w};
                        // kslab_sw32_off(i*m2 + ky, kx2, N1) with the i-dependent terms separated
                        const uint32_t base = dbase + (part ? g.b1_im : 0) + (uint32_t)(kx2 >> 3) * g.N1 * 32 + ky * 32 + ((kx2 & 3) << 2);
                        const uint32_t sw0 = ((kx2 >> 2) ^ (ky >> 2)) & 1;
#pragma unroll
                        for (int i = 0; i < G; ++i) {
                            const uint32_t off = base + i * g.zs * 32 + (((sw0 ^ (i * q4)) & 1) << 4);
                            if (NSPLIT == 3) {
                                const float h = tf32_rn(v[i]);
                                sts32(off, h);
                                sts32(off + g.b1_one, v[i] - h);
                            } else {
                                sts32(off, v[i]);
                            }
                        }
                    }
                }
            }
            if (!waited) wait_t(&bar_b1_empty[bb], ((ng / g.nb1) & 1) ^ 1, w0);
            if (warp == 2) TRACE(3, 2);
            fence_proxy_async_smem();
            mbar_arrive(&bar_b1_full[bb]);
            if (warp == 2) TRACE(3, 3);
        }
    } else if (is_trans) {
        // ================================ transposer ================================
        // thread = stage-1 lane (row h) x accumulator halves [part_lo, part_hi) (re / im)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool t_active_warp = q * 32 < g.H;          // warp holds valid stage-1 lanes (h < H)
        int ng = 0, nT = 0;

        auto transpose_tile = [&](int j) {
            const int zb = nT % g.nzt;
            wait_t(&bar_zt_empty[zb], ((nT / g.nzt) & 1) ^ 1, w0);
            if (warp == 8) TRACE(1, 2);
            if (t_active_warp) {
                const uint32_t zt = smem_u32(smem + g.off_zt) + zb * g.zt_buf;
                for (int il = 0; il < g.IPT; ++il) {
                    const int img = j * g.IPT + il;                  // image within the group
                    const int zrow = il * g.H + row;                 // row of the stage-2 A tile (row < H here)
                    const uint32_t rbase = zrow * 32, rsw = (zrow >> 2) & 1;
                    for (int part = part_lo; part < part_hi; ++part)
                    for (int kyb = 0; kyb < g.m2; kyb += 32) {          // up to 8 x4 loads in flight, one wait
                        uint32_t r[8][4];
                        tmem_ld_cols(tmem_addr(tmem, q * 32, part * g.N1 + img * g.m2 + kyb), g.m2 - kyb, r);   // x16 / x8 / x4
                        tmem_ld_wait();
                        if (warp == 8) TRACE(1, 5);
                        if (row < g.H) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const int ky0 = kyb + 4 * c;
                                if (ky0 < g.m2) {
                                    const float4 v = make_float4(__uint_as_float(r[c][0]), __uint_as_float(r[c][1]),
                                                                 __uint_as_float(r[c][2]), __uint_as_float(r[c][3]));
                                    const int k = part * g.m2 + ky0;          // multiple of 4: one 16-byte chunk of a slab row
                                    const uint32_t off = (uint32_t)(k >> 3) * (128 * 32) + rbase + ((((k >> 2) & 1) ^ rsw) << 4);
                                    if (NSPLIT == 3) {
                                        const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                                        sts128(zt + off, h);
                                        sts128(zt + g.zt_one + off, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
                                    } else {
                                        sts128(zt + off, v);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            if (warp == 8) TRACE(1, 3);
            tc_fence_before_sync();
            fence_proxy_async_smem();
            mbar_arrive(&bar_zt_full[zb]);
            if (warp == 8) TRACE(1, 4);
            ++nT;
        };

        for (int grp = blockIdx.x; grp < g.n_groups; grp += gridDim.x, ++ng) {
            wait_t(&bar_d1_full, ng & 1, w1);
            if (warp == 8) TRACE(1, 1);
            tc_fence_after_sync();
            for (int j = 0; j < g.TPG; ++j) transpose_tile(j);
            mbar_arrive(&bar_d1_empty);                   // every stage-1 accumulator of this group has been read
        }
    } else if (warp >= 10) {
        // ================================ epilogue ================================
        // epi_warps / 4 warps per TMEM lane quarter, each taking an equal share of the W columns in 16-column passes
        const int ew = warp - 10;
        const int q = warp & 3, chunk = ew >> 2, n_chunks = g.epi_warps >> 2;
        const int row = q * 32 + lane;
        float4* sc = reinterpret_cast<float4*>(smem + g.off_scratch) + ew * 128;
        const int c_lo = chunk * (g.W / n_chunks), c_hi = c_lo + g.W / n_chunks;
        const int il_col = TS ? (q * 32 / g.H) * g.W : 0;      // TS: image (32 q) / H of the tile has its own accumulator columns
        int nE = 0;
        int db = 0, dph = 0, sa = 0, aph = 0;            // accumulator / staged-addend ring positions and phases (no divisions)
        for (int grp = blockIdx.x; grp < g.n_groups; grp += gridDim.x) {
            const int img0 = grp * G;
            for (int j = 0; j < g.TPG; ++j, ++nE) {
                const size_t tile_row0 = (size_t)(img0 + j * g.IPT) * g.H;     // rows of the [B*C*H][W] view are consecutive
                const size_t gofs = (tile_row0 + row) * g.W;
                const size_t wofs = (tile_row0 + q * 32) * g.W;              // first row of this warp
                const uint32_t ta = tmem_addr(tmem, q * 32, d2_col + db * g.d2_stride + il_col);
                // addend: staged by TMA (add_stages > 0) or fetched with LDG before the accumulator is ready
                const bool staged = g.add_stages > 0 && addend != nullptr;
                const uint32_t add_s = smem_u32(smem + g.off_add) + sa * g.add_bytes + row * 128;
                float4 a4[4];
                auto fetch = [&](int c) {
                    if (staged) {                  // swizzled [W/32][128 rows][128 B]: 16-byte chunk index XOR (row & 7)
                        const uint32_t blk = add_s + (c >> 5) * (128 * 128);
                        const int j0 = (c & 31) >> 2;
#pragma unroll
                        for (int k = 0; k < 4; ++k) a4[k] = lds128(blk + (((j0 + k) ^ (row & 7)) << 4));
                    } else if (addend && g.ldg_t) {
                        // row-major loads (8 rows x 64 B per instruction = 8 lines instead of 32: a lane-per-row load costs the
                        // load / store unit 32 passes) turned into lane-per-row through this warp's store scratch: 256 -> 250 us
                        // (fp32-parity engine, headline shape).  Staggering the store bursts of the four column chunks the way
                        // the forward transform paces its epilogue does NOT pay here (100 / 250 / 500 ns: 256 / 265 / 292 us): this
                        // epilogue is on the tile's critical path.
                        const float* src = addend + wofs + c + (size_t)(lane >> 2) * g.W + 4 * (lane & 3);
                        const uint32_t s0 = smem_u32(sc);
                        float4 t[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) t[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)8 * i * g.W));
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = 8 * i + (lane >> 2), ch = lane & 3;
                            sts128(s0 + 16u * (rr * 4 + (ch ^ (rr & 3))), t[i]);
                        }
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 4; ++k) a4[k] = lds128(s0 + 16u * (lane * 4 + (k ^ (lane & 3))));
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            a4[k] = addend ? __ldg(reinterpret_cast<const float4*>(addend + gofs + c + 4 * k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
                // A staged addend is waited for AFTER the accumulator: the TMA that fills it completes in many partial
                // transactions and each one wakes every warp parked on the barrier (~11 retries per wait, 12% of all executed
                // instructions of the kernel); by the time stage 2 of the tile is done the tile landed long ago.
                if (!staged) fetch(c_lo);
                // ... and (LDG path) the chunk of the tile after this one is pulled into L2 now, so that its fetch is an L2 hit
                if (addend && !staged) {
                    const size_t nxt_row0 = j + 1 < g.TPG ? tile_row0 + (size_t)g.IPT * g.H
                                                           : (size_t)(grp + gridDim.x) * G * g.H;
                    if (nxt_row0 + 128 <= (size_t)g.B * g.C * g.H) prefetch_l2(addend + (nxt_row0 + row) * g.W + c_lo);
                }
                wait_t(&bar_d2_full[db], dph, w2);
                if (ew == 0) TRACE(2, 1);
                if (ew == g.epi_warps - 1) TRACE(4, 1);
                tc_fence_after_sync();
                if (staged) {
                    mbar_wait(&bar_add_full[sa], aph);
                    if (ew == 0) TRACE(2, 3);
                    fetch(c_lo);
                }
#pragma unroll 1
                for (int c = c_lo; c < c_hi; c += 16) {
                    float v[16];
#ifdef PNO_TRACE
                    long long tq = clock64();
#endif
                    tmem_ld16(ta + c, v);
                    if (c != c_lo) fetch(c);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[4 * k] += a4[k].x; v[4 * k + 1] += a4[k].y; v[4 * k + 2] += a4[k].z; v[4 * k + 3] += a4[k].w;
                    }
                    // the staged tile has been consumed by this thread (the adds above depend on the loaded values)
                    if (staged && !g.tma_out && c + 16 >= c_hi) mbar_arrive(&bar_add_empty[sa]);
#ifdef PNO_TRACE
                    { const long long t2 = clock64(); w0 += t2 - tq; tq = t2; }
#endif
                    if (pre_act) warp_store_rows16(sc, lane, v, pre_act + wofs + c, g.W);
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = act_apply_fast(v[k], ACT);
#ifdef PNO_TRACE
                    { const long long t2 = clock64(); w1 += t2 - tq; tq = t2; }
#endif
                    if (g.tma_out) {               // y replaces the addend in the staged tile (same swizzled chunks this thread read)
                        const uint32_t blk = add_s + (c >> 5) * (128 * 128);
                        const int j0 = (c & 31) >> 2;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            sts128(blk + (((j0 + k) ^ (row & 7)) << 4), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                    } else {
                        warp_store_rows16(sc, lane, v, y + wofs + c, g.W);
                    }
#ifdef PNO_TRACE
                    { const long long t2 = clock64(); w3 += t2 - tq; tq = t2; }
#endif
                }
                tc_fence_before_sync();
                mbar_arrive(&bar_d2_empty[db]);
                if (g.tma_out) {                   // this thread's part of the output tile is in shared memory
                    fence_proxy_async_smem();
                    mbar_arrive(&bar_add_empty[sa]);
                }
                if (ew == 0) TRACE(2, 2);
                if (ew == g.epi_warps - 1) TRACE(4, 2);
                if (++db == g.nd2) { db = 0; dph ^= 1; }
                if (staged && ++sa == g.add_stages) { sa = 0; aph ^= 1; }
            }
        }
    }
#ifdef PNO_TRACE
    if (g.dbg && lane == 0 && (warp == 1 || warp == 8 || warp == 10)) {
        long long* d = g.dbg + ((size_t)blockIdx.x * 3 + (warp == 1 ? 0 : warp == 8 ? 1 : 2)) * 7;
        d[0] = w0; d[1] = w1; d[2] = w2; d[3] = w3; d[4] = clock64() - trace_t0; d[5] = w4; d[6] = w5;
    }
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

inline float tf32_round_h(double v) {
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&f, &u, 4);
    return f;
}

int pad_up(int v, int m) { return (v + m - 1) / m * m; }

struct InvTables {
    float* tabA;      // shared-memory image: (C hi, C lo, S hi, S lo), each S1 slabs of [H rows x 32 B], dead kx2 columns zeroed
    float* tabB;      // shared-memory image: (hi, lo), each S2 slabs of [W rows x 32 B]; entries carry c_ky / (H W)
    float* tabB_adj;  // same without the scale (adjoint != 0)
    // tensor-memory operand mode (single pass): hi parts only; the stage-1 table repeats its H rows over 128 rows, the stage-2
    // table has K = 2 * zs rows per w (zs = m2 rounded up to 8, rows ky >= m2 zero)
    float* tabA_ts;   // (C, unused, S), each S1 slabs of [128 rows x 32 B]
    float* tabB_ts;
    float* tabB_ts_adj;
};

}  // namespace

// Tables are cached on the plan through a second opaque slot owned by this translation unit.
static int inv_tables_get(const pno_plan* cp, InvTables** out) {
    pno_plan* p = const_cast<pno_plan*>(cp);
    if (p->tc_tables_inv) { *out = static_cast<InvTables*>(p->tc_tables_inv); return PNO_OK; }
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2, K1 = 2 * m1, K2 = 2 * m2;
    const double two_pi = 6.283185307179586476925286766559;
    const size_t oneA = (size_t)(K1 / 8) * H * 8, oneB = (size_t)(K2 / 8) * W * 8;       // floats per table image
    std::vector<float> a(4 * oneA, 0.f), bt(2 * oneB, 0.f), bta(2 * oneB, 0.f);
    for (int h = 0; h < H; ++h)
        for (int r = 0; r < K1; ++r) {
            const int kx = r < m1 ? r : H - 2 * m1 + r;
            const bool alive = r >= m1 || kx < H - m1;          // models.py:93-94 overwrite
            const double ang = two_pi * (double)((long long)kx * h % H) / H;
            const double c = alive ? cos(ang) : 0.0, s = alive ? sin(ang) : 0.0;
            const float ch = tf32_round_h(c), sh = tf32_round_h(s);
            const size_t e = sm100::kslab_sw32_off(h, r, H) / 4;
            a[0 * oneA + e] = ch; a[1 * oneA + e] = tf32_round_h(c - (double)ch);
            a[2 * oneA + e] = sh; a[3 * oneA + e] = tf32_round_h(s - (double)sh);
        }
    for (int w = 0; w < W; ++w)
        for (int k = 0; k < K2; ++k) {
            const int ky = k % m2;
            const bool self_conj = ky == 0 || (W % 2 == 0 && ky == W / 2);
            const double scale = (self_conj ? 1.0 : 2.0) / ((double)H * W);      // c_ky / (H W): the irfft2 weights
            const double ang = two_pi * (double)((long long)ky * w % W) / W;
            const double v = k < m2 ? cos(ang) : -sin(ang);
            const size_t e = sm100::kslab_sw32_off(w, k, W) / 4;
            const float hi = tf32_round_h(v * scale), hia = tf32_round_h(v);
            bt[e] = hi; bt[oneB + e] = tf32_round_h(v * scale - (double)hi);
            bta[e] = hia; bta[oneB + e] = tf32_round_h(v - (double)hia);
        }
    const int zs = (m2 + 7) / 8 * 8;
    const size_t oneAt = (size_t)(K1 / 8) * 128 * 8, oneBt = (size_t)(2 * zs / 8) * W * 8;
    std::vector<float> at(3 * oneAt, 0.f), btt(oneBt, 0.f), btta(oneBt, 0.f);
    for (int row = 0; row < 128; ++row)
        for (int r = 0; r < K1; ++r) {
            const size_t src = sm100::kslab_sw32_off(row % H, r, H) / 4, dst = sm100::kslab_sw32_off(row, r, 128) / 4;
            at[dst] = a[src];
            at[2 * oneAt + dst] = a[2 * oneA + src];
        }
    for (int w = 0; w < W; ++w)
        for (int k = 0; k < K2; ++k) {
            const size_t src = sm100::kslab_sw32_off(w, k, W) / 4;
            const size_t dst = sm100::kslab_sw32_off(w, (k / m2) * zs + k % m2, W) / 4;
            btt[dst] = bt[src];
            btta[dst] = bta[src];
        }
    InvTables* t = new InvTables();
    PNO_CUDA(cudaMalloc(&t->tabA_ts, at.size() * 4));
    PNO_CUDA(cudaMalloc(&t->tabB_ts, btt.size() * 4));
    PNO_CUDA(cudaMalloc(&t->tabB_ts_adj, btta.size() * 4));
    PNO_CUDA(cudaMemcpy(t->tabA_ts, at.data(), at.size() * 4, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMemcpy(t->tabB_ts, btt.data(), btt.size() * 4, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMemcpy(t->tabB_ts_adj, btta.data(), btta.size() * 4, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMalloc(&t->tabA, a.size() * 4));
    PNO_CUDA(cudaMalloc(&t->tabB, bt.size() * 4));
    PNO_CUDA(cudaMalloc(&t->tabB_adj, bta.size() * 4));
    PNO_CUDA(cudaMemcpy(t->tabA, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMemcpy(t->tabB, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice));
    PNO_CUDA(cudaMemcpy(t->tabB_adj, bta.data(), bta.size() * 4, cudaMemcpyHostToDevice));
    p->tc_tables_inv = t;
    *out = t;
    return PNO_OK;
}

void tc_plan_release_inv(pno_plan* p) {
    if (!p->tc_tables_inv) return;
    InvTables* t = static_cast<InvTables*>(p->tc_tables_inv);
    cudaFree(t->tabA);
    cudaFree(t->tabB);
    cudaFree(t->tabB_adj);
    cudaFree(t->tabA_ts);
    cudaFree(t->tabB_ts);
    cudaFree(t->tabB_ts_adj);
    delete t;
    p->tc_tables_inv = nullptr;
}

int tc_dft_inv_fused(const pno_plan* p, int engine, const float* yr, const float* yi, int B, int C, int adjoint,
                     const float* addend, int act, float* y, float* pre_act, cudaStream_t st) {
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2;
    if (!(H == 32 || H == 64 || H == 128) || W % 32 || W > 256 || W < 32 || m1 % 4 || m2 % 4 || m2 > 64)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_inv tiles H in {32,64,128}, W%%32 == 0, W <= 256, m1%%4 == 0, m2%%4 == 0");
    const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
    InvGeom g;
    memset(&g, 0, sizeof(g));
    g.H = H; g.W = W; g.m1 = m1; g.m2 = m2; g.B = B; g.C = C; g.adjoint = adjoint; g.act = act;
    g.K1 = 2 * m1; g.K2 = 2 * m2; g.S1 = g.K1 / 8; g.S2 = g.K2 / 8;
    g.IPT = 128 / H;
    const int nbuf = x3 ? 2 : 1;          // hi (+ lo)
    bool ok = false;
    const bool spec32 = !((reinterpret_cast<uintptr_t>(yr) | reinterpret_cast<uintptr_t>(yi)) & 31);      // 32-byte aligned spectra
    const char* e_epi = getenv("PNO_INV_EPI");         // tuning override: 8 = narrow epilogue
    const int epi_max = (W % 64 == 0 && !(e_epi && atoi(e_epi) == 8)) ? 16 : 8;
    const int add_bytes = 128 * W * 4;                 // one 128-row tile of the addend
    const char* e_add = getenv("PNO_INV_ADD");         // tuning override: 0 = LDG addend
    const int add_max = addend == nullptr ? 0 : (e_add ? atoi(e_add) : 3);
    const char* e_cap = getenv("PNO_INV_CAP");         // tuning override: shared-memory budget (KB) of the tensor-memory operand mode
    const int cap_kb = e_cap ? atoi(e_cap) : 200;
    const char* e_nd2 = getenv("PNO_INV_ND2");
    const char* e_nb1 = getenv("PNO_INV_NB1");
    const char* e_nzt = getenv("PNO_INV_NZT");
    const char* e_ts = getenv("PNO_INV_TS");           // tuning override: 0 = transposer path in single-pass mode too
    const char* e_to = getenv("PNO_INV_TMAOUT");       // tuning override: 0 = global store instructions instead of the TMA tile store
    // y can leave through the staged addend tile when there is an addend to stage and no second output
    const bool tma_out_ok = addend != nullptr && pre_act == nullptr && !(e_to && atoi(e_to) == 0);
    // ---- single pass: stage 2 reads Z from tensor memory (no transposer) ----
    if (!x3 && !(e_ts && atoi(e_ts) == 0)) {
        const int zs = pad_up(m2, 8);
        int best = -1;
        for (int G = 8; G >= g.IPT && G >= 4; G >>= 1) {
            const int N1 = G * zs;
            if (C % G || N1 % 16 || N1 > 256) continue;
            if (G == 8 && !spec32) continue;                 // the loaders fetch 8-channel runs with one 256-bit load
            if (getenv("PNO_INV_G") && atoi(getenv("PNO_INV_G")) != G) continue;      // tuning override
            int nd2 = (512 - 2 * N1) / (g.IPT * W);
            if (nd2 < 1) continue;
            if (nd2 > 3) nd2 = 3;
            if (e_nd2 && atoi(e_nd2) >= 1 && atoi(e_nd2) < nd2) nd2 = atoi(e_nd2);
            const int tabA_one = g.S1 * 128 * 32, tabB_one = (2 * zs / 8) * W * 32, b1_one = g.S1 * N1 * 32;
            const int b1_buf = pad_up(2 * b1_one, 1024);
            const int off_tabB = pad_up(2 * tabA_one, 1024), off_b1 = off_tabB + pad_up(tabB_one, 1024);
            const int nb1_max = e_nb1 ? (atoi(e_nb1) == 2 ? 2 : 1) : 2;
            for (int epi = epi_max; epi >= 8; epi -= 8)
                for (int nb1 = nb1_max; nb1 >= 1; --nb1) {
                    const int off_scratch = off_b1 + nb1 * b1_buf;
                    for (int na = add_max; na >= 0; na = (na == 0 ? -1 : (na > 2 ? na - 1 : 0))) {
                        const bool to = tma_out_ok && na > 0;                   // no store scratch needed then
                        const int off_add = pad_up(off_scratch + (to ? 0 : epi * 2048), 1024);
                        const int bytes = off_add + na * add_bytes + 1024;
                        if (bytes > cap_kb * 1024) continue;   // keep an L1 carve-out for the spectrum gathers (see below)
                        const int score = (nd2 >= 2 ? 10000 : 0) + epi * 100 + (na > 0 ? 50 : 0) + G * 4 + nb1 * 2 + na;
                        if (score > best) {
                            best = score; ok = true;
                            g.ts = 1; g.zs = zs; g.G = G; g.N1 = N1; g.TPG = G / g.IPT; g.nd2 = nd2; g.n_lw = 8; g.rowsA = 128;
                            g.d2_stride = g.IPT * W;
                            g.tabA_one = tabA_one; g.tabB_one = tabB_one; g.b1_one = b1_one; g.b1_im = b1_one; g.b1_buf = b1_buf;
                            g.off_tabA = 0; g.off_tabB = off_tabB; g.off_b1 = off_b1; g.nb1 = nb1; g.nzt = 0; g.off_zt = 0;
                            g.zt_one = 0; g.zt_buf = 0;
                            g.epi_warps = epi; g.off_scratch = off_scratch; g.off_add = off_add; g.add_stages = na;
                            g.add_bytes = add_bytes; g.smem_bytes = bytes; g.tma_out = to;
                        }
                    }
                }
        }
    }
    // ---- transposer path (3xTF32, or no tensor-memory geometry fits) ----
    for (int G = 8; G >= g.IPT && G >= 4 && !ok; G >>= 1) {
        if (C % G || (G * m2) % 16 || G * m2 > 256 || 2 * G * m2 + 2 * W > 512) continue;
        if (G == 8 && !spec32) continue;                     // the loaders fetch 8-channel runs with one 256-bit load
        if (getenv("PNO_INV_G") && atoi(getenv("PNO_INV_G")) != G) continue;      // tuning override
        g.ts = 0; g.zs = m2; g.n_lw = 4; g.rowsA = H; g.d2_stride = W;
        g.G = G; g.N1 = G * m2; g.TPG = G / g.IPT;
        g.tabA_one = g.S1 * H * 32;
        g.tabB_one = g.S2 * W * 32;
        g.b1_one = g.S1 * g.N1 * 32;
        g.b1_im = nbuf * g.b1_one;              // B1 layout: [Yre hi][Yre lo]? [Yim hi][Yim lo]?
        g.b1_buf = pad_up(2 * nbuf * g.b1_one, 1024);
        g.zt_one = g.S2 * 128 * 32;
        g.zt_buf = nbuf * g.zt_one;
        g.off_tabA = 0;
        g.off_tabB = pad_up(2 * nbuf * g.tabA_one + 128 * 32, 1024);            // slack: M = 128 reads past H rows
        g.off_b1 = g.off_tabB + pad_up(nbuf * g.tabB_one, 1024);
        g.nd2 = (512 - 2 * g.N1) / W;           // stage-2 accumulators that fit in tensor memory
        if (g.nd2 > 3) g.nd2 = 3;
        // tuning overrides (diagnostics): PNO_INV_ND2 / PNO_INV_NB1 / PNO_INV_NZT cap the buffer counts
        if (e_nd2 && atoi(e_nd2) >= 1 && atoi(e_nd2) < g.nd2) g.nd2 = atoi(e_nd2);
        // Measured on B200 (tools/inv_sweep.py): a second B1 set or a third Zt set pushes shared memory past ~200 KB, the L1
        // carve-out that is left can no longer track the in-flight addend / spectrum loads, and the kernel gets ~10% SLOWER
        // (224 -> 246 us at the headline shape).  Default: one B1 set, two Zt sets.
        const int nb1_max = e_nb1 && atoi(e_nb1) == 2 ? 2 : 1, nzt_max = e_nzt && atoi(e_nzt) == 3 ? 3 : 2;
        // deepest buffering that fits: prefer 16 epilogue warps, then a second B1 set, then a third Zt set
        int best = -1;
        for (int epi = epi_max; epi >= 8; epi -= 8)
            for (int nb1 = nb1_max; nb1 >= 1; --nb1)
                for (int nzt = nzt_max; nzt >= 1; --nzt) {      // a single Zt set is the last resort (128x128 grids with 32 modes)
                    const int off_zt = g.off_b1 + nb1 * g.b1_buf;
                    const int off_scratch = off_zt + nzt * g.zt_buf;
                    for (int na = add_max; na >= 0; na = (na == 0 ? -1 : (na > 2 ? 2 : 0))) {      // 3, 2 or 0 staged addend tiles
                        const bool to = tma_out_ok && na > 0;
                        const int off_add = pad_up(off_scratch + (to ? 0 : epi * 2048), 1024);
                        const int bytes = off_add + na * add_bytes + 1024;
                        if (bytes > 226 * 1024) continue;            // 227 KB minus the static barriers
                        const int score = epi * 100 + (na > 0 ? 50 : 0) + nb1 * 10 + nzt + na;
                        if (score > best) {
                            best = score; ok = true;
                            g.epi_warps = epi; g.nb1 = nb1; g.nzt = nzt; g.off_zt = off_zt; g.off_scratch = off_scratch;
                            g.off_add = off_add; g.add_stages = na; g.add_bytes = add_bytes; g.smem_bytes = bytes; g.tma_out = to;
                        }
                    }
                }
    }
    if (!ok) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_inv: no channel grouping fits (C = %d, m2 = %d, W = %d)", C, m2, W);
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(addend) | reinterpret_cast<uintptr_t>(pre_act) |
         reinterpret_cast<uintptr_t>(yr) | reinterpret_cast<uintptr_t>(yi)) & 15)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_inv needs 16-byte aligned tensors");
    if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched
    g.n_groups = B * C / g.G;
    g.ldg_t = !g.tma_out && !(getenv("PNO_INV_LDGT") && atoi(getenv("PNO_INV_LDGT")) == 0);
    if (getenv("PNO_INV_PAD")) g.smem_bytes += atoi(getenv("PNO_INV_PAD")) * 1024;   // diagnostics: shrink the L1 carve-out
    g.dbg = g_pno_debug_buffer;
    g.dbg_mma = g_pno_debug_buffer != nullptr && getenv("PNO_DEBUG_MMA") != nullptr;
    InvTables* t = nullptr;
    int rc = inv_tables_get(p, &t);
    if (rc) return rc;
    const float* tabA = g.ts ? t->tabA_ts : t->tabA;
    const float* tabB = g.ts ? (adjoint ? t->tabB_ts_adj : t->tabB_ts) : (adjoint ? t->tabB_adj : t->tabB);
    CUtensorMap tmAdd, tmOut;
    memset(&tmAdd, 0, sizeof(tmAdd));
    memset(&tmOut, 0, sizeof(tmOut));
    if (g.add_stages > 0) {
        const uint64_t dims[2] = {(uint64_t)W, (uint64_t)B * C * H};
        const uint64_t str[1] = {(uint64_t)W * 4};
        const uint32_t box[2] = {32, 128};
        if ((rc = pno_encode_tmap(&tmAdd, addend, 2, dims, str, box, 1))) return rc;
        if (g.tma_out && (rc = pno_encode_tmap(&tmOut, y, 2, dims, str, box, 1))) return rc;
    }
    int dev = 0, sms = 148;
    PNO_CUDA(cudaGetDevice(&dev));
    PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms = pno_usable_sms(sms);
    const int grid = g.n_groups < sms ? g.n_groups : sms;
#define PNO_LAUNCH_INV(NS, GG, AC, TSM)                                                                                         \
    do {                                                                                                                        \
        PNO_CUDA(cudaFuncSetAttribute(k_tc_dft_inv<NS, GG, AC, TSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes)); \
        PNO_CUDA(pno_launch(k_tc_dft_inv<NS, GG, AC, TSM>, grid, kFront + g.epi_warps * 32, g.smem_bytes, st, g,                \
                            tabA, tabB, (const float*)p->coef, yr, yi, addend, y, pre_act, tmAdd, tmOut));                      \
    } while (0)
#define PNO_LAUNCH_INV_ACT(NS, GG, TSM)                      \
    do {                                                     \
        if (act == PNO_ACT_GELU) PNO_LAUNCH_INV(NS, GG, PNO_ACT_GELU, TSM);      \
        else if (act == PNO_ACT_RELU) PNO_LAUNCH_INV(NS, GG, PNO_ACT_RELU, TSM); \
        else PNO_LAUNCH_INV(NS, GG, PNO_ACT_NONE, TSM);      \
    } while (0)
    if (x3 && g.G == 8) PNO_LAUNCH_INV_ACT(3, 8, false);
    else if (x3) PNO_LAUNCH_INV_ACT(3, 4, false);
    else if (g.ts && g.G == 8) PNO_LAUNCH_INV_ACT(1, 8, true);
    else if (g.ts) PNO_LAUNCH_INV_ACT(1, 4, true);
    else if (g.G == 8) PNO_LAUNCH_INV_ACT(1, 8, false);
    else PNO_LAUNCH_INV_ACT(1, 4, false);
#undef PNO_LAUNCH_INV_ACT
#undef PNO_LAUNCH_INV
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
