// Pruned forward 2-D transform on the tensor cores (engines PNO_ENGINE_TC_*):
//   xhat[kx2*m2+ky][b][c] = sum_{h,w} x[b][c][h][w] e^{-2 pi i (kx h/H + ky w/W)}        (reference models.py:89, 93-94)
// as a chain of two dense twiddle-matrix GEMMs per image with the intermediate kept on chip:
//   stage 1 (rows):    R[(img,h)][c2]   = X[(img,h)][w] * TwT[c2][w]^T     c2 = ky (cos) | m2+ky (-sin);  M=128 rows
//   transposer:        TMEM R -> registers -> (tf32 hi/lo) -> K-major swizzled smem operands Rre[(img,ky)][h], Rim[..][h]
//   stage 2 (columns): Xre[kx2][(img,ky)] = C Rre^T + S Rim^T ;  Xim = C Rim^T - S Rre^T        (A = twiddles, negate-A bit)
// Stage-2 accumulators of a "super tile" (up to 8 consecutive channels) stay in TMEM so that the epilogue writes
// [mode][b][c] with 32-byte contiguous channel runs.  All twiddle tables are shared-memory resident; the ring only carries x.
// Forms of stage 2 (FwdGeom::stack / ncat / kym; the launcher picks, PNO_FWD_* override):
//   * stacked table T = [S; C; -S]: one accumulator whose rows [0, R2) are Re and [R2, 2 R2) are Im, half the MMAs;
//   * ky-major accumulator columns (ky * GB + image): the epilogue reads 4 ky x 4 channels with one 16-column TMEM load;
//   * fp32 parity, N-concatenated operands: stage 1 = {x, lo(x)} x [TwT hi | TwT lo], stage 2 = T x [R | lo(R)] + lo(T) x R, the
//     halves (main | correction columns) added by the transposer / epilogue; the raw fp32 tile is the hi operand (the tensor
//     core truncates it), lo(v) = RN_tf32(v - trunc(v)).
// In this kernel the MMAs are bound by the shared-memory reads of their own operands (an N = 80 MMA fetches 6.5 KB in the 53
// cycles it lasts alone), so everything that enlarges N per MMA or takes other traffic off shared memory pays.
//   warp 0        TMA producer (tables once, then x row tiles)
//   warps 1, 22   MMA issuers (warp 22 only in 3xTF32 mode, where stage 2 gets its own issuer).  A tcgen05.mma costs >= ~53 cycles whatever its N
//                 (tools/micro/mma_rate.cu), so stage 2 runs once per batch of TPB tiles with N = TPB*IPT*m2 (80 at the
//                 headline shape) instead of once per tile with N = 48; B2 operands and stage-2 accumulators are double
//                 buffered per batch so that transposer, stage 2 and epilogue of consecutive batches overlap
//   warps 2-5     tf32 hi/lo splitter of the x tiles (3xTF32 only)
//   warps 6-13    transposer: two warps per TMEM lane quarter (cos columns / sin columns), all loads of a tile in flight
//   warps 14-21   epilogue: two warps per lane quarter (re / im plane); only quarters holding kx2 < 2*m1 have work
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int kThreads = 64 + 128 + 256 + 256 + 32;   // + the stage-2 issuer warp

struct FwdGeom {
    int H, W, m1, m2, M;
    int B, C;
    int IPT;        // images per 128-row tile (H <= 128)
    int N1p;        // stage-1 N: 2*m2 padded to 16
    int N2p;        // stage-2 N: TPB*IPT*m2 padded to 16
    int R2;         // rows of the C / S tables: 2*m1 padded to 8
    int KB1;        // W / 32
    int KB2;        // H / 32
    int TPS;        // tiles per super tile
    int TPB;        // tiles per stage-2 batch (TPS % TPB == 0)
    int n_super;    // super tiles in the launch
    int grad_scale;
    int KPS;        // x k-blocks per ring stage (a whole tile when W <= 64: one barrier round trip per tile instead of per k-block)
    int stages, nslot, nb2;   // ring depth, stage-2 accumulator buffers, B2 operand buffers
    int swap13;               // physical warps 1 and 3 swap roles (see the kernel)
    int kym;                  // single pass with stacked tables: ky-major accumulator columns and 16-column epilogue loads (the
                              // fp32-parity form with N-concatenated operands always has them)
    int ncat;                 // fp32-parity form with N-concatenated operands (see issue_stage1 / issue_stage2): 2 + 4 instead of 3 + 6 MMAs
                              // per k-step; needs the stacked tables
    int D1W;                  // columns of one stage-1 accumulator buffer: N1p, or 2*tw_rows (main | correction) with ncat
    int stack;                // stage 2 with the stacked table T = [S; C; -S] (see issue_stage2): half the MMAs, one accumulator
    int ablate;               // measurement only (PNO_FWD_ABLATE, wrong results): 1 = no correction MMAs in stage 1, 2 = none in stage 2
    int epace;                // fp32-parity form: nanoseconds an epilogue warp sleeps after each of its scattered 16-byte stores (see there)
    int pf;                   // L2 prefetch distance of the x tiles, in 128-row tiles (0 = off)
    int tw_rows;              // rows of one TwT k-block in shared memory (N1p, or 2*m2 rounded to 8 when the padding rows are not stored)
    // shared-memory byte offsets (from the 1024-aligned base)
    int off_tab, off_tw, off_ring, off_b2, stage_bytes, x_bytes, tw_one, tab_one, b2_one, b2_bytes, smem_bytes;
    long long* dbg;   // optional event trace (PNO_TRACE builds)
};

#define TRACE(role, ev) PNO_TRACE_EV(g.dbg, role, ev)

struct TcTables {
    float *twT_hi, *twT_lo;   // [N1p][W]
    float *c_hi, *c_lo, *s_hi, *s_lo;   // [R2][H]
    float *t_hi, *t_lo;                 // [3*R2][H]: S, C, -S stacked (rows [0, 2 R2) = [S; C], rows [R2, 3 R2) = [C; -S])
    // inverse-transform tables (pno_tc_dft_inv.cu)
    float *inv_a_hi, *inv_a_lo, *inv_w_hi, *inv_w_lo;
    int N1p, R2;
};

inline float tf32_round_host(double v) {
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&f, &u, 4);
    return f;
}

__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}

template <int NSPLIT, bool NCAT>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_dft_fwd(const FwdGeom g, const float* __restrict__ coef, float* __restrict__ xr, float* __restrict__ xi,
             const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmTwHi,
             const __grid_constant__ CUtensorMap tmTwLo, const __grid_constant__ CUtensorMap tmCHi,
             const __grid_constant__ CUtensorMap tmCLo, const __grid_constant__ CUtensorMap tmSHi,
             const __grid_constant__ CUtensorMap tmSLo) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kMaxStages = 8;
    __shared__ __align__(8) uint64_t bar_raw[kMaxStages], bar_split[kMaxStages], bar_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_tab, bar_d1_full[2], bar_d1_empty[2], bar_b2_full[2], bar_b2_empty[2], bar_d2_full[2],
        bar_d2_empty[2];
    __shared__ uint32_t tmem_slot;

    // Role index of this warp.  Physical warps 1 and 3 swap roles: the stage-1 MMA issuer (role 1) then runs on sub-partition 3 and
    // the stage-2 issuer of the 3xTF32 form (role 22) on sub-partition 2 -- the two sub-partitions whose epilogue warps hold TMEM
    // lane quarters beyond the 2*m1 retained rows and idle.  An issuer that shares its sub-partition with busy transposer AND
    // epilogue warps paid 140 (stage 1) / 80 (stage 2) cycles per MMA instead of ~53 (tools/trace_fwd.py, round 2).
    const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = g.swap13 ? (pwarp == 1 ? 3 : pwarp == 3 ? 1 : pwarp) : pwarp;
    PNO_TRACE_DECL();
    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_split[s], 128);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_tab, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_d1_full[s], 1);
            mbar_init(&bar_d1_empty[s], 256);
            mbar_init(&bar_b2_full[s], 256 * g.TPB);
            mbar_init(&bar_b2_empty[s], 1);
            mbar_init(&bar_d2_full[s], 1);
            mbar_init(&bar_d2_empty[s], 256);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                        // x (the previous kernel's output) is complete from here on

    constexpr int NT = NSPLIT == 3 ? 2 : 1;            // hi (+ lo) copies of every operand
    constexpr bool KYM = NCAT;                         // ky-major stage-2 accumulator columns + 16-column epilogue loads
    constexpr bool CAT = NCAT && NSPLIT == 3;          // N-concatenated hi / lo operands (fp32-parity form)
    // TMEM columns: [0, 2*N1p) stage-1 accumulators (double buffered); then nslot buffers of 2*N2p (re | im), one batch each
    const int d2_base = 2 * g.D1W;
    // tables: (C hi, [C lo], S hi, [S lo]), each KB2 k-blocks of [R2 x 128 B]; TwT (hi, [lo]) each KB1 k-blocks of [N1p x 128 B]
    unsigned char* tab = smem + g.off_tab;
    unsigned char* twt = smem + g.off_tw;
    const int tab_kb = (g.stack ? 3 : 1) * g.R2 * 128;
    const int tw_kb = (CAT ? 2 : 1) * g.tw_rows * 128;      // ncat: [hi rows | lo rows] per k-block
    // B2 operands: (Rre hi, [Rre lo], Rim hi, [Rim lo]), each KB2 k-blocks of [N2p x 128 B]; rows = (tile in batch, image, ky)
    // ncat: [hi rows | lo rows] per k-block, so that one N = 2*N2p MMA multiplies a table by both
    const int b2_kb = (CAT ? 2 : 1) * g.N2p * 128;
    const int b2_lo = CAT ? g.N2p * 128 : g.b2_one;          // lo copy of an element, relative to its hi copy

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar_tab, (g.stack ? 1 : 2) * NT * g.tab_one + NT * g.tw_one);
            for (int kb = 0; kb < g.KB2; ++kb) {
                if (g.stack) {          // tmCHi / tmCLo carry the stacked table (hi / lo)
                    tma_load_2d(tab + kb * tab_kb, &tmCHi, &bar_tab, kb * 32, 0);
                    if (NSPLIT == 3) tma_load_2d(tab + g.tab_one + kb * tab_kb, &tmCLo, &bar_tab, kb * 32, 0);
                    continue;
                }
                tma_load_2d(tab + 0 * g.tab_one + kb * tab_kb, &tmCHi, &bar_tab, kb * 32, 0);
                tma_load_2d(tab + NT * g.tab_one + kb * tab_kb, &tmSHi, &bar_tab, kb * 32, 0);
                if (NSPLIT == 3) {
                    tma_load_2d(tab + 1 * g.tab_one + kb * tab_kb, &tmCLo, &bar_tab, kb * 32, 0);
                    tma_load_2d(tab + 3 * g.tab_one + kb * tab_kb, &tmSLo, &bar_tab, kb * 32, 0);
                }
            }
            for (int kb = 0; kb < g.KB1; ++kb) {
                tma_load_2d(twt + kb * tw_kb, &tmTwHi, &bar_tab, kb * 32, 0);
                if (NSPLIT == 3) tma_load_2d(twt + (CAT ? g.tw_rows * 128 : g.tw_one) + kb * tw_kb, &tmTwLo, &bar_tab, kb * 32, 0);
            }
            int stage = 0, phase = 0;
            // this CTA's tiles are TPS consecutive 128-row tiles per super tile, super tiles gridDim.x apart: pull the tile that is
            // g.pf tiles ahead into L2 so that the ring (2-3 stages: tables and B2 operands take most of the shared memory) is filled
            // at L2 latency
            const int total_rows = g.n_super * g.TPS * 128;
            for (int sup = blockIdx.x; sup < g.n_super; sup += gridDim.x) {
                for (int t = 0; t < g.TPS; ++t) {
                    const int row0 = (sup * g.TPS + t) * 128;          // row of the [B*C*H][W] view
                    if (g.pf) {
                        const int ta = t + g.pf;                       // tile index ahead, in this CTA's sequence
                        const int prow = ((sup + (ta / g.TPS) * (int)gridDim.x) * g.TPS + ta % g.TPS) * 128;
                        if (prow < total_rows)
                            for (int kb = 0; kb < g.KB1; ++kb) tma_prefetch_l2_2d(&tmX, kb * 32, prow);
                    }
                    for (int kb = 0; kb < g.KB1; kb += g.KPS) {
                        mbar_wait(&bar_empty[stage], phase ^ 1);
                        TRACE(3, 1);
                        unsigned char* st = smem + g.off_ring + (size_t)stage * g.stage_bytes;
                        mbar_arrive_expect_tx(&bar_raw[stage], g.KPS * g.x_bytes);
                        for (int j = 0; j < g.KPS; ++j) tma_load_2d(st + j * g.x_bytes, &tmX, &bar_raw[stage], (kb + j) * 32, row0);
                        if (++stage == g.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 22) {
        // ================================ MMA issuers: warp 1 = stage 1, warp 22 = stage 2 ================================
        {   // all 32 lanes run the loop (uniform registers); `el` predicates the tcgen05 instructions
            const uint32_t el = lane == 0;
            const uint32_t idesc1 = make_idesc(FMT_TF32, 128, g.N1p, MAJOR_K, MAJOR_K);
            const uint32_t idesc2 = make_idesc(FMT_TF32, 128, g.N2p, MAJOR_K, MAJOR_K);
            const uint32_t idesc2n = make_idesc(FMT_TF32, 128, g.N2p, MAJOR_K, MAJOR_K, 1, 0);
            const uint32_t idesc1c = make_idesc(FMT_TF32, 128, 2 * g.tw_rows, MAJOR_K, MAJOR_K);      // ncat forms
            const uint32_t idesc2c = make_idesc(FMT_TF32, 128, 2 * g.N2p, MAJOR_K, MAJOR_K);
            const uint32_t khi = sdesc_base(0, 16, 1024, SWIZZLE_128B).hi;     // all operands: K-major SW128
            const uint32_t tb = sdesc_base(smem_u32(tab), 16, 1024, SWIZZLE_128B).lo;
            const uint32_t twb = sdesc_base(smem_u32(twt), 16, 1024, SWIZZLE_128B).lo;
            const uint32_t t1 = g.tab_one >> 4, bo1 = g.b2_one >> 4;
            int stage = 0, phase = 0;
            int n1 = 0;      // stage-1 tiles issued
            int n2 = 0;      // stage-2 batches issued
            int i_slot = 0, i_sph = 0, i_bb = 0, i_bph = 0, i_tb = 0;   // ring positions / phases (no divisions in the issue loop)
            mbar_wait(&bar_tab, 0);

            auto issue_stage1 = [&]() {
                const int buf = n1 & 1;
                mbar_wait(&bar_d1_empty[buf], ((n1 >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d1 = tmem + buf * g.D1W;
                for (int kb0 = 0; kb0 < g.KB1; kb0 += g.KPS) {
                    mbar_wait(NSPLIT == 3 ? &bar_split[stage] : &bar_raw[stage], phase);
                    TRACE(0, 5);
                    tc_fence_after_sync();
                    const uint32_t st_lo = sdesc_base(smem_u32(smem + g.off_ring + (size_t)stage * g.stage_bytes), 16, 1024, SWIZZLE_128B).lo;
                    for (int j = 0; j < g.KPS; ++j) {
                        const int kb = kb0 + j;
                        const uint32_t x_hi = st_lo + ((j * g.x_bytes) >> 4);
                        const uint32_t x_lo = x_hi + ((g.KPS * g.x_bytes) >> 4);
                        const uint32_t tw_hi = twb + ((kb * tw_kb) >> 4), tw_lo = tw_hi + (g.tw_one >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t accf = (kb | ks) ? 1u : 0u;
                            const uint32_t o = ks * 2;
                            if (CAT) {
                                // B = [TwT hi | TwT lo] (N = 2*tw_rows): accumulator columns [0, tw_rows) = x * hi, [tw_rows, ..) = x * lo,
                                // for A = the raw x tile (the tensor core reads its tf32 truncation) and A = its lo part: all four
                                // terms of the product in two MMAs instead of three terms in three.  The transposer adds the halves.
                                mma_tf32(d1, x_hi + o, khi, tw_hi + o, khi, idesc1c, accf, el);
                                mma_tf32(d1, x_lo + o, khi, tw_hi + o, khi, idesc1c, 1u, el);
                                continue;
                            }
                            mma_tf32(d1, x_hi + o, khi, tw_hi + o, khi, idesc1, accf, el);
                            if (NSPLIT == 3 && g.ablate != 1) {
                                mma_tf32(d1, x_lo + o, khi, tw_hi + o, khi, idesc1, 1u, el);
                                mma_tf32(d1, x_hi + o, khi, tw_lo + o, khi, idesc1, 1u, el);
                            }
                        }
                    }
                    mma_commit(&bar_empty[stage], el);
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
                mma_commit(&bar_d1_full[buf], el);
                TRACE(0, 1);
                ++n1;
            };

            auto issue_stage2 = [&]() {                         // one batch of TPB tiles
                const int slot = i_slot, bb = i_bb;
                mbar_wait(&bar_d2_empty[slot], i_sph ^ 1);   // the epilogue has drained this accumulator buffer
                TRACE(0, 4);
                mbar_wait(&bar_b2_full[bb], i_bph);
                TRACE(0, 2);
                tc_fence_after_sync();
                const uint32_t d_re = tmem + d2_base + slot * 2 * g.N2p, d_im = d_re + g.N2p;
                const uint32_t bbase = sdesc_base(smem_u32(smem + g.off_b2 + bb * g.b2_bytes), 16, 1024, SWIZZLE_128B).lo;
                // tables (C hi, [C lo], S hi, [S lo]) at tb + {0, t1, NT*t1, (NT+1)*t1}; B2 (Rre hi, [Rre lo], Rim hi, [Rim lo]) likewise
                for (int kb = 0; kb < g.KB2; ++kb) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t first = (kb | ks) ? 1u : 0u;
                        const uint32_t c = tb + ((kb * tab_kb) >> 4) + ks * 2, r = bbase + ((kb * b2_kb) >> 4) + ks * 2;
                        const uint32_t sn = c + NT * t1, ri = r + NT * bo1;
                        if (CAT) {
                            // stacked tables (below) times B = [R | R lo] (N = 2*N2p): columns [0, N2p) main, [N2p, 2 N2p) correction;
                            // the lo tables times R go to the correction columns.  4 MMAs per k-step instead of 6.
                            const uint32_t a_cs = c + ((g.R2 * 128) >> 4), a_sc = c;
                            mma_tf32(d_re, a_cs, khi, r, khi, idesc2c, first, el);
                            mma_tf32(d_re, a_sc, khi, ri, khi, idesc2c, 1u, el);
                            mma_tf32(d_im, a_cs + t1, khi, r, khi, idesc2, 1u, el);
                            mma_tf32(d_im, a_sc + t1, khi, ri, khi, idesc2, 1u, el);
                            continue;
                        }
                        if (g.stack) {
                            // one accumulator, rows [0, R2) = Xre = C Rre + S Rim, rows [R2, 2 R2) = Xim = -S Rre + C Rim:
                            // A = [C; -S] (table rows R2 ..) with Rre, A = [S; C] (table rows 0 ..) with Rim
                            const uint32_t a_cs = c + ((g.R2 * 128) >> 4), a_sc = c;
                            mma_tf32(d_re, a_cs, khi, r, khi, idesc2, first, el);
                            mma_tf32(d_re, a_sc, khi, ri, khi, idesc2, 1u, el);
                            if (NSPLIT == 3 && g.ablate != 2) {
                                mma_tf32(d_re, a_cs + t1, khi, r, khi, idesc2, 1u, el);
                                mma_tf32(d_re, a_cs, khi, r + bo1, khi, idesc2, 1u, el);
                                mma_tf32(d_re, a_sc + t1, khi, ri, khi, idesc2, 1u, el);
                                mma_tf32(d_re, a_sc, khi, ri + bo1, khi, idesc2, 1u, el);
                            }
                            continue;
                        }
                        // Xre += C Rre + S Rim ; Xim += C Rim - S Rre
                        mma_tf32(d_re, c, khi, r, khi, idesc2, first, el);
                        mma_tf32(d_re, sn, khi, ri, khi, idesc2, 1u, el);
                        mma_tf32(d_im, c, khi, ri, khi, idesc2, first, el);
                        mma_tf32(d_im, sn, khi, r, khi, idesc2n, 1u, el);
                        if (NSPLIT == 3) {
                            mma_tf32(d_re, c + t1, khi, r, khi, idesc2, 1u, el);
                            mma_tf32(d_re, c, khi, r + bo1, khi, idesc2, 1u, el);
                            mma_tf32(d_re, sn + t1, khi, ri, khi, idesc2, 1u, el);
                            mma_tf32(d_re, sn, khi, ri + bo1, khi, idesc2, 1u, el);
                            mma_tf32(d_im, c + t1, khi, ri, khi, idesc2, 1u, el);
                            mma_tf32(d_im, c, khi, ri + bo1, khi, idesc2, 1u, el);
                            mma_tf32(d_im, sn + t1, khi, r, khi, idesc2n, 1u, el);
                            mma_tf32(d_im, sn, khi, r + bo1, khi, idesc2n, 1u, el);
                        }
                    }
                }
                mma_commit(&bar_b2_empty[bb], el);
                mma_commit(&bar_d2_full[slot], el);
                TRACE(0, 3);
                ++n2;
                if (++i_slot == g.nslot) { i_slot = 0; i_sph ^= 1; }
                if (++i_bb == g.nb2) { i_bb = 0; i_bph ^= 1; }
            };

            int my_tiles = 0;
            for (int sup = blockIdx.x; sup < g.n_super; sup += gridDim.x) my_tiles += g.TPS;
            if (NSPLIT == 3) {
                // 3xTF32 triples the MMA count and the issue + commit cost of one thread becomes the bound: the two stages are
                // issued by different warps (measured 230 -> 197 us at the headline shape)
                if (warp == 1) {
                    for (int t = 0; t < my_tiles; ++t) issue_stage1();
                } else {
                    for (int t = 0; t < my_tiles; t += g.TPB) issue_stage2();
                }
            } else if (warp == 1) {
                // single pass: one issuer, software pipelined (two issuers contend and measure slower: 123 vs 132 us)
                if (my_tiles > 0) issue_stage1();
                for (int t = 0; t < my_tiles; ++t) {
                    if (t + 1 < my_tiles) issue_stage1();
                    if (++i_tb == g.TPB) { i_tb = 0; issue_stage2(); }
                }
            }
        }
    } else if (warp < 6) {
        // ================================ hi/lo splitter of the x tile (3xTF32) ================================
        if (NSPLIT == 3) {
            const int t = warp * 32 + lane - 64;
            int stage = 0, phase = 0;
            for (int sup = blockIdx.x; sup < g.n_super; sup += gridDim.x) {
                for (int tt = 0; tt < g.TPS * (g.KB1 / g.KPS); ++tt) {
                    mbar_wait(&bar_raw[stage], phase);
                    const uint32_t hi = smem_u32(smem + g.off_ring + (size_t)stage * g.stage_bytes), lo = hi + g.KPS * g.x_bytes;
#pragma unroll 4
                    for (int i = t; i < g.KPS * g.x_bytes / 16; i += 128) {
                        // split by truncation: kind::tf32 ignores the 13 low mantissa bits, so the staged tile itself is the hi
                        // operand (no store) and lo = v - trunc(v)
                        const float4 v = lds128(hi + 16 * i);
                        sts128(lo + 16 * i, tf32_lo_rn4(v));
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&bar_split[stage]);
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 14) {
        // ================================ transposer ================================
        // thread = row (img_l, h) of the tile (TMEM lane) x one half of the stage-1 columns (cos -> Rre, -sin -> Rim)
        const int q = warp & 3, im = (warp - 6) >> 2;
        const int row = q * 32 + lane;
        const int img_l = row / g.H, hk = row % g.H;
        // element (n2 = img_l*m2 + ky, k = hk) of a K-major SW128 [N2p x 32] k-block: only the row term depends on ky
        const uint32_t col_off = (uint32_t)((hk >> 5) * b2_kb + ((hk & 3) << 2));
        const int chunk = (hk & 31) >> 2;
        int n = 0;
        int tb_ = 0, bb = 0, bph = 0;           // tile within its stage-2 batch, B2 buffer and its phase (no divisions in the loop)
        for (int sup = blockIdx.x; sup < g.n_super; sup += gridDim.x) {
            for (int t = 0; t < g.TPS; ++t, ++n) {
                const int buf = n & 1;
                mbar_wait(&bar_d1_full[buf], (n >> 1) & 1);
                if (warp == 8) TRACE(1, 1);
                tc_fence_after_sync();
                if (tb_ == 0) mbar_wait(&bar_b2_empty[bb], bph ^ 1);   // stage 2 of the batch that used this buffer is done
                if (warp == 8) TRACE(1, 2);
                const uint32_t t1 = tmem_addr(tmem, q * 32, buf * g.D1W + im * g.m2);
                const uint32_t dst0 = smem_u32(smem + g.off_b2) + bb * g.b2_bytes + im * NT * g.b2_one + col_off;
                // row of the B2 operand = stage-2 accumulator column: ky * GB + image of the batch, so that the GB images (consecutive
                // channels) of one ky are adjacent columns and the epilogue reads whole 16-column runs (4 ky x 4 images = four
                // 16-byte stores) with one wide tensor-memory load instead of one narrow load per image
                // (fp32-parity form; the single-pass kernel keeps image-major columns and its narrow loads: 128 vs 136 us)
                const int GBt = KYM ? g.TPB * g.IPT : 1;
                const int r0 = KYM ? tb_ * g.IPT + img_l : (tb_ * g.IPT + img_l) * g.m2;
                if (CAT) {
                    // main + correction halves of the stage-1 accumulator, 16 columns a pass (two 16-register loads in flight)
                    for (int c0 = 0; c0 < g.m2; c0 += 16) {
                        uint32_t ra[4][4], rb[4][4];
                        tmem_ld_cols16(t1 + c0, g.m2 - c0, ra);
                        tmem_ld_cols16(t1 + g.tw_rows + c0, g.m2 - c0, rb);
                        tmem_ld_wait();
                        if (warp == 8) TRACE(1, 3);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (c0 + 4 * u < g.m2) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int r = r0 + (c0 + 4 * u + j) * GBt;
                                    const uint32_t dst = dst0 + r * 128 + ((chunk ^ (r & 7)) << 4);
                                    const float v = __uint_as_float(ra[u][j]) + __uint_as_float(rb[u][j]);
                                    sts32(dst, v);                       // raw: the tensor core truncates it
                                    sts32(dst + b2_lo, tf32_lo_rn(v));
                                }
                            }
                        }
                    }
                } else
                for (int c0 = 0; c0 < g.m2; c0 += 32) {          // up to 32 columns with the widest loads (x16 / x8 / x4), one wait
                    uint32_t rv[8][4];
                    tmem_ld_cols(t1 + c0, g.m2 - c0, rv);
                    tmem_ld_wait();
                    float v[8][4];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[u][j] = __uint_as_float(rv[u][j]);
                    if (warp == 8) TRACE(1, 3);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (c0 + 4 * u < g.m2) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int r = r0 + (c0 + 4 * u + j) * GBt;
                                const uint32_t dst = dst0 + r * 128 + ((chunk ^ (r & 7)) << 4);
                                if (NSPLIT == 3) {
                                    const float h = tf32_rn(v[u][j]);
                                    sts32(dst, h);
                                    sts32(dst + g.b2_one, v[u][j] - h);
                                } else {
                                    sts32(dst, v[u][j]);
                                }
                            }
                        }
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(&bar_d1_empty[buf]);
                fence_proxy_async_smem();
                mbar_arrive(&bar_b2_full[bb]);
                if (warp == 8) TRACE(1, 4);
                if (++tb_ == g.TPB) {
                    tb_ = 0;
                    if (++bb == g.nb2) { bb = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp < 22) {
        // ================================ epilogue ================================
        // lanes kx2 < 2*m1 of a D2 buffer hold Xre / Xim of one batch = GB = TPB*IPT images (consecutive channels);
        // warp = (lane quarter, re / im plane).  The batches of a super tile are adjacent pieces of the same 32-byte runs.
        const int q = warp & 3, part = (warp - 14) >> 2;
        const int lrow = q * 32 + lane;                  // TMEM lane
        const int GB = g.TPB * g.IPT;
        // separate accumulators: lanes [0, 2 m1) of buffer `part` (re / im).  Stacked: ONE accumulator, lanes [0, R2) = re and
        // [R2, 2 R2) = im rows; the two warps of a quarter take alternate groups of four ky
        const int half = g.stack ? lrow / g.R2 : part;
        const int row = g.stack ? lrow - half * g.R2 : lrow;        // kx2
        const bool warp_active = q * 32 < (g.stack ? 2 * g.R2 : 2 * g.m1);      // warp-uniform: tcgen05.ld is warp-collective
        const bool active = row < 2 * g.m1 && half < 2;
        float* dst_plane = half ? xi : xr;
        // (x8 loads of 8 ky per image were tried here: 64 values in flight spill at the 80-register cap of this 22-23 warp CTA:
        //  128 -> 193 us)
        const int ch_first = g.stack ? part : 0, ch_step = g.stack ? 2 : 1;      // NCAT: 16-column passes of this warp
        const int ky_first = g.stack ? part * 4 : 0, ky_step = g.stack ? 8 : 4;
        const int bps = g.TPS / g.TPB;                   // batches per super tile
        int slot = 0, sph = 0;
        for (int sup = blockIdx.x; sup < g.n_super; sup += gridDim.x) {
            const int simg0 = sup * bps * GB;                // first image of the super tile = b*C + c0 (one division per super tile)
            const int sb = simg0 / g.C, sc0 = simg0 - sb * g.C;
            for (int k = 0; k < bps; ++k) {
                mbar_wait(&bar_d2_full[slot], sph);
                if (warp == 16) TRACE(2, 1);
                tc_fence_after_sync();
                if (warp_active) {
                    const int b = sb, c0 = sc0 + k * GB;         // a super tile never straddles a batch item (C % (bps*GB) == 0)
                    const uint32_t tb0 = tmem_addr(tmem, q * 32, d2_base + slot * 2 * g.N2p + (g.stack ? 0 : part * g.N2p));
                    if (KYM) {
                    // accumulator columns = ky * GB + image: 16 columns a pass = 16 / GB values of ky x GB consecutive channels
                    const int ncols = g.m2 * GB;
                    for (int cc = ch_first; cc * 16 < ncols; cc += ch_step) {
                        uint32_t ra[4][4];
                        float v[4][4];
                        tmem_ld_cols16(tb0 + cc * 16, 16, ra);
                        if (CAT) {                                 // main + correction columns (N2p apart)
                            uint32_t rb[4][4];
                            tmem_ld_cols16(tb0 + g.N2p + cc * 16, 16, rb);
                            tmem_ld_wait();
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[u][j] = __uint_as_float(ra[u][j]) + __uint_as_float(rb[u][j]);
                        } else {
                            tmem_ld_wait();
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[u][j] = __uint_as_float(ra[u][j]);
                        }
                        if (!active) continue;
                        if (GB == 4) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int mode = row * g.m2 + cc * 4 + u;
                                const float cf = g.grad_scale ? coef[mode] : 1.f;
                                float* dst = dst_plane + ((size_t)mode * g.B + b) * g.C + c0;
                                *reinterpret_cast<float4*>(dst) = make_float4(v[u][0] * cf, v[u][1] * cf, v[u][2] * cf, v[u][3] * cf);
                                // Pacing.  Every lane of such a store hits its own 128-byte line (modes are B*C*4 bytes apart), i.e. 32
                                // LSU passes per instruction; issued back to back by six warps they fill the load / store queue for
                                // ~3.5 k cycles per batch and the transposer's shared-memory stores of the NEXT batch -- the critical
                                // chain transposer -> stage 2 -> transposer -- queue behind them (its 16-column pass took 3.7 k cycles
                                // under the epilogue, 0.8 k alone: tools/trace_fwd.py).  The epilogue has a whole batch period to
                                // spare (two accumulator slots), so it spreads its stores: 169 -> 159 us at 150 ns (100: 162, 200:
                                // 161, 300: 197 = the epilogue becomes the bound).  The single-pass kernel has four active epilogue
                                // warps and no slack: paced the same way it goes from 127 to 202 us, so it is not.
                                if (g.epace) __nanosleep(g.epace);
                            }
                        } else if (GB == 8) {
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int mode = row * g.m2 + cc * 2 + u;
                                const float cf = g.grad_scale ? coef[mode] : 1.f;
                                float* dst = dst_plane + ((size_t)mode * g.B + b) * g.C + c0;
                                reinterpret_cast<float4*>(dst)[0] = make_float4(v[2 * u][0] * cf, v[2 * u][1] * cf, v[2 * u][2] * cf, v[2 * u][3] * cf);
                                reinterpret_cast<float4*>(dst)[1] = make_float4(v[2 * u + 1][0] * cf, v[2 * u + 1][1] * cf, v[2 * u + 1][2] * cf, v[2 * u + 1][3] * cf);
                            }
                        } else {                                    // GB = 1 or 2
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int col = cc * 16 + 4 * u + j;
                                    if (col < ncols) {
                                        const int ky = GB == 2 ? col >> 1 : col, gi = GB == 2 ? col & 1 : 0;
                                        const int mode = row * g.m2 + ky;
                                        const float cf = g.grad_scale ? coef[mode] : 1.f;
                                        dst_plane[((size_t)mode * g.B + b) * g.C + c0 + gi] = v[u][j] * cf;
                                    }
                                }
                        }
                    }
                    } else {
                    for (int ky0 = ky_first; ky0 < g.m2; ky0 += ky_step) {
                        float vals[8][4];
#pragma unroll
                        for (int gi = 0; gi < 8; ++gi) {
                            if (gi < GB) tmem_ld4(tb0 + gi * g.m2 + ky0, vals[gi]);
                            else vals[gi][0] = vals[gi][1] = vals[gi][2] = vals[gi][3] = 0.f;
                        }
                        tmem_ld_wait();
                        if (active) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int mode = row * g.m2 + ky0 + j;
                                const float cf = g.grad_scale ? coef[mode] : 1.f;
                                float* dst = dst_plane + ((size_t)mode * g.B + b) * g.C + c0;
                                if (GB == 8) {
                                    reinterpret_cast<float4*>(dst)[0] = make_float4(vals[0][j] * cf, vals[1][j] * cf, vals[2][j] * cf, vals[3][j] * cf);
                                    reinterpret_cast<float4*>(dst)[1] = make_float4(vals[4][j] * cf, vals[5][j] * cf, vals[6][j] * cf, vals[7][j] * cf);
                                } else if (GB == 4) {
                                    reinterpret_cast<float4*>(dst)[0] = make_float4(vals[0][j] * cf, vals[1][j] * cf, vals[2][j] * cf, vals[3][j] * cf);
                                } else {
#pragma unroll
                                    for (int gi = 0; gi < 8; ++gi)
                                        if (gi < GB) dst[gi] = vals[gi][j] * cf;
                                }
                            }
                        }
                    }
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(&bar_d2_empty[slot]);
                if (warp == 16) TRACE(2, 2);
                if (++slot == g.nslot) { slot = 0; sph ^= 1; }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

int pad_to(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

// ---- plan-owned tensor-core tables ------------------------------------------------------------------------------------
static int upload(float** dst, const std::vector<float>& v) {
    PNO_CUDA(cudaMalloc(dst, v.size() * sizeof(float)));
    PNO_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return PNO_OK;
}

int tc_tables_get(const pno_plan* cp, void** out) {
    pno_plan* p = const_cast<pno_plan*>(cp);
    if (p->tc_tables) { *out = p->tc_tables; return PNO_OK; }
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2;
    TcTables* t = new TcTables();
    memset(t, 0, sizeof(*t));
    t->N1p = pad_to(2 * m2, 16);
    t->R2 = pad_to(2 * m1, 8);
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> tw_hi((size_t)t->N1p * W, 0.f), tw_lo(tw_hi.size(), 0.f);
    for (int c2 = 0; c2 < 2 * m2; ++c2)
        for (int w = 0; w < W; ++w) {
            const int ky = c2 % m2;
            const double ang = two_pi * (double)((long long)ky * w % W) / W;
            const double v = c2 < m2 ? cos(ang) : -sin(ang);
            const float hi = tf32_round_host(v);
            tw_hi[(size_t)c2 * W + w] = hi;
            tw_lo[(size_t)c2 * W + w] = tf32_round_host(v - (double)hi);
        }
    std::vector<float> c_hi((size_t)t->R2 * H, 0.f), c_lo(c_hi.size(), 0.f), s_hi(c_hi.size(), 0.f), s_lo(c_hi.size(), 0.f);
    for (int r = 0; r < 2 * m1; ++r) {
        const int kx = r < m1 ? r : H - 2 * m1 + r;
        for (int h = 0; h < H; ++h) {
            const double ang = two_pi * (double)((long long)kx * h % H) / H;
            const double c = cos(ang), s = sin(ang);
            const float ch = tf32_round_host(c), sh = tf32_round_host(s);
            c_hi[(size_t)r * H + h] = ch; c_lo[(size_t)r * H + h] = tf32_round_host(c - (double)ch);
            s_hi[(size_t)r * H + h] = sh; s_lo[(size_t)r * H + h] = tf32_round_host(s - (double)sh);
        }
    }
    // stacked form [S; C; -S] (hi / lo parts negate exactly)
    const size_t n1 = (size_t)t->R2 * H;
    std::vector<float> st_hi(3 * n1), st_lo(3 * n1);
    for (size_t i = 0; i < n1; ++i) {
        st_hi[i] = s_hi[i]; st_hi[n1 + i] = c_hi[i]; st_hi[2 * n1 + i] = -s_hi[i];
        st_lo[i] = s_lo[i]; st_lo[n1 + i] = c_lo[i]; st_lo[2 * n1 + i] = -s_lo[i];
    }
    int rc;
    if ((rc = upload(&t->twT_hi, tw_hi)) || (rc = upload(&t->twT_lo, tw_lo)) || (rc = upload(&t->c_hi, c_hi)) ||
        (rc = upload(&t->c_lo, c_lo)) || (rc = upload(&t->s_hi, s_hi)) || (rc = upload(&t->s_lo, s_lo)) ||
        (rc = upload(&t->t_hi, st_hi)) || (rc = upload(&t->t_lo, st_lo)))
        return rc;
    p->tc_tables = t;
    *out = t;
    return PNO_OK;
}

void tc_plan_release(pno_plan* p) {
    if (!p->tc_tables) return;
    TcTables* t = static_cast<TcTables*>(p->tc_tables);
    float* ptrs[] = {t->twT_hi, t->twT_lo, t->c_hi, t->c_lo, t->s_hi, t->s_lo, t->t_hi, t->t_lo, t->inv_a_hi, t->inv_a_lo, t->inv_w_hi,
                     t->inv_w_lo};
    for (float* q : ptrs)
        if (q) cudaFree(q);
    delete t;
    p->tc_tables = nullptr;
}

int tc_dft_fwd_fused(const pno_plan* p, int engine, const float* x, int B, int C, int grad_scale, float* xr, float* xi,
                     cudaStream_t st) {
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2;
    if (!(H == 32 || H == 64 || H == 128) || W % 32 || W > 256 || m2 % 4 || 2 * m1 > 128)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd tiles H in {32,64,128}, W%%32 == 0, W <= 256, m2%%4 == 0, 2*m1 <= 128");
    FwdGeom g;
    memset(&g, 0, sizeof(g));
    g.H = H; g.W = W; g.m1 = m1; g.m2 = m2; g.M = p->M; g.B = B; g.C = C; g.grad_scale = grad_scale;
    const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
    const int NT = x3 ? 2 : 1;
    g.IPT = 128 / H;
    g.N1p = pad_to(2 * m2, 16);
    g.R2 = pad_to(2 * m1, 8);
    g.KB1 = W / 32;
    g.KB2 = H / 32;
    if (H > 128) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd: H > 128");
    g.x_bytes = 128 * 128;                        // [128 rows][32 floats]
    g.KPS = 1;
    g.stage_bytes = NT * g.x_bytes;
    // stacked stage-2 tables (see the kernel): half the stage-2 MMAs, one accumulator, epilogue work on three lane quarters instead
    // of one and a quarter.  With image-major accumulator columns and one narrow tensor-memory load per image it measured
    // 202 -> 192 us (fp32 parity) but 121 -> 137 us (single pass); with ky-major columns (one 16-column load per four stores,
    // `kym`) the single-pass kernel gains too: 127 -> 114 us.  Single pass: only while the stacked table stays small (<= 32 KB:
    // 64-row images), larger ones cost ring stages and the second B2 buffer.  PNO_FWD_STACK / PNO_FWD_KYM = 0 / 1 force the choice.
    // If the larger tables do not fit, the un-stacked form is tried next.
    const int kym_env = getenv("PNO_FWD_KYM") ? atoi(getenv("PNO_FWD_KYM")) : 1;
    int want_stack = getenv("PNO_FWD_STACK") ? atoi(getenv("PNO_FWD_STACK")) : (x3 ? 1 : (kym_env && g.KB2 * 3 * g.R2 * 128 <= 32 * 1024));
    if (2 * g.R2 > 128 || 3 * g.R2 > 256) want_stack = 0;      // accumulator lanes; TMA box rows
    // N-concatenated operands (see the issuers): fp32-parity engine with the stacked tables; PNO_FWD_NCAT=0 switches back (tuning)
    const int want_ncat = x3 && !(getenv("PNO_FWD_NCAT") && atoi(getenv("PNO_FWD_NCAT")) == 0);
    bool fits = false;
    int best = -1;
    FwdGeom bg = g;
  for (int stack_try = want_stack; stack_try >= 0 && !fits; --stack_try) {
    g.stack = stack_try;
    g.tab_one = g.KB2 * (g.stack ? 3 : 1) * g.R2 * 128;
    // the padding rows [2 m2, N1p) of a TwT k-block are not stored when the stacked tables need the room: the MMA reads the rows
    // that follow instead, and the accumulator columns they feed are never read
    g.tw_rows = g.stack ? pad_to(2 * m2, 8) : g.N1p;
    g.ncat = want_ncat && g.stack;
    g.D1W = g.ncat ? 2 * g.tw_rows : g.N1p;
    g.tw_one = g.KB1 * g.tw_rows * 128;
    g.off_tab = 0;
    g.off_tw = pad_to((g.stack ? 1 : 2) * NT * g.tab_one, 1024);
    g.off_b2 = g.off_tw + pad_to(NT * g.tw_one, 1024);
    // M = 128 operand reads run past the R2 rows of the last table k-block: the regions that follow (TwT, B2, ring) are at
    // least 16 KB, so those reads stay inside the allocation (the rows they produce are never read back).
    // super tile = up to 8 consecutive channels (32-byte runs in the [mode][b][c] output); stage-2 batch = as many of its
    // tiles as TMEM (512 columns), the MMA (N <= 256) and shared memory (>= 3 ring stages) allow
    const int budget = 226 * 1024 - 1024;
    int want = 8 / g.IPT;
    if (want < 1) want = 1;
    for (int ns = want; ns >= 1; ns >>= 1) {
        if (C % (ns * g.IPT)) continue;
        for (int tpb = ns; tpb >= 1; tpb >>= 1) {
            for (int nb2 = 2; nb2 >= 1; --nb2)
            for (int kps = (g.KB1 <= 2 ? g.KB1 : 1); kps >= 1; kps -= (kps > 1 ? kps - 1 : 1)) {
                FwdGeom c = g;
                c.TPS = ns; c.TPB = tpb; c.nb2 = nb2; c.KPS = kps;
                c.stage_bytes = NT * kps * g.x_bytes;
                c.N2p = pad_to(tpb * g.IPT * m2, 16);
                c.nslot = 2 * g.D1W + 4 * c.N2p <= 512 ? 2 : 1;
                c.b2_one = g.KB2 * c.N2p * 128;
                c.b2_bytes = pad_to(2 * NT * c.b2_one, 1024);
                c.off_ring = g.off_b2 + nb2 * c.b2_bytes;
                c.stages = (budget - c.off_ring) / c.stage_bytes;
                if (c.stages > 8) c.stages = 8;
                if (getenv("PNO_FWD_STAGES") && atoi(getenv("PNO_FWD_STAGES")) < c.stages) c.stages = atoi(getenv("PNO_FWD_STAGES"));   // diagnostics
                if (getenv("PNO_FWD_KPS") && atoi(getenv("PNO_FWD_KPS")) != kps) continue;
                if (c.N2p > 256 || g.N1p > 256 || 2 * g.D1W + c.nslot * 2 * c.N2p > 512 || c.stages < 2) continue;
                if (g.ncat && (c.N2p > 128 || g.tw_rows > 128)) continue;        // N = 2*N2p / 2*tw_rows <= 256
                // prefer >= 16-byte output runs, then overlap (two accumulator / operand buffers), then ring depth
                const int gb = tpb * g.IPT;
                const int score = (gb >= 4 ? 1000 : 0) + (ns * g.IPT >= 8 ? 500 : 0) + c.nslot * 100 + nb2 * 10 + (kps > 1 ? 5 : 0) + (c.stages < 4 ? c.stages : 4);   // measured: whole-tile stages 130 us vs 142 us
                if (score > best) { best = score; bg = c; fits = true; }
            }
        }
    }
  }
    g = bg;
    if (!fits) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd: tables and operands do not fit in shared / tensor memory");
    g.n_super = B * C / (g.TPS * g.IPT);
    // measured at the headline shape: fp32-parity engine (2 ring stages) 191 -> 185 us with a distance of 1, single pass (3 stages)
    // 128 -> 131 / 136 us with 1 / 4: on for the former only
    g.ablate = getenv("PNO_FWD_ABLATE") ? atoi(getenv("PNO_FWD_ABLATE")) : 0;
    g.epace = getenv("PNO_FWD_EPACE") ? atoi(getenv("PNO_FWD_EPACE")) : (x3 ? 150 : 0);
    g.kym = !x3 && g.stack && kym_env;
    g.pf = getenv("PNO_FWD_PF") ? atoi(getenv("PNO_FWD_PF")) : (x3 ? 1 : 0);
    g.swap13 = !(getenv("PNO_FWD_SWAP13") && atoi(getenv("PNO_FWD_SWAP13")) == 0);
    g.smem_bytes = g.off_ring + g.stages * g.stage_bytes + 1024;
    g.dbg = g_pno_debug_buffer;
    if ((reinterpret_cast<uintptr_t>(x) & 15)) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd needs a 16-byte aligned input");
    if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched

    void* tv = nullptr;
    int rc = tc_tables_get(p, &tv);
    if (rc) return rc;
    const TcTables* t = static_cast<const TcTables*>(tv);
    CUtensorMap tmX, tmTwHi, tmTwLo, tmCHi, tmCLo, tmSHi, tmSLo;
    {
        const uint64_t dims[2] = {(uint64_t)W, (uint64_t)B * C * H};
        const uint64_t str[1] = {(uint64_t)W * 4};
        const uint32_t box[2] = {32, 128};
        if ((rc = pno_encode_tmap(&tmX, x, 2, dims, str, box, 1))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)W, (uint64_t)g.N1p};
        const uint64_t str[1] = {(uint64_t)W * 4};
        const uint32_t box[2] = {32, (uint32_t)g.tw_rows};
        if ((rc = pno_encode_tmap(&tmTwHi, t->twT_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmTwLo, t->twT_lo, 2, dims, str, box, 1))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)H, (uint64_t)g.R2};
        const uint64_t str[1] = {(uint64_t)H * 4};
        const uint32_t box[2] = {32, (uint32_t)g.R2};
        if ((rc = pno_encode_tmap(&tmCHi, t->c_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmCLo, t->c_lo, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmSHi, t->s_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmSLo, t->s_lo, 2, dims, str, box, 1))) return rc;
    }
    if (g.stack) {
        const uint64_t dims[2] = {(uint64_t)H, (uint64_t)3 * g.R2};
        const uint64_t str[1] = {(uint64_t)H * 4};
        const uint32_t box[2] = {32, (uint32_t)(3 * g.R2)};
        if ((rc = pno_encode_tmap(&tmCHi, t->t_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmCLo, t->t_lo, 2, dims, str, box, 1))) return rc;
    }
    int dev = 0, sms = 148;
    PNO_CUDA(cudaGetDevice(&dev));
    PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms = pno_usable_sms(sms);
    const int grid = g.n_super < sms ? g.n_super : sms;
    if (x3) {
        auto kern = g.ncat ? k_tc_dft_fwd<3, true> : k_tc_dft_fwd<3, false>;
        PNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        PNO_CUDA(pno_launch(kern, grid, kThreads, g.smem_bytes, st, g, (const float*)p->coef, xr, xi, tmX, tmTwHi, tmTwLo, tmCHi, tmCLo, tmSHi, tmSLo));
    } else {
        auto kern = g.kym ? k_tc_dft_fwd<1, true> : k_tc_dft_fwd<1, false>;
        PNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        PNO_CUDA(pno_launch(kern, grid, kThreads - 32, g.smem_bytes, st, g, (const float*)p->coef, xr, xi, tmX, tmTwHi, tmTwLo, tmCHi, tmCLo, tmSHi, tmSLo));
    }
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
