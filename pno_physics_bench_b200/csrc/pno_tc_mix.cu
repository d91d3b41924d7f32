// Fused weight sampling + per-mode complex channel mix on the tensor cores (engines PNO_ENGINE_TC_*):
//   yhat[mode][b][o] = sum_i xhat[mode][b][i] * (mu + exp(lv/2) * eps)[mode][i][o]        (reference models.py:63-78, 33-35, 93-94)
// The sampled weights exist only as swizzled shared-memory operand tiles: eps comes from the Philox stream in-kernel, mu and
// log_var are streamed from HBM exactly once per call with coalesced 16-byte loads, nothing is written back.
// Work item = (mode, tile of 128 output channels); per item the transposed product D[o][b] is accumulated in TMEM:
//   Dre = Wre^T Xre^T - Wim^T Xim^T ,  Dim = Wim^T Xre^T + Wre^T Xim^T           (A = generated W tiles, negate-A bit; B = xhat via TMA)
//   warp 0        TMA producer: xhat tiles [B rows][32 i] (re, im) of the item's mode
//   warp 1        MMA issuer
//   warps 2-17    generators (16 warps = 4 per scheduler: the kernel is bound by Philox / Box-Muller issue slots and by the
//                 latency of the mu / log_var stream, so the loads of k-block n+1 are in flight in registers while k-block n
//                 is computed): mu/lv -> Philox/Box-Muller -> W = mu + sigma*eps -> (tf32 hi/lo) -> K-major SW128 tiles
//                 [128 o][32 i]; they also split the xhat tiles into hi/lo (3xTF32)
//   warps 18-21   epilogue: TMEM -> yhat[mode][b][o] (coalesced over o)
// The same kernel with BWD = true is the input-gradient pass  dxhat[mode][b][i] = sum_o ghat[mode][b][o] * conj(W[mode][i][o])
// (W re-sampled from the same Philox counters, never stored): the reduction runs over o, the 128-row tiles over i, the generated
// tiles are [128 i][32 o] (o is contiguous in the parameter arrays, so no transposition) and the MMA sign pattern conjugates W.
#include <stdlib.h>

// tcgen05.mma / commit predicated by an elect.sync INSIDE the asm statement instead of by a lane == 0 flag: ptxas then knows the
// instruction has one executing lane and drops its per-instruction ELECT loop.  Measured over all kernels (round 2): neutral except
// here -- fp32-parity mix 301 -> 284 us (48 MMAs per k-block) -- and the forward transform, which loses 5 us; so only this file.
#define PNO_ELECT_ASM 1
#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int kGen = 512;        // generator threads
constexpr int kThreads = 64 + kGen + 128 + 32;   // + a second MMA issuer warp (3xTF32 only: one issuer per accumulator)
// single pass: the second issuer warp is not launched.  Register cap: each scheduler owns 16 K registers, so 22-23 warps (6 on
// some schedulers) allow 80 per thread, and only <= 20 warps would allow 96 -- the generators (24 registers of prefetched
// mu / log_var) sit right at the cap: small additions to their loop spill and cost 20-50% (measured: an L2 prefetch added to
// the loop took the kernel from 206 to 245 us, 3xTF32 from 245 to 326-358 us).
constexpr int threads_for(int /*nsplit*/) { return kThreads; }
constexpr int TM = 128;          // output channels per item
constexpr int TK = 32;           // input channels per stage
constexpr int W_TILE = TM * TK * 4;      // 16 KB

struct MixGeom {
    int B, Bp, Ci, Co, m1m2, Cop, n_modes, o_tiles, n_items, KB;   // Bp = B rounded up to 16 (MMA N, TMA box rows)
    int MD;            // channels of the output ("row") side: Co forward, Ci backward; o_tiles = MD / 128, KB = (other side) / 32
    int x_tile;        // B * 128 bytes
    int stage_bytes, stages, smem_bytes;
    int raw_off, raw_bytes;       // RAW mode: two staging buffers [mu 32 KB][log_var 16 KB] behind the operand stages
    // mode window (mode-parallel sharding): the launch covers global modes [mode_first, mode_first + n_modes) of ONE weight block;
    // spectra are indexed by the LOCAL mode, the Philox counters by the GLOBAL one, the weight arrays by (mode in block - w_sub)
    int mode_first, w_sub;
    // 3xTF32: the correction terms (lo x hi, hi x lo) accumulate in their OWN tensor-memory columns and are added to the main term
    // in the epilogue with a round-to-nearest fp32 add.  The tensor core's fp32 accumulator truncates on every MMA, so the error of
    // an output grows with the number of MMAs chained into it: 64 instead of 192 for a 256-channel complex product this way
    // (measured at the cfg-3 layer: rel-L2 5.6e-6 -> see DESIGN.md).  Needs 4 * Bp <= 256 columns per buffer, i.e. batch <= 64.
    int split_acc;
    int ablate;        // measurement only (PNO_MIX_ABLATE, wrong results): 1 = no correction MMAs, 2 = no MMAs at all, 3 = no noise,
                       // 4 = no hi / lo pass over the xhat tiles, 5 = no lo stores of the weight tiles.  Round 2, headline
                       // shape: fp32-parity engine 301 -> 267 / 282 us, single pass 179 -> 178 / 166 us: the generators bound the kernel
    int trunc;         // 3xTF32 split by TRUNCATION instead of rounding (kind::tf32 ignores the 13 low mantissa bits, so the raw value is the
                       // hi operand and lo = RN_tf32(v - trunc(v)), see tf32_lo_rn): bit 0 = weight tiles, bit 1 = xhat tiles (no hi store).
                       // Default 2.  The lo part has to be rounded: left to the tensor core's own truncation the split scales every
                       // operand down by ~2e-7 (a bias, not noise) and a bias gradient of test_model_tensor_core_engines_against_oracle
                       // went from < 5e-5 to 5.6e-5.  With the rounding all settings measure the same within the box-to-box spread
                       // (274 / 274 / 276 us for 0 / 2 / 3).  Run-time branch on purpose: the same change written as the only code
                       // path made ptxas schedule the generator loop differently and the kernel 6% SLOWER (293 us, same-box A/B)
    int pf_w;          // register-prefetch mode: the producer pulls the mu / log_var tiles of the k-block pf_w ahead into L2 (0 = off)
    int iss2;          // two MMA issuer warps (one per accumulator: Dre / Dim) on different sub-partitions
    const float *mu1, *mu2, *lv1, *lv2;
    float *yr, *yi;
    PhiloxKey key;
};

__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float4 tf32_rn4(const float4 v) { return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w)); }
__device__ __forceinline__ float4 sub4(const float4 a, const float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// stage layout, NSPLIT == 1: [Wre][Wim][Xre][Xim];  NSPLIT == 3: [Wre hi][Wim hi][Wre lo][Wim lo][Xre hi][Xim hi][Xre lo][Xim lo]
// RAW (single pass, sampled): the mu / log_var tiles of a k-block are staged in shared memory by TMA, two k-blocks deep,
// instead of being prefetched into generator registers one k-block deep.  That takes 24 registers off the generators (which sit
// at the 80-register cap) and puts 96 KB per SM in flight against HBM latency instead of 48 KB.
template <int NSPLIT, bool SAMPLED, bool BWD, bool RAW>
__global__ void __launch_bounds__(threads_for(NSPLIT), 1)
k_tc_mix_fwd(const MixGeom g, const __grid_constant__ CUtensorMap tmXr, const __grid_constant__ CUtensorMap tmXi,
             const __grid_constant__ CUtensorMap tmMu1, const __grid_constant__ CUtensorMap tmMu2,
             const __grid_constant__ CUtensorMap tmLv1, const __grid_constant__ CUtensorMap tmLv2) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kMaxStages = 4;
    constexpr int NW = NSPLIT == 3 ? 4 : 2;         // W tiles per stage
    __shared__ __align__(8) uint64_t bar_raw[kMaxStages], bar_ready[kMaxStages], bar_empty[kMaxStages], bar_free[kMaxStages];
    __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
    __shared__ __align__(8) uint64_t bar_rawfull[2], bar_rawempty[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_ready[s], kGen);
            mbar_init(&bar_free[s], 1);
            mbar_init(&bar_empty[s], g.iss2 ? 2 : 1);           // one commit per issuer warp
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_tfull[s], g.iss2 ? 2 : 1);
            mbar_init(&bar_tempty[s], 128);
            mbar_init(&bar_rawfull[s], 1);
            mbar_init(&bar_rawempty[s], kGen);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmXr);
        tma_prefetch_desc(&tmXi);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                     // the spectrum written by the forward transform is complete from here on
    // Dre | Dim (columns >= B are padding: rows of the next mode / zero fill) [| Dre corrections | Dim corrections]
    const int acc_cols = (g.split_acc ? 4 : 2) * g.Bp;
    const int x_off = NW * W_TILE;                  // byte offset of the xhat tiles inside a stage

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0, phase = 0, rs = 0, rph = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
                const int mode = item / g.o_tiles;
                for (int kb = 0; kb < g.KB; ++kb) {
                    if (RAW) {
                        // weight tiles of this k-block: rows = reduction-side... forward: 32 i rows x 128 o; backward: 128 i rows x 32 o
                        const int gm = mode + g.mode_first;
                        const int ot = item - mode * g.o_tiles, blk = gm >= g.m1m2, mb = gm - blk * g.m1m2 - g.w_sub;
                        const int row0 = mb * g.Ci + (BWD ? ot * TM : kb * TK), col0 = BWD ? kb * TK : ot * TM;
                        mbar_wait(&bar_rawempty[rs], rph ^ 1);
                        unsigned char* rw = smem + g.raw_off + (size_t)rs * g.raw_bytes;
                        mbar_arrive_expect_tx(&bar_rawfull[rs], SAMPLED ? 3 * W_TILE : 2 * W_TILE);
                        tma_load_2d(rw, blk ? &tmMu2 : &tmMu1, &bar_rawfull[rs], 2 * col0, row0);
                        if (SAMPLED) tma_load_2d(rw + 2 * W_TILE, blk ? &tmLv2 : &tmLv1, &bar_rawfull[rs], col0, row0);
                        if (++rs == 2) { rs = 0; rph ^= 1; }
                    }
                    if (!RAW && g.pf_w) {
                        // the generators fetch their weights with plain loads one k-block ahead (registers); pulling the tiles of a
                        // later k-block into L2 from here turns that load's HBM latency into L2 latency
                        int pkb = kb + g.pf_w, pitem = item;
                        while (pkb >= g.KB) { pkb -= g.KB; pitem += gridDim.x; }
                        if (pitem < g.n_items) {
                            const int pmode = pitem / g.o_tiles, pot = pitem - pmode * g.o_tiles, gm = pmode + g.mode_first;
                            const int blk = gm >= g.m1m2, mb = gm - blk * g.m1m2 - g.w_sub;
                            const int row0 = mb * g.Ci + (BWD ? pot * TM : pkb * TK), col0 = BWD ? pkb * TK : pot * TM;
                            tma_prefetch_l2_2d(blk ? &tmMu2 : &tmMu1, 2 * col0, row0);
                            if (SAMPLED) tma_prefetch_l2_2d(blk ? &tmLv2 : &tmLv1, col0, row0);
                        }
                    }
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    mbar_arrive(&bar_free[stage]);                 // the generators may fill the W tiles while the xhat tiles are in flight
                    unsigned char* st = smem + (size_t)stage * g.stage_bytes + x_off;
                    mbar_arrive_expect_tx(&bar_raw[stage], 2 * g.x_tile);
                    tma_load_2d(st, &tmXr, &bar_raw[stage], kb * TK, mode * g.B);
                    tma_load_2d(st + g.x_tile, &tmXi, &bar_raw[stage], kb * TK, mode * g.B);
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == 2 + kGen / 32 + 4) {
        // ================================ MMA issuer(s) ================================
        // single pass: warp 1 issues everything.  3xTF32: 48 MMAs per k-block at >= ~53 issue cycles each outrun the generators, so
        // warp 1 issues the Dre accumulator and the extra warp the Dim accumulator.
        const bool do_re = !g.iss2 || warp == 1, do_im = !g.iss2 || warp != 1;
        if (g.iss2 || warp == 1) {   // all 32 lanes run the loop (uniform registers); `el` predicates the tcgen05 instructions
            const uint32_t el = lane == 0;
            const uint32_t idesc = make_idesc(FMT_TF32, TM, g.Bp, MAJOR_K, MAJOR_K);
            const uint32_t idesc_n = make_idesc(FMT_TF32, TM, g.Bp, MAJOR_K, MAJOR_K, 1, 0);
            const uint32_t dhi = sdesc_base(0, 16, 1024, SWIZZLE_128B).hi;
            int stage = 0, phase = 0, it = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&bar_tempty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d_re = tmem + acc * acc_cols, d_im = d_re + g.Bp;
                // correction accumulators (3xTF32 with split_acc), else the main ones
                const uint32_t c_re = g.split_acc ? d_re + 2 * g.Bp : d_re, c_im = g.split_acc ? d_im + 2 * g.Bp : d_im;
                for (int kb = 0; kb < g.KB; ++kb) {
                    mbar_wait(&bar_ready[stage], phase);
                    if (NSPLIT != 3) mbar_wait(&bar_raw[stage], phase);   // xhat tiles landed (3xTF32: the generators already waited)
                    tc_fence_after_sync();
                    const uint32_t s0 = smem_u32(smem + (size_t)stage * g.stage_bytes);
                    // low descriptor words of the operand tiles of this stage (all K-major SW128: LBO 16, SBO 1024)
                    const uint32_t wre = sdesc_base(s0, 16, 1024, SWIZZLE_128B).lo;
                    const uint32_t wim = wre + (W_TILE >> 4), wre_l = wre + 2 * (W_TILE >> 4), wim_l = wre + 3 * (W_TILE >> 4);
                    const uint32_t xre = wre + (x_off >> 4), xim = xre + (g.x_tile >> 4), xre_l = xim + (g.x_tile >> 4),
                                   xim_l = xre_l + (g.x_tile >> 4);
#pragma unroll
                    for (int ks = 0; ks < TK / 8; ++ks) {
                        const uint32_t first = (kb | ks) ? 1u : 0u;
                        const uint32_t o = ks * 2;              // 32 bytes per K step
                        // forward:  Dre = Wre Xre - Wim Xim ,  Dim = Wim Xre + Wre Xim
                        // backward: Dre = Wre Gre + Wim Gim ,  Dim = Wre Gim - Wim Gre        (conj W)
                        const uint32_t id_re2 = BWD ? idesc : idesc_n;          // sign of the Wim term of Dre
                        const uint32_t a_im1 = BWD ? wre : wim, a_im2 = BWD ? wim : wre;        // A operands of the two Dim terms
                        const uint32_t a_im1l = BWD ? wre_l : wim_l, a_im2l = BWD ? wim_l : wre_l;
                        const uint32_t b_im1 = BWD ? xim : xre, b_im2 = BWD ? xre : xim;        // and their B operands
                        const uint32_t b_im1l = BWD ? xim_l : xre_l, b_im2l = BWD ? xre_l : xim_l;
                        const uint32_t id_im2 = BWD ? idesc_n : idesc;          // sign of the second Dim term
                        if (g.ablate == 2) continue;
                        if (do_re) {
                            mma_tf32(d_re, wre + o, dhi, xre + o, dhi, idesc, first, el);
                            mma_tf32(d_re, wim + o, dhi, xim + o, dhi, id_re2, 1u, el);
                        }
                        if (do_im) {
                            mma_tf32(d_im, a_im1 + o, dhi, b_im1 + o, dhi, idesc, first, el);
                            mma_tf32(d_im, a_im2 + o, dhi, b_im2 + o, dhi, id_im2, 1u, el);
                        }
                        if (NSPLIT == 3 && g.ablate != 1) {
                            const uint32_t cfirst = g.split_acc ? first : 1u;      // first MMA into the correction accumulator
                            if (do_re) {
                                mma_tf32(c_re, wre_l + o, dhi, xre + o, dhi, idesc, cfirst, el);
                                mma_tf32(c_re, wre + o, dhi, xre_l + o, dhi, idesc, 1u, el);
                                mma_tf32(c_re, wim_l + o, dhi, xim + o, dhi, id_re2, 1u, el);
                                mma_tf32(c_re, wim + o, dhi, xim_l + o, dhi, id_re2, 1u, el);
                            }
                            if (do_im) {
                                mma_tf32(c_im, a_im1l + o, dhi, b_im1 + o, dhi, idesc, cfirst, el);
                                mma_tf32(c_im, a_im1 + o, dhi, b_im1l + o, dhi, idesc, 1u, el);
                                mma_tf32(c_im, a_im2l + o, dhi, b_im2 + o, dhi, id_im2, 1u, el);
                                mma_tf32(c_im, a_im2 + o, dhi, b_im2l + o, dhi, id_im2, 1u, el);
                            }
                        }
                    }
                    mma_commit(&bar_empty[stage], el);
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
                mma_commit(&bar_tfull[acc], el);
            }
        }
    } else if (warp < 2 + kGen / 32) {
        // ================================ generators ================================
        // thread = (o pair p, quad iq of 4 consecutive i): one Philox call per (i, o pair) feeds (re, im) of o and o+1
        const int t = threadIdx.x - 64;             // 0..511
        const PhiloxKey key = resolve_key(g.key);   // (seed, sample) from device memory under pno_set_dynamic_noise
        // forward:  p = output-channel pair (64 per tile), iq = quad of 4 consecutive i (8 per k-block)
        // backward: p = o pair inside the k-block (16),    iq = quad of 4 consecutive i rows of the tile (32)
        const int p = BWD ? (t & 15) : (t & 63), iq = BWD ? (t >> 4) : (t >> 6);
        float4 bm[4];                               // mu of (o, o+1) at the 4 i of this thread, next k-block in flight
        float2 bl[4];                               // log_var of (o, o+1)
        const float* mu_n = nullptr;                // addresses of the k-block the buffers hold / will hold
        const float* lv_n = nullptr;
        uint64_t quad_n = 0;
        auto locate = [&](int item, int kb) {
            const int lmode = item / g.o_tiles, ot = item - lmode * g.o_tiles;
            const int mode = lmode + g.mode_first;                            // global mode: block and Philox counters
            const int blk = mode >= g.m1m2;
            const int mb = mode - blk * g.m1m2;
            const int i0 = BWD ? ot * TM + iq * 4 : kb * TK + iq * 4;         // first of this thread's 4 consecutive i
            const int o0 = BWD ? kb * TK + 2 * p : ot * TM + 2 * p;           // its (even) output channel
            const size_t base = ((size_t)(mb - g.w_sub) * g.Ci + i0) * g.Co + o0;
            mu_n = (blk ? g.mu2 : g.mu1) + 2 * base;
            lv_n = SAMPLED ? (blk ? g.lv2 : g.lv1) + base : nullptr;
            quad_n = ((uint64_t)mode * g.Ci + i0) * g.Cop + (o0 >> 1);
        };
        auto fetch = [&](int j) {
            bm[j] = __ldg(reinterpret_cast<const float4*>(mu_n + (size_t)2 * j * g.Co));
            if (SAMPLED) bl[j] = __ldg(reinterpret_cast<const float2*>(lv_n + (size_t)j * g.Co));
        };
        int item_n = blockIdx.x, kb_n = 0;
        if (item_n < g.n_items) {
            locate(item_n, 0);
            if (!RAW) {
#pragma unroll
                for (int j = 0; j < 4; ++j) fetch(j);
            }
        }
        int stage = 0, phase = 0, rs = 0, rph = 0;
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            for (int kb = 0; kb < g.KB; ++kb) {
                const uint64_t quad = quad_n;       // Philox counters of the k-block held in the buffers
                bool more = true;
                if (++kb_n == g.KB) {                             // next k-block belongs to the next item: full address computation
                    kb_n = 0;
                    item_n += gridDim.x;
                    more = item_n < g.n_items;
                    if (more) locate(item_n, 0);
                } else if (BWD) {                                 // same item: the streams advance by TK output channels
                    mu_n += 2 * TK;
                    if (SAMPLED) lv_n += TK;
                    quad_n += TK / 2;
                } else {                                          // same item: the streams advance by TK input channels
                    mu_n += (size_t)2 * TK * g.Co;
                    if (SAMPLED) lv_n += (size_t)TK * g.Co;
                    quad_n += (uint64_t)TK * g.Cop;
                }
                uint32_t raw_mu = 0, raw_lv = 0;
                if (RAW) {
                    mbar_wait(&bar_rawfull[rs], rph);           // mu / log_var tiles of this k-block landed
                    const uint32_t rw = smem_u32(smem + g.raw_off + (size_t)rs * g.raw_bytes);
                    // forward: [32 i rows][128 o complex]; backward: [128 i rows][32 o complex]; this thread: rows iq*4 + j, o pair p
                    raw_mu = rw + (BWD ? iq * 4 * 256 : iq * 4 * 1024) + p * 16;
                    raw_lv = rw + 2 * W_TILE + (BWD ? iq * 4 * 128 : iq * 4 * 512) + p * 8;
                }
                mbar_wait(&bar_free[stage], phase);             // the MMAs that read this stage are done: its W tiles are ours to fill
                const uint32_t st = smem_u32(smem + (size_t)stage * g.stage_bytes);
                float wr[2][4], wi[2][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 m4 = RAW ? lds128(raw_mu + j * (BWD ? 256 : 1024)) : bm[j];
                    const float2 l2 = RAW ? (SAMPLED ? lds64(raw_lv + j * (BWD ? 128 : 512)) : make_float2(0.f, 0.f)) : bl[j];
                    if (!RAW && more) fetch(j);                  // refill the slot just consumed: next k-block's loads in flight
                    wr[0][j] = m4.x; wi[0][j] = m4.y; wr[1][j] = m4.z; wi[1][j] = m4.w;
                    if (SAMPLED && g.ablate != 3) {
                        const float4 z = philox_quad_normals(key, quad + (uint64_t)j * g.Cop);
                        const float s0 = fast_ex2(0.72134752044448170f * l2.x), s1 = fast_ex2(0.72134752044448170f * l2.y);   // exp(lv/2)
                        wr[0][j] = fmaf(s0, z.x, wr[0][j]); wi[0][j] = fmaf(s0, z.y, wi[0][j]);
                        wr[1][j] = fmaf(s1, z.z, wr[1][j]); wi[1][j] = fmaf(s1, z.w, wi[1][j]);
                    }
                }
                if (BWD) {
                    // tile rows = i (iq*4 + j), columns = o: this thread owns the 8-byte pair (o, o+1) of four rows
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int row = iq * 4 + j;
                        const uint32_t off = row * 128 + ((((p >> 1) ^ (row & 7)) << 4) | ((p & 1) << 3));
                        const float2 vr = make_float2(wr[0][j], wr[1][j]), vi = make_float2(wi[0][j], wi[1][j]);
                        if (NSPLIT == 3) {
                            const float2 hr = make_float2(tf32_rn(vr.x), tf32_rn(vr.y)), hi = make_float2(tf32_rn(vi.x), tf32_rn(vi.y));
                            sts64(st + off, hr);
                            sts64(st + W_TILE + off, hi);
                            sts64(st + 2 * W_TILE + off, make_float2(vr.x - hr.x, vr.y - hr.y));
                            sts64(st + 3 * W_TILE + off, make_float2(vi.x - hi.x, vi.y - hi.y));
                        } else {
                            sts64(st + off, vr);
                            sts64(st + W_TILE + off, vi);
                        }
                    }
                } else {
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2) {
                    const int row = 2 * p + r2;
                    const uint32_t off = row * 128 + ((iq ^ (row & 7)) << 4);
                    const float4 vr = make_float4(wr[r2][0], wr[r2][1], wr[r2][2], wr[r2][3]);
                    const float4 vi = make_float4(wi[r2][0], wi[r2][1], wi[r2][2], wi[r2][3]);
                    if (NSPLIT == 3) {
                        if (g.trunc & 1) {
                            sts128(st + off, vr);
                            sts128(st + W_TILE + off, vi);
                            sts128(st + 2 * W_TILE + off, tf32_lo_rn4(vr));
                            sts128(st + 3 * W_TILE + off, tf32_lo_rn4(vi));
                        } else {
                        const float4 hr = tf32_rn4(vr), hi = tf32_rn4(vi);
                        sts128(st + off, hr);
                        sts128(st + W_TILE + off, hi);
                        if (g.ablate != 5) {
                            sts128(st + 2 * W_TILE + off, sub4(vr, hr));
                            sts128(st + 3 * W_TILE + off, sub4(vi, hi));
                        }
                        }
                    } else {
                        sts128(st + off, vr);
                        sts128(st + W_TILE + off, vi);
                    }
                }
                }
                // ---- xhat tiles: hi in place, lo behind (layout-agnostic elementwise pass) ----
                if (NSPLIT == 3) {
                    mbar_wait(&bar_raw[stage], phase);          // xhat tiles landed
                    const uint32_t xh = st + x_off, xl = st + x_off + 2 * g.x_tile;
                    for (int i = t; i < 2 * g.x_tile / 16 && g.ablate != 4; i += kGen) {
                        const float4 v = lds128(xh + 16 * i);
                        if (g.trunc & 2) {
                            sts128(xl + 16 * i, tf32_lo_rn4(v));
                        } else {
                        const float4 h = tf32_rn4(v);
                        sts128(xh + 16 * i, h);
                        sts128(xl + 16 * i, sub4(v, h));
                        }
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(&bar_ready[stage]);
                if (RAW) {
                    mbar_arrive(&bar_rawempty[rs]);             // (the loads above completed: their values were consumed)
                    if (++rs == 2) { rs = 0; rph ^= 1; }
                }
                if (++stage == g.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue ================================
        const int q = warp & 3;
        int it = 0;
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
            const int mode = item / g.o_tiles, ot = item % g.o_tiles;
            const int acc = it & 1;
            mbar_wait(&bar_tfull[acc], (it >> 1) & 1);
            tc_fence_after_sync();
            const int o = ot * TM + q * 32 + lane;
            float* dr = g.yr + (size_t)mode * g.B * g.MD + o;
            float* di = g.yi + (size_t)mode * g.B * g.MD + o;
            const uint32_t ta = tmem_addr(tmem, q * 32, acc * acc_cols);
            for (int b0 = 0; b0 < g.B; b0 += 16) {
                float vr[16], vi[16];
                tmem_ld16(ta + b0, vr);
                tmem_ld16(ta + g.Bp + b0, vi);
                if (NSPLIT == 3 && g.split_acc) {
                    float cr[16], ci[16];
                    tmem_ld16(ta + 2 * g.Bp + b0, cr);
                    tmem_ld16(ta + 3 * g.Bp + b0, ci);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { vr[j] += cr[j]; vi[j] += ci[j]; }
                } else {
                    tmem_ld_wait();
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (b0 + j < g.B) {             // warp-uniform: skips the padding columns when B is not a multiple of 16
                        dr[(size_t)(b0 + j) * g.MD] = vr[j];
                        di[(size_t)(b0 + j) * g.MD] = vi[j];
                    }
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

// Shared launcher.  in_r / in_i: spectra [mode][b][KD] (KD = channels of the reduction side), out_r / out_i: [mode][b][MD].
static int mix_launch(const pno_plan* p, int engine, bool bwd, const float* in_r, const float* in_i, int B, int Ci, int Co,
                      const float* mu1, const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* out_r,
                      float* out_i, cudaStream_t st, int mode_first, int mode_count) {
    const int KD = bwd ? Co : Ci, MD = bwd ? Ci : Co;
    const char* who = bwd ? "tc_mix_bwd_dx" : "tc_mix_fwd";
    if (B < 1 || B > 128 || KD % TK || MD % TM || Co % 2)
        PNO_FAIL(PNO_E_UNSUPPORTED, "%s tiles 1 <= B <= 128, reduction channels %%32 == 0, output channels %%128 == 0 (got %d, %d, %d)",
                 who, B, KD, MD);
    if ((reinterpret_cast<uintptr_t>(in_r) | reinterpret_cast<uintptr_t>(in_i) | reinterpret_cast<uintptr_t>(mu1) |
         reinterpret_cast<uintptr_t>(mu2) | reinterpret_cast<uintptr_t>(lv1) | reinterpret_cast<uintptr_t>(lv2)) & 15)
        PNO_FAIL(PNO_E_UNSUPPORTED, "%s needs 16-byte aligned tensors", who);
    const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
    MixGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Bp = (B + 15) / 16 * 16; g.Ci = Ci; g.Co = Co; g.m1m2 = p->m1 * p->m2; g.Cop = (Co + 1) / 2; g.n_modes = p->M;
    g.MD = MD;
    int w_rows_modes = g.m1m2;                  // modes held by each weight array
    if (mode_count >= 0) {                      // mode window: [mode_first, mode_first + mode_count) inside one weight block
        const int blk0 = mode_first >= g.m1m2;
        if (mode_first < 0 || mode_count < 1 || mode_first + mode_count > (blk0 + 1) * g.m1m2)
            PNO_FAIL(PNO_E_INVALID, "%s: mode window [%d, %d) must lie inside one weight block of %d modes", who, mode_first,
                     mode_first + mode_count, g.m1m2);
        g.n_modes = mode_count;
        g.mode_first = mode_first;
        g.w_sub = mode_first - blk0 * g.m1m2;
        w_rows_modes = mode_count;
    }
    g.o_tiles = MD / TM; g.n_items = g.n_modes * g.o_tiles; g.KB = KD / TK;
    g.mu1 = mu1; g.mu2 = mu2; g.lv1 = lv1; g.lv2 = lv2; g.yr = out_r; g.yi = out_i; g.key = key;
    g.x_tile = g.Bp * 128;
    g.trunc = getenv("PNO_MIX_TRUNC") ? atoi(getenv("PNO_MIX_TRUNC")) : 2;
    g.ablate = getenv("PNO_MIX_ABLATE") ? atoi(getenv("PNO_MIX_ABLATE")) : 0;
    g.iss2 = x3 || !(getenv("PNO_MIX_ISS2") && atoi(getenv("PNO_MIX_ISS2")) == 0);
    g.split_acc = engine == PNO_ENGINE_TC_TF32X3 && 2 * 4 * g.Bp <= 512 && !(getenv("PNO_MIX_SPLITACC") && atoi(getenv("PNO_MIX_SPLITACC")) == 0);
    g.stage_bytes = (x3 ? 4 : 2) * (W_TILE + g.x_tile);
    g.stage_bytes = (g.stage_bytes + 1023) / 1024 * 1024;
    g.stages = getenv("PNO_MIX_STAGES") ? atoi(getenv("PNO_MIX_STAGES")) : 4;   // diagnostics override
    if (g.stages < 2 || g.stages > 4) g.stages = 4;
    while (g.stages > 2 && g.stages * g.stage_bytes + 1024 > 220 * 1024) --g.stages;
    g.smem_bytes = g.stages * g.stage_bytes + 1024;
    // RAW mode (see the kernel): single pass with sampled weights; two operand stages + two 48 KB mu / log_var staging buffers
    const bool sampled = lv1 != nullptr;
    const char* e_raw = getenv("PNO_MIX_RAW");          // tuning override: 0 = register-prefetched weights
    // (fp32-parity engine: hi / lo copies make a stage 96 KB, so the staging buffers would only fit next to ONE operand stage;
    //  measured 359 vs 287 us -- generation of k-block n+1 then waits for the MMAs of k-block n -- so it keeps the register prefetch
    //  plus the producer's L2 prefetch, pf_w)
    const bool raw = !x3 && sampled && !(e_raw && atoi(e_raw) == 0) && 2 * g.stage_bytes + 2 * 3 * W_TILE + 1024 <= 225 * 1024;
    if (raw) {
        g.stages = 2;
        g.raw_off = g.stages * g.stage_bytes;
        g.raw_bytes = 3 * W_TILE;
        g.smem_bytes = g.raw_off + 2 * g.raw_bytes + 1024;
    }
    if (g.smem_bytes > 225 * 1024) PNO_FAIL(PNO_E_UNSUPPORTED, "%s: %d bytes of shared memory needed", who, g.smem_bytes);
    if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched
    CUtensorMap tmXr, tmXi;
    int rc;
    {
        const uint64_t dims[2] = {(uint64_t)KD, (uint64_t)g.n_modes * B};
        const uint64_t str[1] = {(uint64_t)KD * 4};
        const uint32_t box[2] = {TK, (uint32_t)g.Bp};      // rows past mode*B + B: the next mode's rows, or zero fill past the end
        if ((rc = pno_encode_tmap(&tmXr, in_r, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmXi, in_i, 2, dims, str, box, 1))) return rc;
    }
    CUtensorMap tmMu1, tmMu2, tmLv1, tmLv2;
    memset(&tmMu1, 0, sizeof(tmMu1)); memset(&tmMu2, 0, sizeof(tmMu2)); memset(&tmLv1, 0, sizeof(tmLv1)); memset(&tmLv2, 0, sizeof(tmLv2));
    g.pf_w = raw ? 0 : (getenv("PNO_MIX_PFW") ? atoi(getenv("PNO_MIX_PFW")) : 1);      // fp32-parity, headline shape: 286 -> 275 us
    if (raw || g.pf_w) {      // mu: [m1m2*Ci rows][2*Co floats], log_var: [m1m2*Ci rows][Co floats]; unswizzled boxes
        const uint64_t rows = (uint64_t)w_rows_modes * Ci;
        const uint64_t dmu[2] = {(uint64_t)2 * Co, rows}, dlv[2] = {(uint64_t)Co, rows};
        const uint64_t smu[1] = {(uint64_t)2 * Co * 4}, slv[1] = {(uint64_t)Co * 4};
        const uint32_t bmu[2] = {bwd ? 2u * TK : 2u * TM, bwd ? (uint32_t)TM : (uint32_t)TK};
        const uint32_t blv[2] = {bwd ? (uint32_t)TK : (uint32_t)TM, bwd ? (uint32_t)TM : (uint32_t)TK};
        if ((rc = pno_encode_tmap(&tmMu1, mu1, 2, dmu, smu, bmu, 0))) return rc;
        if ((rc = pno_encode_tmap(&tmMu2, mu2, 2, dmu, smu, bmu, 0))) return rc;
        if (sampled) {
            if ((rc = pno_encode_tmap(&tmLv1, lv1, 2, dlv, slv, blv, 0))) return rc;
            if ((rc = pno_encode_tmap(&tmLv2, lv2, 2, dlv, slv, blv, 0))) return rc;
        }
    }
    int dev = 0, sms = 148;
    PNO_CUDA(cudaGetDevice(&dev));
    PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms = pno_usable_sms(sms);
    const int grid = g.n_items < sms ? g.n_items : sms;
#define PNO_LAUNCH_MIX(NS, SM, BW, RW)                                                                                          \
    do {                                                                                                                       \
        PNO_CUDA(cudaFuncSetAttribute(k_tc_mix_fwd<NS, SM, BW, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)); \
        PNO_CUDA(pno_launch(k_tc_mix_fwd<NS, SM, BW, RW>, grid, threads_for(NS), g.smem_bytes, st, g, tmXr, tmXi, tmMu1, tmMu2, \
                            tmLv1, tmLv2));                                                                                    \
    } while (0)
#define PNO_LAUNCH_MIX_DIR(NS, SM, RW)               \
    do {                                             \
        if (bwd) PNO_LAUNCH_MIX(NS, SM, true, RW);   \
        else PNO_LAUNCH_MIX(NS, SM, false, RW);      \
    } while (0)
    if (x3 && sampled) PNO_LAUNCH_MIX_DIR(3, true, false);
    else if (x3) PNO_LAUNCH_MIX_DIR(3, false, false);
    else if (raw) PNO_LAUNCH_MIX_DIR(1, true, true);
    else if (sampled) PNO_LAUNCH_MIX_DIR(1, true, false);
    else PNO_LAUNCH_MIX_DIR(1, false, false);
#undef PNO_LAUNCH_MIX_DIR
#undef PNO_LAUNCH_MIX
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

int tc_mix_fwd(const pno_plan* p, int engine, const float* xr, const float* xi, int B, int Ci, int Co, const float* mu1,
               const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* yr, float* yi, cudaStream_t st,
               int mode_first, int mode_count) {
    return mix_launch(p, engine, false, xr, xi, B, Ci, Co, mu1, mu2, lv1, lv2, key, yr, yi, st, mode_first, mode_count);
}

// dxhat[mode][b][i] = sum_o ghat[mode][b][o] * conj(W[mode][i][o])      (reference: autograd of models.py:33-35, 93-94)
int tc_mix_bwd_dx(const pno_plan* p, int engine, const float* gr, const float* gi, int B, int Ci, int Co, const float* mu1,
                  const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* dxr, float* dxi, cudaStream_t st,
                  int mode_first, int mode_count) {
    return mix_launch(p, engine, true, gr, gi, B, Ci, Co, mu1, mu2, lv1, lv2, key, dxr, dxi, st, mode_first, mode_count);
}
