// Local branch of the PNO block on the tensor cores (engines PNO_ENGINE_TC_*):
//   out[b][o][p] = (Wm x[b])[o][p] + bm[o] + eps * clamp(exp(0.5 clamp((Wl x[b])[o][p] + bl[o], -10, 10)), 1e-6)
// (reference models.py:298-303 + 264-275) as ONE persistent, warp-specialised kernel:
//   warp 0      TMA producer: x tile [32 i][128 p] (MN-major B operand, straight from the NCHW tensor) and the Wm / Wl tiles
//               [128 o][32 i] (K-major A operands) into a multi-stage shared-memory ring
//   warps 2-5   (fp32-parity engine only) write the bf16 copies bf16(x), bf16(x - trunc_tf32(x)) of every staged x tile
//   warp 1      one thread issues tcgen05.mma into two TMEM accumulators (mean, log-var), double buffered.  Single pass: kind::tf32.
//               fp32-parity engine ("x3b", pno_sm100.cuh): per product ONE kind::tf32 MMA on the raw fp32 tiles plus TWO kind::f16
//               MMAs on bf16 copies (the hi x lo cross terms) -- 2/3 of the tensor time and operand reads of three tf32 MMAs
//   then EW epilogue warps (8 by default = 2 per scheduler): tcgen05.ld -> bias, clamp/exp, Philox noise, (d out / d lv for
//               backward) -> each lane stores its own output row straight from registers with 256-bit stores
// A tile is (batch item b, 128 output channels, TN pixels); tiles are statically strided over the persistent grid with the
// output-channel tile fastest so that the CTAs sharing an x tile run at the same time (L2 reuse); the producer also pulls the x
// boxes of its next tile into L2 one tile ahead.
// TN = 128 by default (double-buffered accumulators).  TN = 256 (PNO_LOCAL_TN=256, single-pass mode only) halves the weight-tile
// fill per output at the price of single-buffered accumulators (mean | log-var = 512 TMEM columns); measured slower.
// Tried and dropped (round 2): the transposed product D[pixel][channel of both convs] (one N = 256 MMA per K step, x tile read
// once, whole-line stores): the MMA side is no faster (90 vs 92 us with the epilogue switched off -- at the HBM floor either way)
// and the epilogue needs per-channel biases, 16 scalar stores and a 4x4 shuffle transpose of every Philox quad: 181 vs 149 us.
// What bounds it (measured on B200, headline shape, ablations of the epilogue): without noise generation the kernel still takes
// 136 of its 164 us -- the MMA side is shared-memory-bandwidth bound (a 128x128x8 SS MMA reads 8 KB per 64 tensor cycles = the
// full 128 B/clk, and the TMA stage fills write another 48 KB per 64 KB of operand reads); TMEM loads and the global stores cost
// nothing measurable; more than 8 epilogue warps only slow the issuer down (4 warps 177 us, 8: 163, 16: 177, 24: 196).
#include <stdlib.h>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int kTileTab = 128;
// CTA: EW epilogue warps first (warp & 3 = the TMEM lane quarter a warp may read), then (fp32-parity engine only) 4 warps that convert
// the x tile, then the TMA producer and LAST the MMA issuer: the warp scheduler favours the highest warp id among the eligible warps
// of a sub-partition, and the issuer shares its sub-partition with ALU-bound epilogue warps -- as warp 1 it waited behind them for
// every instruction of its issue loop and a k-block took 1.4 k cycles while an epilogue ran against 0.8 k when none did
// (tools/trace_local.py).
constexpr int aux_warps(int nsplit) { return nsplit == 3 ? 7 : 3; }      // (4 converters +) producer + two issuers
constexpr int threads_for(int nsplit, int ew) { return (aux_warps(nsplit) + ew) * 32; }
constexpr int TM = 128;          // output channels per tile (UMMA M)
constexpr int TK = 32;           // input channels per stage (one 128-byte swizzle row)
constexpr int TILE_BYTES = TM * TK * 4;   // 16 KB, same for A [128][32] and B [32][128]

struct LocalArgs {
    int B, Ci, Co, HW;
    int noisy, want_dlv;
    int prefetch;           // L2 prefetch of the next tile's x boxes
    int tma_out;            // the epilogue stages each pass ([32 rows][16 px]) in shared memory and a TMA store writes it (see below)
    int ablate;             // measurement only (PNO_LOCAL_ABLATE, wrong results): 1 = weight tiles fetched for the first k-block of a tile only
    const float *bm, *bl;
    float *out, *dlv;
    PhiloxKey key;
    uint64_t batch_offset;
    float sign;             // EPI == 1: sign of the cross terms of the complex combine
    int n_tiles, p_tiles, o_tiles;
    long long* dbg;   // optional event trace (PNO_TRACE builds)
};

#define TRACE(role, ev) PNO_TRACE_EV(a.dbg, role, ev)

template <int NSPLIT, int TN, int EW>
struct Cfg {
    static constexpr int kXBytes = TK * TN * 4;                                  // X tile [TN/32 p-blocks][32 rows][128 B]
    static constexpr int kOperandBytes = 2 * TILE_BYTES + kXBytes;               // Am, Al, X (fp32)
    // fp32-parity engine: bf16 copies behind the fp32 operands: [Am][Al][Am lo][Al lo] (K-major, 64-byte rows, 8 KB each), then
    // [X][X lo] (MN-major, blocks of 64 pixels x 32 rows x 128 B)
    static constexpr int kW16 = TM * TK * 2;
    static constexpr int kX16 = ((TN + 63) / 64) * TK * 128;
    static constexpr int kStageBytes = kOperandBytes + (NSPLIT == 3 ? 4 * kW16 + 2 * kX16 : 0);

    // output staging of the epilogue warps (tma_out): two [32 rows][64 B] buffers per warp behind the operand ring
    static constexpr int kOutStage = 32 * 64;
    static constexpr int kOutBytes = EW == 8 ? EW * 2 * kOutStage : 0;
    static constexpr int kStages = (224 * 1024 - kOutBytes) / kStageBytes;       // 4 / 3 (TN 256) / 2 (fp32 parity)
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024;
    static constexpr int kAcc = 4 * TN <= 512 ? 2 : 1;                           // accumulator buffers in tensor memory
    static_assert(kStages >= 2, "stage does not fit");
};

// EPI == 1 (NOISY = true: two weight matrices): the H-axis pass of the two-pass pruned transforms (pno_tc_dft2.cu).  The "weights"
// are the cos / sin twiddle tables (A_m = C, A_l = S), the "pixels" of an image are its TN = 2 * m2p spectrum columns (re | im), and
// the epilogue combines the two accumulators into the complex product instead of drawing noise:
//   out[:, c] = Dm[:, c] + sign * Dl[:, c + TN/2]  (real part),   out[:, c + TN/2] = Dm[:, c + TN/2] - sign * Dl[:, c]  (imaginary part)
// CL > 1: thread-block clusters of CL CTAs work on CL consecutive pixel tiles of the SAME output-channel tile in lock step; each
// CTA fetches 1/CL of every weight tile and multicasts it to the whole cluster.  The weight tiles are re-streamed from L2 for every
// pixel tile (2 x 16 KB of fp32 + 4 x 8 KB of bf16 per 16 KB of x): without sharing, the kernel moves 9-11 TB/s out of L2 -- its
// measured bound at the headline shape, not the tensor pipe or shared memory.
// CL == 3 ("pair"): a cluster of two CTAs runs ONE MMA over both SMs (tcgen05 cta_group::2, M = 256): both CTAs work on the same
// pixel tile, the CTA of rank r on output-channel tile 2 j + r.  Each CTA loads its own weight tiles (its 128 rows of A) but only
// HALF of the x tile (64 of the 128 columns of B) and converts only that half, so per k-block the MMA reads 6 KB instead of 8 KB of
// operands per instruction and the x fill / conversion halve -- the shared-memory traffic that bounds the fp32-parity kernel drops
// from 240 to 184 KB per k-block.  The leader (rank 0) issues; all issuer-side barriers live in the leader.
template <int NSPLIT, bool NOISY, bool DLV, int TN, int EW, int EPI = 0, int CL = 1>
__global__ void __launch_bounds__(threads_for(NSPLIT, EW), 1)
k_tc_local(const LocalArgs a, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWm,
           const __grid_constant__ CUtensorMap tmWl, const __grid_constant__ CUtensorMap tmW16,
           const __grid_constant__ CUtensorMap tmOut) {
    using C = Cfg<NSPLIT, TN, EW>;
    // two MMA issuers on different sub-partitions, one per accumulator (mean conv / log-var conv): an issuer needs ~53 of the 64
    // cycles an N = 128 MMA takes just to issue it, so ONE issuer feeding both accumulators has no slack at all when the ALU-bound
    // epilogue warps of its sub-partition compete for issue / MIO slots (measured: skipping the epilogue math on the issuer's
    // sub-partition alone gave 167 -> 140 us, on any other sub-partition nothing)
    constexpr int kWarps = aux_warps(NSPLIT) + EW, kIssuer = kWarps - 2, kIssuer2 = kWarps - 1, kProducer = kWarps - 3;
    constexpr int NI = NOISY ? 2 : 1;
    constexpr bool PAIR = CL == 3;                  // 2-CTA MMA
    constexpr int CS = PAIR ? 2 : CL;               // cluster size
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_raw[C::kStages], bar_split[C::kStages], bar_empty[C::kStages];
    __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_tile[kTileTab][3];      // (ot, pt, b) of this CTA's first kTileTab tiles: no divisions in the role loops

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = a.Ci / TK;
    PNO_TRACE_DECL();

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_split[s], PAIR ? 8 : 128);          // pair: one (remote) arrival per converter warp of both CTAs, on the leader's
            mbar_init(&bar_empty[s], PAIR ? NI : CL * NI);     // one commit from every MMA issuer of the cluster that reads this CTA's stage
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_tfull[s], NI);
            mbar_init(&bar_tempty[s], PAIR ? 2 * EW : EW * 32);       // pair: one arrival per epilogue warp of both CTAs, on the leader's
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmWm);
        tma_prefetch_desc(&tmWl);
        tma_prefetch_desc(&tmW16);
    }
    // work items: "super tiles" su = (output-channel tile, group of CL consecutive pixel tiles), output-channel tile fastest; the CTA
    // of cluster rank r takes pixel tile CL * group + r
    const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int su0 = blockIdx.x / CS, su_step = gridDim.x / CS, n_super = a.n_tiles / CS;
    auto decode = [&](int su, int& ot, int& pt, int& b) {
        // pair: (pair of output-channel tiles, pixel tile); multicast cluster: (output-channel tile, CL consecutive pixel tiles)
        const int ots = PAIR ? a.o_tiles / 2 : a.o_tiles;
        ot = PAIR ? (su % ots) * 2 + crank : su % ots;
        const int ptile = PAIR ? su / ots : (su / ots) * CS + crank;
        pt = ptile % a.p_tiles;
        b = ptile / a.p_tiles;
    };
    for (int it = threadIdx.x; it < kTileTab; it += blockDim.x) decode(su0 + it * su_step, s_tile[it][0], s_tile[it][1], s_tile[it][2]);
    if (warp == kIssuer) {
        if (PAIR) tmem_alloc_pair(&tmem_slot, 512);
        else tmem_alloc(&tmem_slot, 512);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CS > 1) cluster_sync_all();                 // every CTA's barriers are initialised before any remote arrival / multicast write
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                     // x (the previous kernel's output) is complete from here on

    auto stage_ptr = [&](int s) { return smem + (size_t)s * C::kStageBytes; };

    if (warp == kProducer) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0, phase = 0, itp = 0;
            constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1);
            constexpr int kRows = PAIR ? TM : TM / CS;   // weight-tile rows fetched (and, multicast cluster, sent to the peers) by this CTA
            constexpr int kWBytes = (NOISY ? 2 : 1) * (TILE_BYTES + (NSPLIT == 3 ? 2 * C::kW16 : 0));
            for (int tile = su0; tile < n_super; tile += su_step, ++itp) {
                int ot, pt, b;
                if (itp < kTileTab) { ot = s_tile[itp][0]; pt = s_tile[itp][1]; b = s_tile[itp][2]; }
                else decode(tile, ot, pt, b);
                // x tiles of this CTA's NEXT tile are pulled into L2 one tile ahead: the stage fills then see L2 latency instead
                // of HBM latency (the ring holds only kStages stages, too few bytes in flight to cover ~1.5 us of HBM latency)
                int npt = 0, nb = 0;
                const int ntile = tile + su_step;
                const bool pf = a.prefetch && ntile < n_super;
                if (pf) {
                    int not_;
                    if (itp + 1 < kTileTab) { npt = s_tile[itp + 1][1]; nb = s_tile[itp + 1][2]; }
                    else decode(ntile, not_, npt, nb);
                }
                for (int kb = 0; kb < KB; ++kb) {
                    if (pf) tma_prefetch_l2_3d(&tmX, 0, nb * a.Ci + kb * TK, npt * (TN / 32) + (PAIR ? crank * (TN / 64) : 0));
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    TRACE(3, 1);
                    unsigned char* st = stage_ptr(stage);
                    // fp32-parity engine: the bf16 copies of the weights arrive pre-computed (k_split_weights: one bf16 array of
                    // rows [Wm][Wl][Wm lo][Wl lo], Co rows each); only the x tile is converted on the fly
                    const bool skip_w = !PAIR && a.ablate == 1 && kb > 0;
                    if (PAIR) {
                        // the weight tiles of BOTH CTAs complete on the leader's barrier (its issuer reads both CTAs' stages).  The x
                        // half: fp32-parity -> this CTA's own barrier (its converter warps wait for it and then arrive on the
                        // leader's bar_split); single pass -> the leader's barrier as well.  tmX has a half-tile box here.
                        if (NSPLIT == 3) mbar_arrive_expect_tx(&bar_raw[stage], (crank == 0 ? 2 * kWBytes : 0) + C::kXBytes / 2);
                        else if (crank == 0) mbar_arrive_expect_tx(&bar_raw[stage], 2 * (kWBytes + C::kXBytes / 2));
                        tma_load_2d_pair(st, &tmWm, &bar_raw[stage], kb * TK, ot * TM);
                        if (NOISY) tma_load_2d_pair(st + TILE_BYTES, &tmWl, &bar_raw[stage], kb * TK, ot * TM);
                        if (NSPLIT == 3) {
                            unsigned char* w16 = st + C::kOperandBytes;
                            tma_load_2d_pair(w16, &tmW16, &bar_raw[stage], kb * TK, ot * TM);
                            tma_load_2d_pair(w16 + 2 * C::kW16, &tmW16, &bar_raw[stage], kb * TK, 2 * a.Co + ot * TM);
                            if (NOISY) {
                                tma_load_2d_pair(w16 + C::kW16, &tmW16, &bar_raw[stage], kb * TK, a.Co + ot * TM);
                                tma_load_2d_pair(w16 + 3 * C::kW16, &tmW16, &bar_raw[stage], kb * TK, 3 * a.Co + ot * TM);
                            }
                            tma_load_3d(st + 2 * TILE_BYTES, &tmX, &bar_raw[stage], 0, b * a.Ci + kb * TK, pt * (TN / 32) + crank * (TN / 64));
                        } else {
                            tma_load_3d_pair(st + 2 * TILE_BYTES, &tmX, &bar_raw[stage], 0, b * a.Ci + kb * TK, pt * (TN / 32) + crank * (TN / 64));
                        }
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&bar_raw[stage], (skip_w ? 0 : kWBytes) + C::kXBytes);
                    if (skip_w) {
                    } else if (CL == 1) {
                        tma_load_2d(st, &tmWm, &bar_raw[stage], kb * TK, ot * TM);
                        if (NOISY) tma_load_2d(st + TILE_BYTES, &tmWl, &bar_raw[stage], kb * TK, ot * TM);
                        if (NSPLIT == 3) {
                            unsigned char* w16 = st + C::kOperandBytes;
                            tma_load_2d(w16, &tmW16, &bar_raw[stage], kb * TK, ot * TM);
                            tma_load_2d(w16 + 2 * C::kW16, &tmW16, &bar_raw[stage], kb * TK, 2 * a.Co + ot * TM);
                            if (NOISY) {
                                tma_load_2d(w16 + C::kW16, &tmW16, &bar_raw[stage], kb * TK, a.Co + ot * TM);
                                tma_load_2d(w16 + 3 * C::kW16, &tmW16, &bar_raw[stage], kb * TK, 3 * a.Co + ot * TM);
                            }
                        }
                    } else {
                        // this CTA's slice (rows [crank * kRows, +kRows) of every weight tile) goes to every CTA of the cluster; the
                        // other slices arrive from the peers and complete on this CTA's barrier as well (the tensor maps have
                        // kRows-row boxes)
                        const int r0 = ot * TM + crank * kRows;
                        unsigned char* wf = st + crank * (TILE_BYTES / CS);
                        tma_load_2d_mc(wf, &tmWm, &bar_raw[stage], kb * TK, r0, kMask);
                        if (NOISY) tma_load_2d_mc(wf + TILE_BYTES, &tmWl, &bar_raw[stage], kb * TK, r0, kMask);
                        if (NSPLIT == 3) {
                            unsigned char* w16 = st + C::kOperandBytes + crank * (C::kW16 / CS);
                            tma_load_2d_mc(w16, &tmW16, &bar_raw[stage], kb * TK, r0, kMask);
                            tma_load_2d_mc(w16 + 2 * C::kW16, &tmW16, &bar_raw[stage], kb * TK, 2 * a.Co + r0, kMask);
                            if (NOISY) {
                                tma_load_2d_mc(w16 + C::kW16, &tmW16, &bar_raw[stage], kb * TK, a.Co + r0, kMask);
                                tma_load_2d_mc(w16 + 3 * C::kW16, &tmW16, &bar_raw[stage], kb * TK, 3 * a.Co + r0, kMask);
                            }
                        }
                    }
                    // x viewed as (p%32, row = b*Ci + i, p/32): box {32, 32, TN/32} lands as [TN/32 p-blocks][32 rows][128 B]
                    tma_load_3d(st + 2 * TILE_BYTES, &tmX, &bar_raw[stage], 0, b * a.Ci + kb * TK, pt * (TN / 32));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kIssuer || (NOISY && warp == kIssuer2)) {
        // ================================ MMA issuers ================================
        const bool second = NOISY && warp == kIssuer2;      // this warp issues the log-var conv
        if (!PAIR || crank == 0) {   // pair: the leader issues for both CTAs
            // all 32 lanes run the loop (uniform registers); `el` predicates the tcgen05 instructions
            const uint32_t el = lane == 0;
            const uint32_t idesc = make_idesc(FMT_TF32, PAIR ? 2 * TM : TM, TN, MAJOR_K, MAJOR_MN);
            const uint32_t idesc16 = make_idesc(FMT_BF16, PAIR ? 2 * TM : TM, TN, MAJOR_K, MAJOR_MN);
            const uint32_t ahi = sdesc_base(0, 16, 1024, SWIZZLE_128B).hi;
            const uint32_t bhi = sdesc_base(0, TK * 128, 512, SWIZZLE_128B_BASE32B).hi;
            const uint32_t ahi16 = sdesc_base(0, 16, 512, SWIZZLE_64B).hi;               // bf16 K-major, 64-byte rows
            const uint32_t bhi16 = sdesc_base(0, TK * 128, 1024, SWIZZLE_128B).hi;       // bf16 MN-major, blocks of 64 pixels
            int stage = 0, phase = 0, it = 0;
            for (int tile = su0; tile < n_super; tile += su_step, ++it) {
                const int acc = it % C::kAcc;
                mbar_wait(&bar_tempty[acc], ((it / C::kAcc) & 1) ^ 1);
                TRACE(0, 1);
                tc_fence_after_sync();
                const uint32_t d_m = tmem + acc * (2 * TN) + (second ? TN : 0), d_l = tmem + acc * (2 * TN) + TN;
                // fp32-parity engine without a second conv: the cross terms accumulate in the otherwise unused second accumulator
                // and are added in the epilogue (round-to-nearest): the tensor core's fp32 accumulator truncates on every MMA, so
                // the error grows with the number of MMAs chained into one accumulator
                const uint32_t d_c = NOISY ? d_m : d_l;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&bar_raw[stage], phase);
                    // pair, fp32-parity: the peer's x half completes on the PEER's barrier (its converter warps wait there); what tells
                    // this issuer that it has landed is their arrival on bar_split -- before the first MMA, not only before the bf16 ones
                    if (PAIR && NSPLIT == 3) mbar_wait(&bar_split[stage], phase);
                    TRACE(0, 2);
                    tc_fence_after_sync();
                    const uint32_t s0 = smem_u32(stage_ptr(stage));
                    const uint32_t am = sdesc_base(s0 + (second ? TILE_BYTES : 0), 16, 1024, SWIZZLE_128B).lo;   // A tile: K-major SW128
                    const uint32_t xb = sdesc_base(s0 + 2 * TILE_BYTES, TK * 128, 512, SWIZZLE_128B_BASE32B).lo;   // B: MN-major
#pragma unroll
                    for (int ks = 0; ks < TK / 8; ++ks) {
                        const uint32_t acc_flag = (kb | ks) ? 1u : 0u;
                        const uint32_t ao = ks * 2, bo = ks * 64;    // 32 B / 1024 B per K step
                        if (PAIR) mma_tf32_pair(d_m, am + ao, ahi, xb + bo, bhi, idesc, acc_flag, el);
                        else mma_tf32(d_m, am + ao, ahi, xb + bo, bhi, idesc, acc_flag, el);
                    }
                    if (NSPLIT == 3) {
                        // cross terms on the bf16 copies (K = 16 per instruction): bf16(W) x bf16(lo x) + bf16(lo W) x bf16(x)
                        mbar_wait(&bar_split[stage], phase);
                        tc_fence_after_sync();
                        const uint32_t w16 = sdesc_base(s0 + C::kOperandBytes + (second ? C::kW16 : 0), 16, 512, SWIZZLE_64B).lo;   // [Wm][Wl][Wm lo][Wl lo]
                        const uint32_t x16 = sdesc_base(s0 + C::kOperandBytes + 4 * C::kW16, TK * 128, 1024, SWIZZLE_128B).lo;
                        const uint32_t x16l = x16 + (C::kX16 >> 4);
                        constexpr uint32_t W1 = C::kW16 >> 4;
#pragma unroll
                        for (int kh = 0; kh < TK / 16; ++kh) {
                            const uint32_t cflag = (NOISY || kb || kh) ? 1u : 0u;
                            const uint32_t ao = kh * 2, bo = kh * 128;   // 32 B / 2048 B per K = 16 step
                            if (PAIR) {
                                mma_bf16_pair(d_c, w16 + ao, ahi16, x16l + bo, bhi16, idesc16, cflag, el);
                                mma_bf16_pair(d_c, w16 + 2 * W1 + ao, ahi16, x16 + bo, bhi16, idesc16, 1u, el);
                            } else {
                                mma_bf16(d_c, w16 + ao, ahi16, x16l + bo, bhi16, idesc16, cflag, el);
                                mma_bf16(d_c, w16 + 2 * W1 + ao, ahi16, x16 + bo, bhi16, idesc16, 1u, el);
                            }
                        }
                    }
                    // frees the stage when these MMAs have read it -- in every CTA of the cluster: a peer's next multicast writes here
                    if (PAIR) mma_commit_pair(&bar_empty[stage], el);
                    else if (CL > 1) mma_commit_mc(&bar_empty[stage], (uint16_t)((1u << CL) - 1), el);
                    else mma_commit(&bar_empty[stage], el);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (PAIR) mma_commit_pair(&bar_tfull[acc], el);
                else mma_commit(&bar_tfull[acc], el);
                TRACE(0, 3);
            }
        }
    } else if (warp >= EW) {
        // ================================ bf16 copies of the x tile (fp32-parity engine) ================================
        if (NSPLIT == 3 && warp < EW + 4) {
            const int t = threadIdx.x - EW * 32;   // 0..127
            int stage = 0, phase = 0;
            for (int tile = su0; tile < n_super; tile += su_step) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&bar_raw[stage], phase);
                    const uint32_t xf = smem_u32(stage_ptr(stage)) + 2 * TILE_BYTES;            // fp32 [TN/32][32 rows][128 B]
                    const uint32_t xh = smem_u32(stage_ptr(stage)) + C::kOperandBytes + 4 * C::kW16, xl = xh + C::kX16;
#pragma unroll 4
                    for (int c = t; c < C::kXBytes / (PAIR ? 32 : 16); c += 128) {      // pair: this CTA's half of the tile
                        // 16-byte chunk c of the fp32 tile: pixel block pb (32 pixels), row r (input channel), chunk j; the fp32 layout
                        // swizzles 32-byte chunks with (r & 3), the bf16 layout 16-byte chunks (8 pixels) with (r & 7)
                        const int pb = c >> 8, r = (c >> 3) & 31, j = c & 7;
                        const int p = pb * 32 + ((((j >> 1) ^ (r & 3)) << 3) | ((j & 1) << 2));   // first of 4 consecutive pixels
                        const float4 v = lds128(xf + 16 * c);
                        uint2 full, low;
                        x3b_split4(v, full, low);
                        const uint32_t off = (uint32_t)((p >> 6) * (TK * 128) + r * 128 + (((((p & 63) >> 3) ^ (r & 7)) << 4) | ((p & 4) << 1)));
                        sts64u(xh + off, full);
                        sts64u(xl + off, low);
                    }
                    fence_proxy_async_smem();
                    if (PAIR) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&bar_split[stage]);
                    } else {
                        mbar_arrive(&bar_split[stage]);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue ================================
        // A tile is NP passes of 16 pixel columns per lane quarter; the EW / 4 warps of a quarter take the passes of the tile
        // stream round-robin (pass s = it * NP + p goes to warp s % (EW/4)), so any warp count divides the work evenly
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int kq = warp >> 2;               // index of this warp among the warps of its quarter
        const PhiloxKey key = resolve_key(a.key);
        constexpr int EWQ = EW / 4, NP = TN / 16;
        static_assert(EW % 4 == 0 && NP >= EWQ, "every epilogue warp needs a pass in every tile");
        int it = 0, base = 0;                   // base = (it * NP) % EWQ
        // tma_out: a lane's 64-byte piece of its output row goes to a swizzled [32 rows][64 B] staging buffer (two per warp) and one
        // TMA store per pass writes the 32 x 16 block: the scattered row stores (32 lines per store instruction) cost L1 / LSU
        // wavefronts that the MMA's operand reads lack (52 of 312 us in the fp32-parity kernel, 17 of 168 single pass)
        const uint32_t ostage = smem_u32(smem + (size_t)C::kStages * C::kStageBytes) + warp * 2 * C::kOutStage;
        const uint32_t orow = ostage + lane * 64, osw = (lane >> 1) & 3;
        int obuf = 0;
        for (int tile = su0; tile < n_super; tile += su_step, ++it) {
            int ot, pt, b;
            if (it < kTileTab) { ot = s_tile[it][0]; pt = s_tile[it][1]; b = s_tile[it][2]; }
            else decode(tile, ot, pt, b);
            const int acc = it % C::kAcc;
            const int o = ot * TM + q * 32 + lane;
            const float bias_m = a.bm ? a.bm[o] : 0.f;
            const float bias_l = (NOISY && a.bl) ? a.bl[o] : 0.f;
            // output row / first Philox element of this thread
            const size_t wrow = ((size_t)b * a.Co + o) * a.HW + (size_t)pt * TN;      // this lane's row
            const uint64_t e_row = ((a.batch_offset + b) * (uint64_t)a.Co + o) * (uint64_t)a.HW + (uint64_t)pt * TN;
            mbar_wait(&bar_tfull[acc], (it / C::kAcc) & 1);
            if (warp == 0) TRACE(2, 1);
            tc_fence_after_sync();
            const uint32_t t_m = tmem_addr(tmem, q * 32, acc * (2 * TN)), t_l = t_m + TN;
            if (EPI == 1) {
                constexpr int H2 = TN / 2;
                static_assert(EPI == 0 || H2 % 16 == 0, "complex combine needs 16-column halves");
#pragma unroll 1
                for (int c = kq * 16; c < H2; c += EWQ * 16) {
                    float mr[16], mi[16], lr[16], li[16];
                    tmem_ld16(t_m + c, mr);
                    tmem_ld16(t_m + H2 + c, mi);
                    tmem_ld16(t_l + c, lr);
                    tmem_ld16(t_l + H2 + c, li);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float re = fmaf(a.sign, li[j], mr[j]), im = fmaf(-a.sign, lr[j], mi[j]);
                        mr[j] = re;
                        mi[j] = im;
                    }
                    lane_store16(a.out + wrow + c, mr);
                    lane_store16(a.out + wrow + H2 + c, mi);
                }
                tc_fence_before_sync();
                mbar_arrive(&bar_tempty[acc]);
                continue;
            }
            int p0 = kq - base;
            if (p0 < 0) p0 += EWQ;
            // measurement only (wrong results): 2 = the epilogue does nothing, 3 = TMEM loads only, 4 = loads + stores (no noise / math),
            // 5 = everything but the stores
            if (a.ablate == 2) p0 = TN;
#pragma unroll 1
            for (int c = p0 * 16; c < TN; c += EWQ * 16) {
                float v[16], l[16];
                tmem_ld16(t_m + c, v);
                if (NOISY) tmem_ld16(t_l + c, l);
                if (a.ablate == 3 || a.ablate == 4 || (a.ablate == 6 && q == (kIssuer & 3)) || (a.ablate == 7 && q == ((kIssuer + 1) & 3))) {
                    tmem_ld_wait();
                    if (a.ablate == 4) lane_store16(a.out + wrow + c, v);
                    else if (v[0] == 1.2345e-30f && l[3] == 7.7e-20f) a.out[0] = v[1];
                    continue;
                }
                if (NOISY) {
                    // the noise does not depend on the accumulators: it is generated while the TMEM loads are in flight
                    float z[16];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 z4 = philox_quad_normals(key, ((e_row + c) >> 2) + k);
                        z[4 * k] = z4.x; z[4 * k + 1] = z4.y; z[4 * k + 2] = z4.z; z[4 * k + 3] = z4.w;
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float lv = l[j] + bias_l;
                        const float s_raw = fast_ex2(0.72134752044448170f * fminf(fmaxf(lv, -10.f), 10.f));
                        const float s = fmaxf(s_raw, 1e-6f);
                        v[j] = fmaf(z[j], s, v[j] + bias_m);
                        if (DLV) l[j] = ((lv >= -10.f) && (lv <= 10.f) && (s_raw >= 1e-6f)) ? 0.5f * z[j] * s : 0.f;
                    }
                } else if (NSPLIT == 3) {
                    tmem_ld16(t_l + c, l);                 // the separately accumulated 3xTF32 corrections
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = (v[j] + l[j]) + bias_m;
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += bias_m;
                }
                // each lane owns one output row: 64 contiguous bytes per pass, stored from registers as two 256-bit stores
                // (whole sectors; measured faster than a shared-memory transpose, which competes with the MMA operand reads)
                if (a.ablate == 5) {
                    if (v[0] == 1.2345e-30f && v[7] == 7.7e-20f) a.out[0] = v[1];
                    continue;
                }
                if (C::kOutBytes && a.tma_out) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the store that read this buffer two passes ago
                    __syncwarp();
                    const uint32_t dst = orow + obuf * C::kOutStage;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        sts128(dst + ((k ^ osw) << 4), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, reinterpret_cast<const void*>(smem + (size_t)C::kStages * C::kStageBytes +
                                                                            (size_t)(warp * 2 + obuf) * C::kOutStage),
                                     pt * TN + c, b * a.Co + ot * TM + q * 32);
                        bulk_commit_group();
                    }
                    obuf ^= 1;
                } else {
                    lane_store16(a.out + wrow + c, v);
                }
                if (NOISY && DLV) lane_store16(a.dlv + wrow + c, l);
            }
            base += NP % EWQ;
            if (base >= EWQ) base -= EWQ;
            tc_fence_before_sync();
            if (PAIR) {
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&bar_tempty[acc]);
            } else {
                mbar_arrive(&bar_tempty[acc]);
            }
            if (warp == 0) TRACE(2, 2);
        }
    }
    if (C::kOutBytes && a.tma_out && warp < EW && lane == 0) bulk_wait0();      // this lane's output stores are complete
    tc_fence_before_sync();
    __syncthreads();
    if (CS > 1) cluster_sync_all();                 // no CTA leaves while a peer may still multicast into it / arrive on its barriers
    if (warp == kIssuer) {
        if (PAIR) tmem_dealloc_pair(tmem, 512);
        else tmem_dealloc(tmem, 512);
    }
}

// bf16 copies of the two 1x1 weight matrices for the fp32-parity engine: out = rows [Wm][Wl][Wm lo][Wl lo], n bf16 each, where
// lo = w - trunc_tf32(w) is what the kind::tf32 main term (run on the raw fp32 weights) leaves out.  The first two are copies of
// trunc_tf32(w), not of w: the cross terms are then  h(W) l(x) + l(W) x  =  h(W) l(x) + l(W) h(x) + l(W) l(x), all three terms
// the main product misses; with bf16(w) there the l(W) l(x) term is counted twice, and because truncation parts carry the sign of
// their value that is a +1.2e-7 scale bias of every output (tests/test_split_numerics.py).
__global__ void k_split_weights(const float* __restrict__ wm, const float* __restrict__ wl, int n, uint16_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a_ = wm[i], b_ = wl ? wl[i] : 0.f;
    const uint32_t f = pack_bf16(tf32_trunc(a_), tf32_trunc(b_)), l = pack_bf16(a_ - tf32_trunc(a_), b_ - tf32_trunc(b_));
    out[i] = (uint16_t)(f & 0xFFFF);
    out[n + i] = (uint16_t)(f >> 16);
    out[2 * n + i] = (uint16_t)(l & 0xFFFF);
    out[3 * n + i] = (uint16_t)(l >> 16);
}

// Tensor maps of the two [Co][Ci] weight matrices: fp32 K-major boxes for the kind::tf32 term and -- fp32-parity engine -- one bf16
// array of rows [Wm][Wl][Wm lo][Wl lo] (written here, once per call: tiny next to the activations) for the kind::f16 cross terms.
static int weight_maps(int engine, const float* Wm, const float* Wl, int Ci, int Co, cudaStream_t st, CUtensorMap* tmWm,
                       CUtensorMap* tmWl, CUtensorMap* tmW16, const char* who, int cl = 1) {
    int rc;
    const uint64_t dims[2] = {(uint64_t)Ci, (uint64_t)Co};
    const uint64_t str[1] = {(uint64_t)Ci * 4};
    const uint32_t box[2] = {TK, (uint32_t)(TM / cl)};      // clusters: every CTA fetches (and multicasts) TM / cl rows of a tile
    if ((rc = pno_encode_tmap(tmWm, Wm, 2, dims, str, box, 1))) return rc;
    if ((rc = pno_encode_tmap(tmWl, Wl ? Wl : Wm, 2, dims, str, box, 1))) return rc;
    *tmW16 = *tmWm;                                   // unused by the single-pass engine
    if (engine == PNO_ENGINE_TC_TF32X3) {
        const size_t n = (size_t)Co * Ci;
        uint16_t* sc = reinterpret_cast<uint16_t*>(pno_scratch(0, 4 * n * 2, st));
        if (!sc) PNO_FAIL(PNO_E_CUDA, "%s: cannot allocate the bf16 weight scratch (%zu bytes)", who, 8 * n);
        k_split_weights<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Wm, Wl, (int)n, sc);
        PNO_LAUNCH_CHECK();
        const uint64_t d16[2] = {(uint64_t)Ci, (uint64_t)4 * Co};
        const uint64_t s16[1] = {(uint64_t)Ci * 2};
        if ((rc = pno_encode_tmap_ex(tmW16, sc, 2, d16, s16, box, 3, 1))) return rc;
    }
    return PNO_OK;
}

template <int NSPLIT, bool NOISY, bool DLV, int TN, int EW = 16, int EPI = 0, int CL = 1>
int launch(const LocalArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmWm, const CUtensorMap& tmWl, const CUtensorMap& tmW16,
           const CUtensorMap& tmOut, cudaStream_t st) {
    using C = Cfg<NSPLIT, TN, EW>;
    auto kern = k_tc_local<NSPLIT, NOISY, DLV, TN, EW, EPI, CL>;
    PNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    int dev = 0, sms = 148;
    PNO_CUDA(cudaGetDevice(&dev));
    PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms = pno_usable_sms(sms);
    constexpr int CS = CL == 3 ? 2 : CL;            // cluster size (CL == 3: a pair running 2-CTA MMAs)
    const int n_super = a.n_tiles / CS, groups = sms / CS;
    const int grid = (n_super < groups ? n_super : groups) * CS;
    if (CS > 1) PNO_CUDA(pno_launch_cluster(kern, grid, threads_for(NSPLIT, EW), C::kSmemBytes, st, CS, a, tmX, tmWm, tmWl, tmW16, tmOut));
    else PNO_CUDA(pno_launch(kern, grid, threads_for(NSPLIT, EW), C::kSmemBytes, st, a, tmX, tmWm, tmWl, tmW16, tmOut));
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

template <int NSPLIT, int TN, int EW, int CL = 1>
int launch_ns(const LocalArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmWm, const CUtensorMap& tmWl, const CUtensorMap& tmW16,
              const CUtensorMap& tmOut, cudaStream_t st) {
    if (!a.noisy) return launch<NSPLIT, false, false, TN, EW, 0, CL>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
    if (a.want_dlv) return launch<NSPLIT, true, true, TN, EW, 0, CL>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
    return launch<NSPLIT, true, false, TN, EW, 0, CL>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
}

}  // namespace

int tc_local_fwd(int engine, const float* x, int B, int Ci, int Co, int HW, const float* Wm, const float* bm,
                 const float* Wl, const float* bl, PhiloxKey key, uint64_t batch_offset, float* out, float* dout_dlv,
                 cudaStream_t st) {
    if (Ci % TK || Co % TM || HW % 128)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_local_fwd tiles Ci%%32 == 0, Co%%128 == 0, HW%%128 == 0 (got %d, %d, %d)", Ci, Co, HW);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(Wm) | reinterpret_cast<uintptr_t>(Wl) |
         reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout_dlv)) & 15)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_local_fwd needs 16-byte aligned tensors");
    if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout_dlv)) & 31)      // 256-bit row stores
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_local_fwd needs 32-byte aligned outputs");
    if (g_pno_dry_run) return PNO_OK;          // pno_engine_supports: the shape is tiled; nothing is launched
    LocalArgs a;
    a.B = B; a.Ci = Ci; a.Co = Co; a.HW = HW;
    a.noisy = Wl != nullptr; a.want_dlv = dout_dlv != nullptr;
    a.dbg = g_pno_debug_buffer;
    a.prefetch = getenv("PNO_LOCAL_PREFETCH") ? atoi(getenv("PNO_LOCAL_PREFETCH")) : 1;
    a.ablate = getenv("PNO_LOCAL_ABLATE") ? atoi(getenv("PNO_LOCAL_ABLATE")) : 0;
    a.bm = bm; a.bl = bl; a.out = out; a.dlv = dout_dlv; a.key = key; a.batch_offset = batch_offset; a.sign = 0.f;
    // pixel-tile width (see the header comment): 128 unless PNO_LOCAL_TN=256 asks for the wide tile
    const char* e_tn = getenv("PNO_LOCAL_TN");
    const int TN = (e_tn && atoi(e_tn) == 256 && engine != PNO_ENGINE_TC_TF32X3 && HW % 256 == 0) ? 256 : 128;
    a.o_tiles = Co / TM; a.p_tiles = HW / TN; a.n_tiles = B * a.o_tiles * a.p_tiles;
    CUtensorMap tmX, tmWm, tmWl;
    int rc;
    {   // x as (p % 32, row = b*Ci + i, p / 32)
        const uint64_t dims[3] = {32, (uint64_t)B * Ci, (uint64_t)HW / 32};
        const uint64_t str[2] = {(uint64_t)HW * 4, 128};
        const uint32_t box[3] = {32, TK, (uint32_t)TN / 32};
        if ((rc = pno_encode_tmap(&tmX, x, 3, dims, str, box, 2))) return rc;
    }
    const CUtensorMap tmXfull = tmX;
    // clusters of 2 CTAs share every weight tile through TMA multicast (see k_tc_local).  Measured at the headline shape: fp32-parity
    // engine (80 KB of weight tiles per k-block) 296 vs 308 us, single pass (32 KB) 150.5 vs 149.3 us -> on for the former only;
    // PNO_LOCAL_CL=1 / 2 forces the choice (tuning)
    const char* e_cl = getenv("PNO_LOCAL_CL");
    const int want_cl = e_cl ? atoi(e_cl) : (engine == PNO_ENGINE_TC_TF32X3 ? 2 : 1);
    // 3 = CTA pairs running 2-CTA MMAs (k_tc_local): needs an even number of output-channel tiles.  Opt-in only: measured at the
    // headline shape 171 vs 156 us (single pass) and 367 vs 282 us (fp32 parity) -- the operand traffic per SM drops as planned, but
    // every k-block now waits for a cross-CTA round trip (peer's TMA / converter warps -> leader's barrier -> MMA -> multicast
    // commit -> peer's producer) that the 2-4 stage ring does not cover, and the fp32-parity form can no longer start its tf32 MMAs
    // before the conversion of the x tile has finished
    const int cl = (want_cl == 3 && TN == 128 && a.o_tiles % 2 == 0) ? 3
                   : (TN == 128 && (B * a.p_tiles) % 2 == 0 && want_cl == 2) ? 2 : 1;
    CUtensorMap tmW16;
    if ((rc = weight_maps(engine, Wm, Wl, Ci, Co, st, &tmWm, &tmWl, &tmW16, "tc_local_fwd", cl == 2 ? 2 : 1))) return rc;
    // output through TMA stores of [32 rows][16 px] blocks (see the epilogue).  Measured at the headline shape: fp32-parity engine
    // 294 -> 280 us, single pass 153 -> 160 us (its MMA side is not shared-memory bound and the staging costs more than the row
    // stores): on for the former only; PNO_LOCAL_TMAOUT=0 / 1 forces the choice (tuning)
    CUtensorMap tmOut = tmWm;
    a.tma_out = getenv("PNO_LOCAL_TMAOUT") ? atoi(getenv("PNO_LOCAL_TMAOUT")) != 0 : engine == PNO_ENGINE_TC_TF32X3;
    if (a.tma_out) {
        const uint64_t dims[2] = {(uint64_t)HW, (uint64_t)B * Co};
        const uint64_t str[1] = {(uint64_t)HW * 4};
        const uint32_t box[2] = {16, 32};
        if ((rc = pno_encode_tmap_ex(&tmOut, out, 2, dims, str, box, 3, 0))) return rc;
    }
    // Epilogue warps: measured on B200 at the headline shape (single pass): 4 warps 177 us, 8 warps 163, 16 warps 177, 24 warps
    // 196 -- two per scheduler keep up with the tensor pipe, more only take issue slots and power from it.  PNO_LOCAL_EW=16
    // selects the wide epilogue (tuning).
    const char* e_ew = getenv("PNO_LOCAL_EW");
    const bool wide = e_ew && atoi(e_ew) == 16;
    if (cl == 3) {       // each CTA of a pair loads half of the x tile: half-tile box
        const uint64_t dims[3] = {32, (uint64_t)B * Ci, (uint64_t)HW / 32};
        const uint64_t str[2] = {(uint64_t)HW * 4, 128};
        const uint32_t box[3] = {32, TK, (uint32_t)TN / 64};
        if ((rc = pno_encode_tmap(&tmX, x, 3, dims, str, box, 2))) return rc;
        return engine == PNO_ENGINE_TC_TF32X3 ? launch_ns<3, 128, 8, 3>(a, tmX, tmWm, tmWl, tmW16, tmOut, st)
                                              : launch_ns<1, 128, 8, 3>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
    }
    (void)tmXfull;
    if (cl == 2)
        return engine == PNO_ENGINE_TC_TF32X3 ? launch_ns<3, 128, 8, 2>(a, tmX, tmWm, tmWl, tmW16, tmOut, st)
                                              : launch_ns<1, 128, 8, 2>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
    if (engine == PNO_ENGINE_TC_TF32X3)
        return wide ? launch_ns<3, 128, 16>(a, tmX, tmWm, tmWl, tmW16, tmOut, st) : launch_ns<3, 128, 8>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
    if (TN == 256) return wide ? launch_ns<1, 256, 16>(a, tmX, tmWm, tmWl, tmW16, tmOut, st) : launch_ns<1, 256, 8>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
    return wide ? launch_ns<1, 128, 16>(a, tmX, tmWm, tmWl, tmW16, tmOut, st) : launch_ns<1, 128, 8>(a, tmX, tmWm, tmWl, tmW16, tmOut, st);
}

// H-axis pass of the two-pass transforms: out[b][o][c] for c in [0, TN): complex combine (see k_tc_local, EPI == 1) of
//   Dm = Wc . x[b]  and  Dl = Ws . x[b],   x [B][Ci][TN], Wc / Ws [Co][Ci] (cos / sin tables), out [B][Co][TN].
// Ci % 32 == 0, Co % 128 == 0, TN in {64, 96, 128}.  3xTF32: the tables are split per call like the 1x1 weights.
int tc_local_combine(int engine, const float* x, int B, int Ci, int Co, int TN, const float* Wc, const float* Ws, float sign, float* out,
                     cudaStream_t st) {
    if (Ci % TK || Co % TM || !(TN == 64 || TN == 96 || TN == 128))
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_local_combine tiles Ci%%32 == 0, Co%%128 == 0, TN in {64, 96, 128} (got %d, %d, %d)", Ci, Co, TN);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(Wc) | reinterpret_cast<uintptr_t>(Ws) | reinterpret_cast<uintptr_t>(out)) & 31)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_local_combine needs 32-byte aligned tensors");
    if (g_pno_dry_run) return PNO_OK;
    LocalArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.Ci = Ci; a.Co = Co; a.HW = TN;
    a.noisy = 1; a.want_dlv = 0;
    a.dbg = g_pno_debug_buffer;
    a.prefetch = 0;
    a.ablate = 0;
    a.out = out; a.sign = sign;
    a.o_tiles = Co / TM; a.p_tiles = 1; a.n_tiles = B * a.o_tiles;
    CUtensorMap tmX, tmWm, tmWl, tmW16;
    int rc;
    {   // x as (c % 32, row = b*Ci + i, c / 32)
        const uint64_t dims[3] = {32, (uint64_t)B * Ci, (uint64_t)TN / 32};
        const uint64_t str[2] = {(uint64_t)TN * 4, 128};
        const uint32_t box[3] = {32, TK, (uint32_t)TN / 32};
        if ((rc = pno_encode_tmap(&tmX, x, 3, dims, str, box, 2))) return rc;
    }
    if ((rc = weight_maps(engine, Wc, Ws, Ci, Co, st, &tmWm, &tmWl, &tmW16, "tc_local_combine"))) return rc;
    const CUtensorMap tmOut = tmWm;              // unused: the combine epilogue stores from registers
    a.tma_out = 0;
    const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
#define PNO_COMBINE(TN_) (x3 ? launch<3, true, false, TN_, 8, 1>(a, tmX, tmWm, tmWl, tmW16, tmOut, st) \
                             : launch<1, true, false, TN_, 8, 1>(a, tmX, tmWm, tmWl, tmW16, tmOut, st))
    if (TN == 64) return PNO_COMBINE(64);
    if (TN == 96) return PNO_COMBINE(96);
    return PNO_COMBINE(128);
#undef PNO_COMBINE
}
