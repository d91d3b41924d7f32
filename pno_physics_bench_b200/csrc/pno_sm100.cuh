// sm_100a building blocks used by the tensor-core engines: mbarrier, TMA (cp.async.bulk.tensor), tcgen05.mma with
// shared-memory matrix descriptors, TMEM allocation / loads.  Inline PTX only; no CUTLASS dependency.
//
// Every mbarrier wait is bounded: a wait that exceeds ~2 s of SM clock traps instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint expires) instead
// of re-issuing the test in a tight loop -- with ~20 mostly-waiting warps next to the 4-16 working ones, a hint-less spin costs
// the working warps a measurable share of the issue slots.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    // (measured: a shorter retry path -- clock read once per 256 retries -- makes the waiters spin faster and the kernels
    //  3-5% SLOWER; the clock read doubles as back-off)
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
#ifdef PNO_WAIT_SLEEP
        __nanosleep(PNO_WAIT_SLEEP);
#endif
        if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug, not a slow kernel -> fail loudly, do not hang
            printf("pno_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ---- proxies / fences --------------------------------------------------------------------------------------------------
// generic-proxy writes to shared memory -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- TMA ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// pull one box of a 3-D tensor map into L2 (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"((uint64_t)m), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"((uint64_t)m), "r"(c0), "r"(c1) : "memory");
}
// shared -> global tile store (bulk async group of the issuing thread); the shared tile must have been made visible to the
// async proxy (fence_proxy_async_smem by the writers, then a barrier) before this is issued
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk groups of this thread are complete (writes performed)
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- thread-block clusters: multicast TMA loads and cluster-wide barrier arrivals -----------------------------------------------
// A multicast load writes the box at the SAME shared-memory offset in every CTA of `mask` and signals complete_tx on the mbarrier
// at the same offset in each of them: CTAs that consume the same operand tile (e.g. one weight tile, different pixel tiles) each
// fetch 1/n of it and receive all of it -- one L2 read instead of n.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// all previously issued MMAs of this thread arrive (when complete) on the mbarrier at this offset in EVERY CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask, uint32_t elected = 1u) {
#ifdef PNO_ELECT_ASM
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
    (void)elected;
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %2, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
        "h"(mask), "r"(elected)
        : "memory");
}

// ---- CTA pairs (cta_group::2): one MMA over the two SMs of a cluster of 2 -------------------------------------------------------
// The pair's MMA has M = 256: each CTA supplies its own 128 rows of A and HALF of the N columns of B from its own shared memory (same
// offsets in both CTAs) and keeps its 128 accumulator rows in its own tensor memory.  Only the leader (cluster rank 0) issues it;
// everything the issuer waits for is signalled on the LEADER's mbarriers: a shared-window address with bit 24 cleared names the
// barrier at the same offset in the even CTA of the pair (cute::Sm100MmaPeerBitMask).
__device__ __forceinline__ uint32_t leader_bar(const uint64_t* bar) { return smem_u32(bar) & 0xFEFFFFFFu; }
__device__ __forceinline__ void mbar_arrive_leader(const uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_bar(bar)) : "memory");
}
// TMA loads into this CTA's shared memory whose bytes complete on the leader's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, const uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(leader_bar(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, const uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(leader_bar(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t ncols) {      // the same warp of BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}
// all previously issued pair MMAs of this thread arrive (when complete) on the barrier at this offset in both CTAs
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
        "mov.b16 m, 3;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(smem_u32(bar)),
        "r"(elected)
        : "memory");
}

// ---- TMEM --------------------------------------------------------------------------------------------------------------
// one full warp; ncols a power of two in [32, 512]; the base address is written to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 8 consecutive 32-bit columns: thread t of the warp gets lane (lane_base + t), columns col .. col+7
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// Up to 32 consecutive columns (n a multiple of 4) with the widest loads that cover them: x16 / x8 / x4.  Several narrow loads of
// one warp do not overlap the way independent global loads do: five x4 loads took ~1.45 k cycles in the inverse transform's
// transposer (tools/trace_inv.py), one x16 + one x4 a third of that.  v[c] = columns 4c .. 4c+3; no wait inside.
__device__ __forceinline__ void tmem_ld_cols(uint32_t ta, int n, uint32_t (&r)[8][4]) {
#define PNO_LD4(c0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"                                   \
                                 : "=r"(r[c0][0]), "=r"(r[c0][1]), "=r"(r[c0][2]), "=r"(r[c0][3]) : "r"(ta + 4 * (c0)))
#define PNO_LD8(c0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                   \
                                 : "=r"(r[c0][0]), "=r"(r[c0][1]), "=r"(r[c0][2]), "=r"(r[c0][3]), "=r"(r[c0 + 1][0]),               \
                                   "=r"(r[c0 + 1][1]), "=r"(r[c0 + 1][2]), "=r"(r[c0 + 1][3]) : "r"(ta + 4 * (c0)))
#define PNO_LD16(c0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "  \
                                  "%13, %14, %15}, [%16];"                                                                           \
                                  : "=r"(r[c0][0]), "=r"(r[c0][1]), "=r"(r[c0][2]), "=r"(r[c0][3]), "=r"(r[c0 + 1][0]),              \
                                    "=r"(r[c0 + 1][1]), "=r"(r[c0 + 1][2]), "=r"(r[c0 + 1][3]), "=r"(r[c0 + 2][0]),                  \
                                    "=r"(r[c0 + 2][1]), "=r"(r[c0 + 2][2]), "=r"(r[c0 + 2][3]), "=r"(r[c0 + 3][0]),                  \
                                    "=r"(r[c0 + 3][1]), "=r"(r[c0 + 3][2]), "=r"(r[c0 + 3][3]) : "r"(ta + 4 * (c0)))
    if (n >= 16) {
        PNO_LD16(0);
        if (n >= 32) PNO_LD16(4);
        else {
            if (n >= 24) { PNO_LD8(4); if (n >= 28) PNO_LD4(6); }
            else if (n >= 20) PNO_LD4(4);
        }
    } else {
        if (n >= 8) { PNO_LD8(0); if (n >= 12) PNO_LD4(2); }
        else PNO_LD4(0);
    }
#undef PNO_LD4
#undef PNO_LD8
#undef PNO_LD16
}
// Up to 16 consecutive columns (n a multiple of 4; more are ignored): x16, or x8 (+ x4), or x4.  No wait inside.
__device__ __forceinline__ void tmem_ld_cols16(uint32_t ta, int n, uint32_t (&r)[4][4]) {
    if (n >= 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0][0]), "=r"(r[0][1]), "=r"(r[0][2]), "=r"(r[0][3]), "=r"(r[1][0]), "=r"(r[1][1]), "=r"(r[1][2]), "=r"(r[1][3]),
                       "=r"(r[2][0]), "=r"(r[2][1]), "=r"(r[2][2]), "=r"(r[2][3]), "=r"(r[3][0]), "=r"(r[3][1]), "=r"(r[3][2]), "=r"(r[3][3])
                     : "r"(ta));
    } else if (n >= 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0][0]), "=r"(r[0][1]), "=r"(r[0][2]), "=r"(r[0][3]), "=r"(r[1][0]), "=r"(r[1][1]), "=r"(r[1][2]), "=r"(r[1][3])
                     : "r"(ta));
        if (n >= 12)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[2][0]), "=r"(r[2][1]), "=r"(r[2][2]), "=r"(r[2][3]) : "r"(ta + 8));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0][0]), "=r"(r[0][1]), "=r"(r[0][2]), "=r"(r[0][3]) : "r"(ta));
    }
}
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) {
    return base + ((uint32_t)lane << 16) + (uint32_t)col;
}

// ---- tcgen05.mma -------------------------------------------------------------------------------------------------------
enum : uint32_t { FMT_BF16 = 1, FMT_TF32 = 2 };
enum : uint32_t { MAJOR_K = 0, MAJOR_MN = 1 };
enum : uint32_t { SWIZZLE_NONE = 0, SWIZZLE_128B_BASE32B = 1, SWIZZLE_128B = 2, SWIZZLE_64B = 4, SWIZZLE_32B = 6 };

// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): fp32 accumulate, dense
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, int M, int N, uint32_t a_major, uint32_t b_major,
                                                  uint32_t a_neg = 0, uint32_t b_neg = 0) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_neg << 13) | (b_neg << 14) | (a_major << 15) | (b_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                               uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}

// Cheap per-instruction descriptors.  A single thread issues every MMA, and with K = 8 per tf32 instruction the issue loop
// is the bottleneck unless it is a handful of instructions: keep the constant high word and a base low word per operand
// and add the byte offset (>> 4) of the K step / k-block to the low word only (the 14-bit address field cannot overflow
// within 228 KB of shared memory).
struct SDesc {
    uint32_t lo, hi;
};
__device__ __forceinline__ SDesc sdesc_base(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    SDesc d;
    d.lo = ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
    d.hi = ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | ((layout_type & 7) << 29);
    return d;
}
// The whole issuer warp runs the issue loop and `elected` (1 on exactly one lane) predicates the instruction itself: with
// warp-uniform control flow ptxas keeps the descriptor arithmetic in uniform registers instead of computing it in vector
// registers and paying an R2UR round trip per operand, which is what bounds the issue rate of small (N <= 64) MMAs.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate, uint32_t elected = 1u) {
#ifdef PNO_ELECT_ASM
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
    (void)elected;
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: A is 128 lanes x 8 consecutive 32-bit columns of tensor memory (e.g. the accumulator of an
// earlier MMA; its column address must be a multiple of 4 - measured, tools/micro/mma_ts.cu), read as fp32 truncated to tf32.
// MMAs of one thread execute in issue order, so an MMA may consume the accumulator of the one issued before it.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate, uint32_t elected = 1u) {
#ifdef PNO_ELECT_ASM
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
    (void)elected;
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}

// The same for 16-bit operands (kind::f16; the instruction descriptor says bf16): K = 16 per instruction, half the shared-memory
// bytes and twice the multiply rate of kind::tf32.  Used for the cross terms of the fp32-parity engine (see pno_x3b below).
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate, uint32_t elected = 1u) {
#ifdef PNO_ELECT_ASM
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
    (void)elected;
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
// (issued by ONE thread inside a divergent branch: the self-tests)
__device__ __forceinline__ void mma_commit_1t(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar, uint32_t elected = 1u) {
#ifdef PNO_ELECT_ASM
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
    (void)elected;
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
        "r"(elected)
        : "memory");
}

// ---- canonical SWIZZLE_128B tile addressing for 4-byte elements (tile base 1024-byte aligned) ---------------------------
// K-major [R rows][K]: K split into blocks of 32 elements (128 B); block kb is a dense [R x 128 B] region.
__host__ __device__ __forceinline__ uint32_t kmajor_sw128_off(int r, int k, int R) {
    const int kb = k >> 5, kk = k & 31;
    return (uint32_t)(kb * R * 128 + r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2)));
}
// MN-major [K rows][N]: N split into blocks of 32 elements; block nb is a dense [Kt x 128 B] region (one K row per line).
// 32-bit MN-major operands only exist in the SWIZZLE_128B_BASE32B form (32-byte chunks XOR (row & 3); atom = 4 K rows):
// descriptor LBO = Kt*128 (next block of 32 N), SBO = 512 (next 4 K rows), one K=8 instruction spans 1024 bytes.
// TMA writes this layout with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B and a {32 elements, Kt rows} box.
__host__ __device__ __forceinline__ uint32_t mnmajor_sw128_off(int k, int n, int Kt) {
    const int nb = n >> 5, nn = n & 31;
    return (uint32_t)(nb * Kt * 128 + k * 128 + ((((nn >> 3) ^ (k & 3)) << 5) | ((nn & 7) << 2)));
}

// K-major "slab" form with SWIZZLE_32B: one slab per K=8 step (32 bytes per row), slab s = dense [R rows x 32 B] region
// (R a multiple of 8).  No padding of K to 32: a K=40 operand is exactly 5 slabs.  Descriptor: start = slab base,
// SBO = 256 (8 rows), layout SWIZZLE_32B.  16-byte chunk index is XORed with bit 2 of the row.
__host__ __device__ __forceinline__ uint32_t kslab_sw32_off(int r, int k, int R) {
    const int s = k >> 3, kk = k & 7;
    return (uint32_t)(s * R * 32 + r * 32 + ((((kk >> 2) ^ ((r >> 2) & 1)) << 4) | ((kk & 3) << 2)));
}


// ---- 16-bit operand tiles (bf16 cross terms of the fp32-parity engine) -------------------------------------------------------
// K-major, one k-block of 32 elements per 64-byte row, SWIZZLE_64B (16-byte chunk index XOR bits 1-2 of the row): tile
// [R rows][64 B], R a multiple of 8, tile base 512-byte aligned.  Descriptor: SBO = 512 (8 rows), layout SWIZZLE_64B; one K = 16
// instruction spans 32 bytes of the row.  TMA writes it with CU_TENSOR_MAP_SWIZZLE_64B and a {32 elements, R rows} bf16 box.
__host__ __device__ __forceinline__ uint32_t kmajor16_sw64_off(int r, int k) {
    return (uint32_t)(r * 64 + ((((k >> 3) ^ ((r >> 1) & 3)) << 4) | ((k & 7) << 1)));
}
// MN-major, 64 elements of N per 128-byte line, SWIZZLE_128B (16-byte chunk XOR (K row & 7)): block nb of 64 N is a dense
// [Kt rows x 128 B] region.  Descriptor: LBO = Kt * 128 (next block of 64 N), SBO = 1024 (next 8 K rows); one K = 16 instruction
// spans 2048 bytes.
__host__ __device__ __forceinline__ uint32_t mnmajor16_sw128_off(int k, int n, int Kt) {
    const int nb = n >> 6, nn = n & 63;
    return (uint32_t)(nb * Kt * 128 + k * 128 + ((((nn >> 3) ^ (k & 7)) << 4) | ((nn & 7) << 1)));
}

// The fp32-parity product on the tensor cores ("x3b"): with  h(v) = v truncated to tf32 (what kind::tf32 reads of a raw fp32
// operand) and  l(v) = v - h(v)  (exact, <= 13 significant bits),
//     a * b  =  h(a) h(b)  +  [ h(a) l(b) + l(a) h(b) ]  +  l(a) l(b)
// the first term is ONE kind::tf32 MMA on the raw fp32 tiles (no hi copy needed), the bracket is TWO kind::f16 MMAs on bf16 copies
// (bf16(h(a)) x bf16(l(b)) and bf16(l(a)) x bf16(b): the 8-bit mantissas cost 2^-9 relative on terms that are 2^-11 of the product,
// i.e. ~1e-6; with the FULL b in the second one the last term l(a) l(b) is included as well -- and with the full a in the first it
// would be included twice, a +1.2e-7 bias since truncation parts carry the sign of their value: k_split_weights copies h(W)).
// Against three kind::tf32 MMAs on hi / lo fp32 copies this is 2/3 of the tensor time and 2/3 of the operand bytes read from shared
// memory, for the same fp32-grade result.
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// lo part of a truncation split, itself ROUNDED to tf32: l(v) has up to 13 significant bits and kind::tf32 would truncate it again,
// always towards zero and l(v) always has the sign of v, i.e. a systematic scale-down of every operand by ~1.4e-7 (seen as a
// 5.6e-5 error on a cancellation-heavy bias gradient); rounded to nearest the lo operand is exact in tf32.  What is left is the
// round-half-up of a value with only 0-2 bits to drop (+3e-8 per operand split this way) and, where the l(a) l(b) term is dropped,
// its mean (-1.2e-7 when BOTH operands are truncation-split; zero when one of them is split by rounding, as the host-made tables
// and the mix weights are): tests/test_split_numerics.py models all the variants.
__device__ __forceinline__ float tf32_lo_rn(float v) {
    const float l = v - tf32_trunc(v);
    return __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 tf32_lo_rn4(float4 v) { return make_float4(tf32_lo_rn(v.x), tf32_lo_rn(v.y), tf32_lo_rn(v.z), tf32_lo_rn(v.w)); }
// bf16 pair (lo half = a, hi half = b), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// four consecutive fp32 -> 8 bytes of bf16(v) and 8 bytes of bf16(v - trunc_tf32(v))
__device__ __forceinline__ void x3b_split4(const float4 v, uint2& full, uint2& low) {
    full = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    low = make_uint2(pack_bf16(v.x - tf32_trunc(v.x), v.y - tf32_trunc(v.y)), pack_bf16(v.z - tf32_trunc(v.z), v.w - tf32_trunc(v.w)));
}
__device__ __forceinline__ void sts64u(uint32_t a, const uint2 v) { asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts128u(uint32_t a, const uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------
// pdl_wait(): everything the preceding kernel of the stream wrote is complete and visible after this returns (a no-op when the
// kernel was launched without the PDL attribute).  Kernels call it after their prologue (barrier init, TMEM allocation, constant
// tables) and before touching any activation / spectrum buffer; pdl_launch_dependents() lets the next kernel's CTAs be scheduled
// on SMs this grid has vacated, so its prologue overlaps this grid's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- explicit shared-space accesses ----------------------------------------------------------------------------------------
// Pointers derived from the dynamic shared-memory base are generic to the compiler (LD.E / ST.E with a window check per
// access); the hot staging loops address shared memory by its 32-bit window offset instead (LDS / STS).
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, const float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- coalesced row stores through a per-warp shared-memory transpose ------------------------------------------------------
// The 32x32b TMEM load gives lane r a run of consecutive columns of row (row0 + r).  Storing those runs directly makes every
// store instruction touch 32 different lines; passing them through a conflict-free [32 rows][16 floats] scratch turns each
// store instruction into 8 rows x 64 contiguous bytes (4x fewer L1 wavefronts).  sc = 32 x 4 float4 per warp.
__device__ __forceinline__ void warp_store_rows16(float4* sc, int lane, const float v[16], float* base, size_t row_stride) {
    const uint32_t s0 = smem_u32(sc);          // explicit shared-space accesses (a generic pointer costs a window check per access)
#pragma unroll
    for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s0 + 16u * (lane * 4 + (c ^ (lane & 3)))), "f"(v[4 * c]),
                     "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                     : "memory");
    __syncwarp();
    float* dst = base + (size_t)(lane >> 2) * row_stride + 4 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + (lane >> 2), ch = lane & 3;
        float4 t;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                     : "r"(s0 + 16u * (rr * 4 + (ch ^ (rr & 3)))));
        *reinterpret_cast<float4*>(dst) = t;
        dst += 8 * row_stride;
    }
    __syncwarp();
}

// Each lane stores its own run of 16 consecutive floats (64 B) with two 256-bit stores: every store instruction writes 32 whole
// sectors (one per lane) and no shared memory is touched.
__device__ __forceinline__ void lane_store16(float* dst, const float v[16]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8), "f"(v[8]), "f"(v[9]), "f"(v[10]),
                 "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
                 : "memory");
}

}  // namespace sm100

// ---- diagnostics (compiled in only with -DPNO_TRACE: `make EXTRA=-DPNO_TRACE`) -------------------------------------------
// Event trace of CTA 0: role r appends (cycles since kernel start << 8 | event) to dbg[4096 + r * 512 + n]
// (pno_debug_set_buffer supplies dbg; tools/trace_*.py print the merged timeline).  Production builds compile these away.
#ifdef PNO_TRACE
#define PNO_TRACE_DECL() const long long trace_t0 = clock64(); int trace_n = 0
#define PNO_TRACE_EV(dbg, role, ev)                                                                    \
    do {                                                                                               \
        if ((dbg) && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_n < 512)                      \
            (dbg)[4096 + (role) * 512 + trace_n++] = ((clock64() - trace_t0) << 8) | (long long)(ev);  \
    } while (0)
#else
#define PNO_TRACE_DECL() do {} while (0)
#define PNO_TRACE_EV(dbg, role, ev) do {} while (0)
#endif


// ---- host: launch, optionally with the programmatic-stream-serialization attribute (PNO_PDL=1) -------------------------------
// Measured on B200 at the headline shape: no gain (20.77k vs 20.75k fields/s) -- every kernel here owns a whole SM (shared memory,
// TMEM), so a dependent CTA can only start when its predecessor has left and only the ~us prologue overlaps.  Off by default.
#include <stdlib.h>
template <typename... KArgs, typename... Args>
static inline cudaError_t pno_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const bool pdl = getenv("PNO_PDL") && atoi(getenv("PNO_PDL")) == 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// the same with a thread-block cluster of `cl` CTAs along x (grid.x must be a multiple of cl)
template <typename... KArgs, typename... Args>
static inline cudaError_t pno_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cl,
                                             Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- host: tensor maps --------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
int pno_encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle /* 0 none, 1 = 128B (16-byte atoms), 2 = 128B with 32-byte atoms */);
// the same with swizzle 3 = 64B and an element type switch (bf16 = 1: 2-byte elements; dims / boxes count elements)
int pno_encode_tmap_ex(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int swizzle, int bf16);
