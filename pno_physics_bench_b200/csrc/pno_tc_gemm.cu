// Row GEMM with a streamed table on the tensor cores (engines PNO_ENGINE_TC_*): the W-axis passes of the TWO-PASS pruned
// transforms used for grids whose twiddle tables do not fit next to the operands of the fused kernels (pno_tc_dft_fwd.cu /
// pno_tc_dft_inv.cu): 256-row images (BASELINE config 5) and the fp32-parity engine at 128 x 128 with 32 modes (config 4).
//   C[r][n] = sum_k A[r][k] * T[n][k]          A: data rows [rows][K] (K contiguous), T: twiddle table [N][K]
//   forward pass 1 : A = x viewed [B*C*H][W],       T[j][w]  = (cos | -sin)(2 pi ky w / W)  -> T1[(b,c,h)][re ky | im ky]
//   inverse pass 2 : A = Z [(b,c,h)][re ky | im ky], T[w][j] = (cos | -sin)(2 pi ky w / W)  -> y = act(C + addend)   (models.py:97, 306)
// Both operands are K-major SWIZZLE_128B tiles staged by TMA through one ring ([128 rows x 32] of A and [N x 32] of T per stage;
// the table is re-read per tile from L2, where it lives), fp32 accumulators in tensor memory (double buffered), 3xTF32 with the
// table pre-split on the host in fp64 and the data tile split on the fly.
//   warp 0 TMA producer; warp 1 MMA issuer; warps 2-5 hi/lo splitter (3xTF32 only); then 8 epilogue warps (two per TMEM lane
//   quarter, 16-column passes dealt alternately): each lane owns one output row and stores 64-byte runs from registers.
#include <stdlib.h>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int kEW = 8;
constexpr int A_BYTES = 128 * 32 * 4;      // [128 rows][32 k] = 16 KB
constexpr int kMaxStages = 6;
constexpr int epi0(int nsplit) { return nsplit == 3 ? 6 : 2; }
constexpr int threads_for(int nsplit) { return (epi0(nsplit) + kEW) * 32; }

struct GemmArgs {
    int n_tiles, N, KB;
    int t_bytes;          // N * 128: one [N x 32] table tile
    int operand_bytes;    // A_BYTES + t_bytes
    int stage_bytes, stages;
    int epi, act;         // epi 0: out = C ; 1: v = C + addend, pre_act = v (optional), out = act(v)
    const float* addend;
    float* out;
    float* pre_act;
};

__device__ __forceinline__ float4 tf32_hi4(const float4 v) {
    float4 r;
    r.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u);
    r.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u);
    r.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u);
    r.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u);
    return r;
}

template <int NSPLIT>
__global__ void __launch_bounds__(threads_for(NSPLIT), 1)
k_tc_rowgemm(const GemmArgs a, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmT,
             const __grid_constant__ CUtensorMap tmTlo) {
    constexpr int kEpi0 = epi0(NSPLIT);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_raw[kMaxStages], bar_split[kMaxStages], bar_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_split[s], 128);
            mbar_init(&bar_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_tfull[s], 1);
            mbar_init(&bar_tempty[s], kEW * 32);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmT);
        tma_prefetch_desc(&tmTlo);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                     // A (the previous kernel's output) is complete from here on

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < a.KB; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    unsigned char* st = smem + (size_t)stage * a.stage_bytes;
                    mbar_arrive_expect_tx(&bar_raw[stage], A_BYTES + (NSPLIT == 3 ? 2 : 1) * a.t_bytes);
                    tma_load_2d(st, &tmA, &bar_raw[stage], kb * 32, tile * 128);
                    tma_load_2d(st + A_BYTES, &tmT, &bar_raw[stage], kb * 32, 0);
                    if (NSPLIT == 3) tma_load_2d(st + a.operand_bytes + A_BYTES, &tmTlo, &bar_raw[stage], kb * 32, 0);
                    if (++stage == a.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        const uint32_t el = lane == 0;
        const uint32_t idesc = make_idesc(FMT_TF32, 128, a.N, MAJOR_K, MAJOR_K);
        const uint32_t khi = sdesc_base(0, 16, 1024, SWIZZLE_128B).hi;
        int stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&bar_tempty[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d = tmem + acc * a.N;
            for (int kb = 0; kb < a.KB; ++kb) {
                mbar_wait(NSPLIT == 3 ? &bar_split[stage] : &bar_raw[stage], phase);
                tc_fence_after_sync();
                const uint32_t s0 = smem_u32(smem + (size_t)stage * a.stage_bytes);
                const uint32_t xa = sdesc_base(s0, 16, 1024, SWIZZLE_128B).lo;
                const uint32_t tb = xa + (A_BYTES >> 4);
                const uint32_t lo = a.operand_bytes >> 4;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t accf = (kb | ks) ? 1u : 0u;
                    const uint32_t o = ks * 2;
                    mma_tf32(d, xa + o, khi, tb + o, khi, idesc, accf, el);
                    if (NSPLIT == 3) {
                        mma_tf32(d, xa + lo + o, khi, tb + o, khi, idesc, 1u, el);
                        mma_tf32(d, xa + o, khi, tb + lo + o, khi, idesc, 1u, el);
                    }
                }
                mma_commit(&bar_empty[stage], el);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            mma_commit(&bar_tfull[acc], el);
        }
    } else if (warp < kEpi0) {
        // ================================ hi/lo splitter of the data tile (3xTF32) ================================
        if (NSPLIT == 3) {
            const int t = threadIdx.x - 64;
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < a.KB; ++kb) {
                    mbar_wait(&bar_raw[stage], phase);
                    const uint32_t hi = smem_u32(smem + (size_t)stage * a.stage_bytes), lo = hi + a.operand_bytes;
#pragma unroll 4
                    for (int i = t; i < A_BYTES / 16; i += 128) {
                        const float4 v = lds128(hi + 16 * i);
                        const float4 h = tf32_hi4(v);
                        sts128(hi + 16 * i, h);
                        sts128(lo + 16 * i, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&bar_split[stage]);
                    if (++stage == a.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue ================================
        const int q = warp & 3, kq = (warp - kEpi0) >> 2;       // TMEM lane quarter; which of the quarter's two warps
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const size_t row = (size_t)tile * 128 + q * 32 + lane;
            const size_t rofs = row * a.N;
            mbar_wait(&bar_tfull[acc], (it >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t ta = tmem_addr(tmem, q * 32, acc * a.N);
#pragma unroll 1
            for (int c = kq * 16; c < a.N; c += 32) {
                float v[16];
                tmem_ld16(ta + c, v);
                if (a.epi == 1) {
                    float4 a4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        a4[k] = a.addend ? __ldg(reinterpret_cast<const float4*>(a.addend + rofs + c + 4 * k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[4 * k] += a4[k].x; v[4 * k + 1] += a4[k].y; v[4 * k + 2] += a4[k].z; v[4 * k + 3] += a4[k].w;
                    }
                    if (a.pre_act) lane_store16(a.pre_act + rofs + c, v);
                    if (a.act == PNO_ACT_GELU) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = gelu_fast(v[k]);
                    } else if (a.act == PNO_ACT_RELU) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
                    }
                } else {
                    tmem_ld_wait();
                }
                lane_store16(a.out + rofs + c, v);
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

// C[rows][N] = A[rows][K] . T[N][K]^T  (+ epilogue).  rows % 128 == 0, K % 32 == 0, N % 16 == 0, 16 <= N <= 256.
// t_hi / t_lo: the table rounded to tf32 and its remainder (t_lo is only read by the 3xTF32 engine).
int tc_rowgemm(int engine, const float* A, long long rows, int K, int N, const float* t_hi, const float* t_lo, int epi, int act,
               const float* addend, float* out, float* pre_act, cudaStream_t st) {
    if (rows <= 0 || rows % 128 || K % 32 || K <= 0 || N % 16 || N < 16 || N > 256 || rows / 128 > 0x7fffffffLL)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_rowgemm tiles rows%%128 == 0, K%%32 == 0, N%%16 == 0, N <= 256 (got %lld, %d, %d)", rows, K, N);
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(addend) |
         reinterpret_cast<uintptr_t>(pre_act)) & 31)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_rowgemm needs 32-byte aligned tensors");
    const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.n_tiles = (int)(rows / 128); a.N = N; a.KB = K / 32;
    a.t_bytes = N * 128;
    a.operand_bytes = A_BYTES + a.t_bytes;
    a.stage_bytes = a.operand_bytes * (x3 ? 2 : 1);
    a.stages = (216 * 1024) / a.stage_bytes;
    if (a.stages > kMaxStages) a.stages = kMaxStages;
    if (a.stages < 2) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_rowgemm: a stage of %d bytes does not fit twice", a.stage_bytes);
    a.epi = epi; a.act = act; a.addend = addend; a.out = out; a.pre_act = pre_act;
    if (g_pno_dry_run) return PNO_OK;
    CUtensorMap tmA, tmT, tmTlo;
    int rc;
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
        const uint64_t str[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {32, 128};
        if ((rc = pno_encode_tmap(&tmA, A, 2, dims, str, box, 1))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        const uint64_t str[1] = {(uint64_t)K * 4};
        const uint32_t box[2] = {32, (uint32_t)N};
        if ((rc = pno_encode_tmap(&tmT, t_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmTlo, x3 ? t_lo : t_hi, 2, dims, str, box, 1))) return rc;
    }
    const int smem = a.stages * a.stage_bytes + 1024;
    int dev = 0, sms = 148;
    PNO_CUDA(cudaGetDevice(&dev));
    PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms = pno_usable_sms(sms);
    const int grid = a.n_tiles < sms ? a.n_tiles : sms;
    if (x3) {
        PNO_CUDA(cudaFuncSetAttribute(k_tc_rowgemm<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        PNO_CUDA(pno_launch(k_tc_rowgemm<3>, grid, threads_for(3), smem, st, a, tmA, tmT, tmTlo));
    } else {
        PNO_CUDA(cudaFuncSetAttribute(k_tc_rowgemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        PNO_CUDA(pno_launch(k_tc_rowgemm<1>, grid, threads_for(1), smem, st, a, tmA, tmT, tmTlo));
    }
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
