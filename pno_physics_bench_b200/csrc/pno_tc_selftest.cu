// Hardware self-test of the sm_100a building blocks (pno_sm100.cuh): one CTA computes D[128][N] = A[128][K] * B[K][N]
// with tcgen05.mma kind::tf32, operands staged either by threads (manual SWIZZLE_128B placement) or by TMA, in K-major
// or MN-major form.  tests/test_gpu_tc.py runs every combination against a CPU product; the fused kernels reuse exactly
// these descriptor conventions.
#include "pno_common.cuh"
#include "pno_sm100.cuh"

using namespace sm100;

int pno_encode_tmap_ex(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int swizzle, int bf16) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        PNO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
        if (!sym || q != cudaDriverEntryPointSuccess) PNO_FAIL(PNO_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
        fn = reinterpret_cast<encode_fn>(sym);
    }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];     // strides of dims 1..rank-1
    const CUtensorMapSwizzle sw = swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                  : swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                          const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PNO_FAIL(PNO_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PNO_OK;
}

int pno_encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle) {
    return pno_encode_tmap_ex(out, base, rank, dims, strides_bytes, box, swizzle, 0);
}

namespace {

struct SelfTestArgs {
    int N, K;
    int a_mn, b_mn;        // operand majorness (0 = K-major)
    int a_tma, b_tma;      // staged by TMA instead of by threads
    int a_neg, b_neg;
    int slab;              // both operands K-major in SWIZZLE_32B slab form (staged by threads)
    const float *A, *B;    // logical A[128][K], B[K][N], row-major
    float* D;              // [128][N]
};

__global__ void __launch_bounds__(128) k_umma_selftest(const SelfTestArgs a, const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;                                   // 128*K*4 bytes
    unsigned char* sB = base + 128 * a.K * 4;                   // K*N*4 bytes (N padded to 32 for MN-major)
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = a.K, N = a.N;

    if (tid == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    // ---- stage operands -------------------------------------------------------------------------------------------
    uint32_t tx = 0;
    if (a.a_tma) tx += 128u * K * 4;
    if (a.b_tma) tx += (uint32_t)K * N * 4;
    if (tid == 0 && tx) {
        mbar_arrive_expect_tx(&bar_tma, tx);
        if (a.a_tma) {
            if (!a.a_mn) for (int kb = 0; kb < K / 32; ++kb) tma_load_2d(sA + kb * 128 * 128, &tmA, &bar_tma, kb * 32, 0);
            else         for (int mb = 0; mb < 4; ++mb)      tma_load_2d(sA + mb * K * 128, &tmA, &bar_tma, mb * 32, 0);
        }
        if (a.b_tma) {
            if (!a.b_mn) for (int kb = 0; kb < K / 32; ++kb) tma_load_2d(sB + kb * N * 128, &tmB, &bar_tma, kb * 32, 0);
            else         for (int nb = 0; nb < N / 32; ++nb) tma_load_2d(sB + nb * K * 128, &tmB, &bar_tma, nb * 32, 0);
        }
    }
    if (!a.a_tma)
        for (int e = tid; e < 128 * K; e += 128) {
            const int m = e / K, k = e % K;
            const uint32_t off = a.slab ? kslab_sw32_off(m, k, 128) : a.a_mn ? mnmajor_sw128_off(k, m, K) : kmajor_sw128_off(m, k, 128);
            *reinterpret_cast<float*>(sA + off) = a.A[e];
        }
    if (!a.b_tma)
        for (int e = tid; e < K * N; e += 128) {
            const int k = e / N, n = e % N;
            const uint32_t off = a.slab ? kslab_sw32_off(n, k, N) : a.b_mn ? mnmajor_sw128_off(k, n, K) : kmajor_sw128_off(n, k, N);
            *reinterpret_cast<float*>(sB + off) = a.B[e];
        }
    fence_proxy_async_smem();
    __syncthreads();
    if (tx) mbar_wait(&bar_tma, 0);

    // ---- MMA ------------------------------------------------------------------------------------------------------
    if (tid == 0) {
        tc_fence_after_sync();
        const uint32_t idesc = make_idesc(FMT_TF32, 128, N, a.a_mn ? MAJOR_MN : MAJOR_K, a.b_mn ? MAJOR_MN : MAJOR_K,
                                          a.a_neg, a.b_neg);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t ad, bd;
            if (a.slab) {
                ad = make_sdesc(a0 + ks * 128 * 32, 16, 256, SWIZZLE_32B);
                bd = make_sdesc(b0 + ks * N * 32, 16, 256, SWIZZLE_32B);
                mma_tf32_ss(tmem, ad, bd, idesc, ks > 0);
                continue;
            }
            if (!a.a_mn) ad = make_sdesc(a0 + (ks >> 2) * 128 * 128 + (ks & 3) * 32, 16, 1024, SWIZZLE_128B);
            else         ad = make_sdesc(a0 + ks * 1024, (uint32_t)K * 128, 512, SWIZZLE_128B_BASE32B);
            if (!a.b_mn) bd = make_sdesc(b0 + (ks >> 2) * N * 128 + (ks & 3) * 32, 16, 1024, SWIZZLE_128B);
            else         bd = make_sdesc(b0 + ks * 1024, (uint32_t)K * 128, 512, SWIZZLE_128B_BASE32B);
            mma_tf32_ss(tmem, ad, bd, idesc, ks > 0);
        }
        mma_commit_1t(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();

    // ---- read back ------------------------------------------------------------------------------------------------
    for (int c = 0; c < N; c += 8) {
        float v[8];
        tmem_ld8(tmem_addr(tmem, warp * 32, c), v);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) a.D[(size_t)tid * N + c + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// flags: bit0 A MN-major, bit1 B MN-major, bit2 A via TMA, bit3 B via TMA, bit4 negate A, bit5 negate B.
// A: [128][K] row-major (K-major source) -- for TMA + MN-major the caller passes At [K][128] in A_src_tma.
extern "C" int pno_selftest_umma(int flags, int N, int K, const float* A, const float* B, const float* A_tma_src,
                                 const float* B_tma_src, float* D, void* stream) {
    const bool slab = (flags >> 6) & 1;
    if (slab ? (N % 16 || N < 16 || N > 256 || K % 8 || K < 8 || K > 128 || (flags & 15))
             : (N % 32 || N < 32 || N > 256 || K % 32 || K < 32 || K > 128))
        PNO_FAIL(PNO_E_INVALID, "selftest: unsupported N / K / flags combination");
    int rc = pno_check_device();
    if (rc) return rc;
    SelfTestArgs a;
    a.N = N; a.K = K;
    a.a_mn = flags & 1; a.b_mn = (flags >> 1) & 1; a.a_tma = (flags >> 2) & 1; a.b_tma = (flags >> 3) & 1;
    a.a_neg = (flags >> 4) & 1; a.b_neg = (flags >> 5) & 1; a.slab = (flags >> 6) & 1;
    a.A = A; a.B = B; a.D = D;
    CUtensorMap tmA, tmB;
    memset(&tmA, 0, sizeof(tmA));
    memset(&tmB, 0, sizeof(tmB));
    if (a.a_tma) {
        // K-major: source [128 rows][K] -> box {32 k, 128 rows}; MN-major: source At [K rows][128] -> box {32 m, K rows}
        uint64_t dims[2], str[1];
        uint32_t box[2];
        if (!a.a_mn) { dims[0] = K; dims[1] = 128; str[0] = (uint64_t)K * 4; box[0] = 32; box[1] = 128; }
        else         { dims[0] = 128; dims[1] = K; str[0] = 128 * 4; box[0] = 32; box[1] = K; }
        if ((rc = pno_encode_tmap(&tmA, A_tma_src, 2, dims, str, box, a.a_mn ? 2 : 1))) return rc;
    }
    if (a.b_tma) {
        // K-major: source Bt [N rows][K] -> box {32 k, N rows}; MN-major: source B [K rows][N] -> box {32 n, K rows}
        uint64_t dims[2], str[1];
        uint32_t box[2];
        if (!a.b_mn) { dims[0] = K; dims[1] = N; str[0] = (uint64_t)K * 4; box[0] = 32; box[1] = N; }
        else         { dims[0] = N; dims[1] = K; str[0] = (uint64_t)N * 4; box[0] = 32; box[1] = K; }
        if ((rc = pno_encode_tmap(&tmB, B_tma_src, 2, dims, str, box, a.b_mn ? 2 : 1))) return rc;
    }
    const size_t smem = 1024 + (size_t)128 * K * 4 + (size_t)K * N * 4;
    PNO_CUDA(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_umma_selftest<<<1, 128, smem, as_stream(stream)>>>(a, tmA, tmB);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Self-test of the fp32-parity product "x3b" (pno_sm100.cuh): D = A * B from RAW fp32 operands as one kind::tf32 MMA on the fp32
// tiles plus two kind::f16 MMAs on bf16 copies (bf16(A) x bf16(lo(B)) and bf16(lo(A)) x bf16(B)), all into one accumulator.
//   flags bit0: B is MN-major (fp32: SWIZZLE_128B_BASE32B, bf16: SWIZZLE_128B) instead of K-major (fp32 SWIZZLE_128B, bf16 SWIZZLE_64B)
//         bit1: the bf16 tiles of A arrive by TMA (SWIZZLE_64B boxes of bf16 arrays split in global memory) instead of by threads
//         bit2: main term only (no cross terms): pins what kind::tf32 does with the low 13 mantissa bits of a raw fp32 operand
// A is always K-major.  N % 64 == 0 when B is MN-major (64-element bf16 blocks), N % 16 == 0 otherwise; K % 32 == 0, K <= 128.
// ---------------------------------------------------------------------------------------------------------------------------
namespace {

struct MixedArgs {
    int N, K, b_mn, a_tma, main_only;
    const float *A, *B;                  // A[128][K], B[K][N] row-major, raw fp32
    float* D;
};

__global__ void k_split_bf16(const float* __restrict__ src, int n, uint16_t* __restrict__ full, uint16_t* __restrict__ low) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = src[i];
    full[i] = (uint16_t)(pack_bf16(v, 0.f) & 0xFFFF);
    low[i] = (uint16_t)(pack_bf16(v - tf32_trunc(v), 0.f) & 0xFFFF);
}

__global__ void __launch_bounds__(128) k_umma_selftest_mixed(const MixedArgs a, const __grid_constant__ CUtensorMap tmAb,
                                                             const __grid_constant__ CUtensorMap tmAl) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int K = a.K, N = a.N, KB = K / 32;
    unsigned char* sA = base;                            // fp32 [KB][128][128 B]
    unsigned char* sB = sA + 128 * K * 4;                // fp32, K-major [KB][N][128 B] or MN-major [N/32][K][128 B]
    unsigned char* sAb = sB + K * N * 4;                 // bf16(A)      [KB][128][64 B]
    unsigned char* sAl = sAb + 128 * K * 2;              // bf16(lo(A))
    unsigned char* sBb = sAl + 128 * K * 2;              // bf16(B): K-major [KB][N][64 B] or MN-major [N/64][K][128 B]
    unsigned char* sBl = sBb + K * N * 2;
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (a.a_tma && tid == 0) {
        mbar_arrive_expect_tx(&bar_tma, 2u * 128 * K * 2);
        for (int kb = 0; kb < KB; ++kb) {
            tma_load_2d(sAb + kb * 128 * 64, &tmAb, &bar_tma, kb * 32, 0);
            tma_load_2d(sAl + kb * 128 * 64, &tmAl, &bar_tma, kb * 32, 0);
        }
    }
    for (int e = tid; e < 128 * K; e += 128) {
        const int m = e / K, k = e % K;
        const float v = a.A[e];
        *reinterpret_cast<float*>(sA + kmajor_sw128_off(m, k, 128)) = v;
        if (!a.a_tma) {
            const uint32_t off = (k >> 5) * 128 * 64 + kmajor16_sw64_off(m, k & 31);
            *reinterpret_cast<uint16_t*>(sAb + off) = (uint16_t)(pack_bf16(v, 0.f) & 0xFFFF);
            *reinterpret_cast<uint16_t*>(sAl + off) = (uint16_t)(pack_bf16(v - tf32_trunc(v), 0.f) & 0xFFFF);
        }
    }
    for (int e = tid; e < K * N; e += 128) {
        const int k = e / N, n = e % N;
        const float v = a.B[e];
        *reinterpret_cast<float*>(sB + (a.b_mn ? mnmajor_sw128_off(k, n, K) : kmajor_sw128_off(n, k, N))) = v;
        const uint32_t off = a.b_mn ? mnmajor16_sw128_off(k, n, K) : (uint32_t)((k >> 5) * N * 64) + kmajor16_sw64_off(n, k & 31);
        *reinterpret_cast<uint16_t*>(sBb + off) = (uint16_t)(pack_bf16(v, 0.f) & 0xFFFF);
        *reinterpret_cast<uint16_t*>(sBl + off) = (uint16_t)(pack_bf16(v - tf32_trunc(v), 0.f) & 0xFFFF);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (a.a_tma) mbar_wait(&bar_tma, 0);

    if (tid == 0) {
        tc_fence_after_sync();
        const uint32_t id32 = make_idesc(FMT_TF32, 128, N, MAJOR_K, a.b_mn ? MAJOR_MN : MAJOR_K);
        const uint32_t id16 = make_idesc(FMT_BF16, 128, N, MAJOR_K, a.b_mn ? MAJOR_MN : MAJOR_K);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t ad = make_sdesc(a0 + (ks >> 2) * 128 * 128 + (ks & 3) * 32, 16, 1024, SWIZZLE_128B);
            const uint64_t bd = a.b_mn ? make_sdesc(b0 + ks * 1024, (uint32_t)K * 128, 512, SWIZZLE_128B_BASE32B)
                                       : make_sdesc(b0 + (ks >> 2) * N * 128 + (ks & 3) * 32, 16, 1024, SWIZZLE_128B);
            mma_tf32_ss(tmem, ad, bd, id32, ks > 0);
        }
        if (!a.main_only) {
            for (int t = 0; t < 2; ++t) {            // t = 0: bf16(A) x bf16(lo B);  t = 1: bf16(lo A) x bf16(B)
                const uint32_t a16 = smem_u32(t ? sAl : sAb), b16 = smem_u32(t ? sBb : sBl);
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint64_t ad = make_sdesc(a16 + (ks >> 1) * 128 * 64 + (ks & 1) * 32, 16, 512, SWIZZLE_64B);
                    const uint64_t bd = a.b_mn ? make_sdesc(b16 + ks * 2048, (uint32_t)K * 128, 1024, SWIZZLE_128B)
                                               : make_sdesc(b16 + (ks >> 1) * N * 64 + (ks & 1) * 32, 16, 512, SWIZZLE_64B);
                    mma_bf16_ss(tmem, ad, bd, id16, 1u);
                }
            }
        }
        mma_commit_1t(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        tmem_ld8(tmem_addr(tmem, warp * 32, c), v);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) a.D[(size_t)tid * N + c + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int pno_selftest_umma_mixed(int flags, int N, int K, const float* A, const float* B, float* D, void* scratch, void* stream) {
    const int b_mn = flags & 1, a_tma = (flags >> 1) & 1, main_only = (flags >> 2) & 1;
    if (K % 32 || K < 32 || K > 128 || N < 16 || N > 256 || N % (b_mn ? 64 : 16))
        PNO_FAIL(PNO_E_INVALID, "selftest (mixed): unsupported N / K / flags combination");
    int rc = pno_check_device();
    if (rc) return rc;
    MixedArgs a;
    a.N = N; a.K = K; a.b_mn = b_mn; a.a_tma = a_tma; a.main_only = main_only; a.A = A; a.B = B; a.D = D;
    CUtensorMap tmAb, tmAl;
    memset(&tmAb, 0, sizeof(tmAb));
    memset(&tmAl, 0, sizeof(tmAl));
    if (a_tma) {      // scratch: 2 * 128 * K bf16
        if (!scratch) PNO_FAIL(PNO_E_INVALID, "selftest (mixed): the TMA form needs a scratch buffer of 512 * K bytes");
        uint16_t* full = static_cast<uint16_t*>(scratch);
        uint16_t* low = full + 128 * K;
        k_split_bf16<<<(128 * K + 255) / 256, 256, 0, as_stream(stream)>>>(A, 128 * K, full, low);
        PNO_LAUNCH_CHECK();
        const uint64_t dims[2] = {(uint64_t)K, 128}, str[1] = {(uint64_t)K * 2};
        const uint32_t box[2] = {32, 128};
        if ((rc = pno_encode_tmap_ex(&tmAb, full, 2, dims, str, box, 3, 1))) return rc;
        if ((rc = pno_encode_tmap_ex(&tmAl, low, 2, dims, str, box, 3, 1))) return rc;
    }
    const size_t smem = 1024 + (size_t)(128 * K + K * N) * 4 + (size_t)(128 * K + K * N) * 4;
    PNO_CUDA(cudaFuncSetAttribute(k_umma_selftest_mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    if (smem > 220 * 1024) PNO_FAIL(PNO_E_INVALID, "selftest (mixed): operands do not fit in shared memory");
    k_umma_selftest_mixed<<<1, 128, smem, as_stream(stream)>>>(a, tmAb, tmAl);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
