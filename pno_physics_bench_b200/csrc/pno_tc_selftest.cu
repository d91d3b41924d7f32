// Hardware self-test of the sm_100a building blocks (pno_sm100.cuh): one CTA computes D[128][N] = A[128][K] * B[K][N]
// with tcgen05.mma kind::tf32, operands staged either by threads (manual SWIZZLE_128B placement) or by TMA, in K-major
// or MN-major form.  tests/test_gpu_tc.py runs every combination against a CPU product; the fused kernels reuse exactly
// these descriptor conventions.
#include "pno_common.cuh"
#include "pno_sm100.cuh"

using namespace sm100;

int pno_encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle) {
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
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PNO_FAIL(PNO_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PNO_OK;
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
        mma_commit(&bar_mma);
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
