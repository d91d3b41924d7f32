// Micro-test: tcgen05.mma kind::tf32 with the A operand read from tensor memory (the accumulator of a previous MMA).
// Stage 1: D1[128][32] = A1[128][8] * B1[32][8]^T (both K-major SW32 slabs in shared memory, small integers -> exact).
// Stage 2: D2[128][16] = D1[:, c0 : c0+8] * B2[16][8]^T with A = tensor-memory columns c0.. of D1.
// Checks (1) the TMEM-A form works at column offsets that are not multiples of 8, (2) whether stage 2 may be issued right
// behind stage 1 without waiting for its completion (in-order execution of the MMA pipe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../pno_physics_bench_b200/csrc -o mma_ts mma_ts.cu
#include <cstdio>
#include <vector>
#include "pno_sm100.cuh"
using namespace sm100;
int pno_encode_tmap(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; }

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) k_ts(const float* a1, const float* b1, const float* b2, int c0, int wait_between,
                                                 int d1_col, float* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sA = reinterpret_cast<float*>(smem);                 // [128 rows][8]
    float* sB1 = reinterpret_cast<float*>(smem + 4096);         // [32 rows][8]
    float* sB2 = reinterpret_cast<float*>(smem + 8192);         // [16 rows][8]
    for (int i = threadIdx.x; i < 128 * 8; i += 128) sA[kslab_sw32_off(i >> 3, i & 7, 128) / 4] = a1[i];
    for (int i = threadIdx.x; i < 32 * 8; i += 128) sB1[kslab_sw32_off(i >> 3, i & 7, 32) / 4] = b1[i];
    for (int i = threadIdx.x; i < 16 * 8; i += 128) sB2[kslab_sw32_off(i >> 3, i & 7, 16) / 4] = b2[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t id1 = make_idesc(FMT_TF32, 128, 32, MAJOR_K, MAJOR_K);
        const uint32_t id2 = make_idesc(FMT_TF32, 128, 16, MAJOR_K, MAJOR_K);
        const uint32_t hi = sdesc_base(0, 16, 256, SWIZZLE_32B).hi;
        const uint32_t a = sdesc_base(smem_u32(smem), 16, 256, SWIZZLE_32B).lo;
        const uint32_t b = sdesc_base(smem_u32(smem + 4096), 16, 256, SWIZZLE_32B).lo;
        const uint32_t b2d = sdesc_base(smem_u32(smem + 8192), 16, 256, SWIZZLE_32B).lo;
        int phase = 0;
        mma_tf32(tmem + d1_col, a, hi, b, hi, id1, 0);
        if (wait_between) {
            mma_commit(&bar);
            mbar_wait(&bar, phase); phase ^= 1;
            tc_fence_after_sync();
        }
        mma_tf32_ts(tmem + 256, tmem + d1_col + c0, b2d, hi, id2, 0);
        mma_commit(&bar);
        mbar_wait(&bar, phase);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    float v[16];
    tmem_ld16(tmem_addr(tmem, warp * 32, 256), v);
    tmem_ld_wait();
    for (int k = 0; k < 16; ++k) out[(warp * 32 + lane) * 16 + k] = v[k];
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    std::vector<float> a1(128 * 8), b1(32 * 8), b2(16 * 8);
    for (int i = 0; i < 128 * 8; ++i) a1[i] = (float)((i * 7 + 3) % 11 - 5);
    for (int i = 0; i < 32 * 8; ++i) b1[i] = (float)((i * 5 + 1) % 7 - 3);
    for (int i = 0; i < 16 * 8; ++i) b2[i] = (float)((i * 3 + 2) % 5 - 2);
    float *da, *db1, *db2, *dout;
    cudaMalloc(&da, a1.size() * 4); cudaMalloc(&db1, b1.size() * 4); cudaMalloc(&db2, b2.size() * 4); cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(da, a1.data(), a1.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db1, b1.data(), b1.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db2, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    std::vector<double> d1(128 * 32);
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 32; ++n) {
            double s = 0;
            for (int k = 0; k < 8; ++k) s += (double)a1[r * 8 + k] * b1[n * 8 + k];
            d1[r * 32 + n] = s;
        }
    for (int d1_col : {0, 36})
    for (int wait_between : {1, 0})
        for (int c0 : {0, 8, 4, 12, 20, 24}) {
            cudaMemset(dout, 0xFF, 128 * 16 * 4);
            k_ts<<<1, 128, 32 * 1024>>>(da, db1, db2, c0, wait_between, d1_col, dout);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> h(128 * 16);
            cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
            double worst = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < 16; ++n) {
                    double s = 0;
                    for (int k = 0; k < 8; ++k) s += d1[r * 32 + c0 + k] * b2[n * 8 + k];
                    const double err = fabs(s - h[r * 16 + n]);
                    if (!(err <= worst)) worst = err;
                }
            printf("d1_col %2d wait %d c0 %2d : max |err| %.3g  [%s]\n", d1_col, wait_between, c0, worst, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
