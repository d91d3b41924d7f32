// Micro-benchmark: issue + execution rate of back-to-back tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of N.
// One CTA per SM; operands are whatever is in shared memory (K-major SW128 descriptors).  Prints cycles per MMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../pno_physics_bench_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include "pno_sm100.cuh"
using namespace sm100;
int pno_encode_tmap(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; }

__global__ void __launch_bounds__(128, 1) k_rate(int N, int n_mma, int per_commit, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = make_idesc(FMT_TF32, 128, N, MAJOR_K, MAJOR_K);
        const uint32_t hi = sdesc_base(0, 16, 1024, SWIZZLE_128B).hi;
        const uint32_t a = sdesc_base(smem_u32(smem), 16, 1024, SWIZZLE_128B).lo;
        const uint32_t b = sdesc_base(smem_u32(smem + 16384), 16, 1024, SWIZZLE_128B).lo;
        int phase = 0;
        const long long t0 = clock64();
        long long t_issue = 0;
        for (int i = 0; i < n_mma; i += per_commit) {
            for (int j = 0; j < per_commit; ++j) mma_tf32(tmem + ((j & 1) ? 256 : 0), a + (j & 3) * 2, hi, b + (j & 3) * 2, hi, idesc, j > 1);
            mma_commit(&bar);
            if (i + per_commit >= n_mma) { t_issue = clock64() - t0; }
            mbar_wait(&bar, phase); phase ^= 1;     // wait for the batch (models a consumer-paced pipeline)
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int Ns[] = {16, 32, 48, 64, 128, 160, 256};
    for (int per_commit : {4, 32, 256}) {
        for (int N : Ns) {
            const int n = 4096;
            k_rate<<<148, 128, 100 * 1024>>>(N, n, per_commit, out);
            cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("per_commit %3d  N %3d : %.1f cycles/MMA total (floor %d)  [%s]\n", per_commit, N, (double)h[0] / n, N / 2,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
