// Micro-benchmark: DRAM write traffic / time of a 268 MB fp32 output written with different per-instruction footprints.
//   mode 0: each warp store = 512 contiguous bytes (float4 per lane)
//   mode 1: each warp store = 4 rows x 128 B   (rows 16 KB apart)
//   mode 2: each warp store = 8 rows x 64 B
//   mode 3: each warp store = 32 rows x 16 B
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_store(float* out, int mode, size_t n_rows, int row_floats) {
    // tile = 32 rows x 32 floats (128 B per row); tiles over (row block, column block)
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int col_blocks = row_floats / 32;
    const size_t n_tiles = n_rows / 32 * col_blocks;
    for (size_t t = warp; t < n_tiles; t += n_warps) {
        const size_t rb = t / col_blocks, cb = t % col_blocks;
        float* base = out + rb * 32 * row_floats + cb * 32;
        const float4 v = make_float4((float)t, 1.f, 2.f, 3.f);
        if (mode == 0) {          // linear: treat the tile as 4 KB contiguous (different address map, same volume)
            float* lin = out + t * 1024;
            for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(lin)[i * 32 + lane] = v;
        } else if (mode == 1) {
            for (int i = 0; i < 8; ++i) { const int rr = 4 * i + (lane >> 3), ch = lane & 7;
                *reinterpret_cast<float4*>(base + (size_t)rr * row_floats + 4 * ch) = v; }
        } else if (mode == 2) {
            for (int h = 0; h < 2; ++h) for (int i = 0; i < 4; ++i) { const int rr = 8 * i + (lane >> 2), ch = lane & 3;
                *reinterpret_cast<float4*>(base + (size_t)rr * row_floats + 16 * h + 4 * ch) = v; }
        } else {
            for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(base + (size_t)lane * row_floats + 4 * c) = v;
        }
    }
}
int main() {
    const size_t n_rows = 64 * 256 * 4, row_floats = 4096 / 4 * 4 / 4;   // rows of 1024 floats? keep 268 MB total
    const size_t rows = 16384, rf = 4096;                                  // [B*C][HW] = 16384 x 4096 floats = 268 MB
    (void)n_rows; (void)row_floats;
    float* out; cudaMalloc(&out, rows * rf * 4);
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_store<<<148 * 4, 512>>>(out, mode, rows, rf);
        cudaEventRecord(e0);
        k_store<<<148 * 4, 512>>>(out, mode, rows, rf);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mode %d: %.1f us  %.0f GB/s  (%s)\n", mode, ms * 1e3, rows * rf * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
