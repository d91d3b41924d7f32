// Pruned forward / inverse 2-D transforms on CUDA cores (engine PNO_ENGINE_FP32).
//
// Replaces torch.fft.rfft2 + mode slicing (reference models.py:89, 93-94) and zeros + scatter + torch.fft.irfft2
// (models.py:92-97): only the 2*m1 x m2 retained modes are ever computed.  One CTA owns G consecutive channels of one
// batch item, runs the separable transform per image in shared memory (row pass over W, column pass over H) and moves
// the truncated spectrum in the [mode][b][c] layout with G-wide contiguous segments.
#include "pno_common.cuh"

namespace {

constexpr int kThreads = 256;

struct DftGeom {
    int H, W, m1, m2, M;
    int Hc;   // rows per chunk
    int G;    // images per CTA
    int wpad; // W + 1
};

__host__ __device__ inline size_t dft_smem_bytes(const DftGeom& g) {
    size_t b = 0;
    b += sizeof(float2) * (g.H + g.W);            // tables
    b += sizeof(int) * 2 * g.m1;                  // kx
    b += sizeof(float) * g.Hc * g.wpad;           // image rows
    b += sizeof(float2) * g.Hc * g.m2;            // row-pass result
    b += sizeof(float2) * (size_t)g.G * g.M;      // spectra of the G images
    return b;
}

static DftGeom make_geom(const pno_plan* p, int C) {
    DftGeom g;
    g.H = p->H; g.W = p->W; g.m1 = p->m1; g.m2 = p->m2; g.M = p->M;
    g.wpad = g.W + 1;
    g.Hc = g.H;
    while ((size_t)g.Hc * g.wpad * 4 > 36 * 1024 && g.Hc > 1) g.Hc = (g.Hc + 1) / 2;
    g.G = 8;
    while (g.G > 1 && (g.G > C || dft_smem_bytes(g) > 160 * 1024)) g.G >>= 1;
    return g;
}

struct Smem {
    float2* tabH; float2* tabW; int* kx; float* img; float2* row; float2* spec;
};

__device__ inline Smem carve(unsigned char* base, const DftGeom& g) {
    Smem s;
    s.tabH = reinterpret_cast<float2*>(base);
    s.tabW = s.tabH + g.H;
    s.spec = s.tabW + g.W;
    s.row = s.spec + (size_t)g.G * g.M;
    s.img = reinterpret_cast<float*>(s.row + (size_t)g.Hc * g.m2);
    s.kx = reinterpret_cast<int*>(s.img + (size_t)g.Hc * g.wpad);
    return s;
}

__device__ inline void load_tables(const pno_plan p, const Smem& s, const DftGeom& g) {
    for (int i = threadIdx.x; i < g.H; i += kThreads) s.tabH[i] = p.tabH[i];
    for (int i = threadIdx.x; i < g.W; i += kThreads) s.tabW[i] = p.tabW[i];
    for (int i = threadIdx.x; i < 2 * g.m1; i += kThreads) s.kx[i] = p.kx[i];
}

// x[B][C][H][W] -> xhat[mode][b][c] (re, im planes)
__global__ void __launch_bounds__(kThreads) k_dft_fwd(const pno_plan p, const DftGeom g, const float* __restrict__ x,
                                                      int B, int C, int grad_scale, float* __restrict__ xr,
                                                      float* __restrict__ xi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, g);
    load_tables(p, s, g);
    const int b = blockIdx.y, c0 = blockIdx.x * g.G;
    const int ng = min(g.G, C - c0);
    for (int i = threadIdx.x; i < g.G * g.M; i += kThreads) s.spec[i] = make_float2(0.f, 0.f);
    __syncthreads();

    for (int gi = 0; gi < ng; ++gi) {
        const float* img = x + ((size_t)b * C + c0 + gi) * g.H * g.W;
        float2* spec = s.spec + (size_t)gi * g.M;
        for (int h0 = 0; h0 < g.H; h0 += g.Hc) {
            const int nh = min(g.Hc, g.H - h0);
            for (int i = threadIdx.x; i < nh * g.W; i += kThreads) {
                const int r = i / g.W, w = i - r * g.W;
                s.img[r * g.wpad + w] = img[(size_t)(h0 + r) * g.W + w];
            }
            __syncthreads();
            // row pass: row[r][ky] = sum_w x[r][w] e^{-i 2 pi ky w / W}
            for (int i = threadIdx.x; i < nh * g.m2; i += kThreads) {
                const int r = i / g.m2, ky = i - r * g.m2;
                const float* xrow = s.img + r * g.wpad;
                float ar = 0.f, ai = 0.f;
                int t = 0;
                for (int w = 0; w < g.W; ++w) {
                    const float2 cs = s.tabW[t];
                    const float v = xrow[w];
                    ar = fmaf(v, cs.x, ar);
                    ai = fmaf(-v, cs.y, ai);
                    t += ky;
                    if (t >= g.W) t -= g.W;
                }
                s.row[r * g.m2 + ky] = make_float2(ar, ai);
            }
            __syncthreads();
            // column pass: spec[kx2][ky] += sum_r row[r][ky] e^{-i 2 pi kx (h0+r) / H}
            for (int m = threadIdx.x; m < g.M; m += kThreads) {
                const int kx2 = m / g.m2, ky = m - kx2 * g.m2;
                const int kx = s.kx[kx2];
                int t = (int)(((long long)kx * h0) % g.H);
                float ar = 0.f, ai = 0.f;
                for (int r = 0; r < nh; ++r) {
                    const float2 cs = s.tabH[t];
                    const float2 v = s.row[r * g.m2 + ky];
                    ar = fmaf(v.x, cs.x, ar);
                    ar = fmaf(v.y, cs.y, ar);
                    ai = fmaf(v.y, cs.x, ai);
                    ai = fmaf(-v.x, cs.y, ai);
                    t += kx;
                    if (t >= g.H) t -= g.H;
                }
                float2 acc = spec[m];
                acc.x += ar;
                acc.y += ai;
                spec[m] = acc;
            }
            __syncthreads();
        }
    }
    // write [mode][b][c0 .. c0+ng)
    for (int i = threadIdx.x; i < g.M * ng; i += kThreads) {
        const int m = i / ng, gi = i - m * ng;
        float2 v = s.spec[(size_t)gi * g.M + m];
        if (grad_scale) {
            const float cf = p.coef[m];
            v.x *= cf;
            v.y *= cf;
        }
        const size_t o = ((size_t)m * B + b) * C + c0 + gi;
        xr[o] = v.x;
        xi[o] = v.y;
    }
}

// yhat[mode][b][c] -> y[B][C][H][W] with fused (+ addend, activation) epilogue
__global__ void __launch_bounds__(kThreads) k_dft_inv(const pno_plan p, const DftGeom g, const float* __restrict__ yr,
                                                      const float* __restrict__ yi, int B, int C, int adjoint,
                                                      const float* __restrict__ addend, int act, float* __restrict__ y,
                                                      float* __restrict__ pre_act) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, g);
    load_tables(p, s, g);
    const int b = blockIdx.y, c0 = blockIdx.x * g.G;
    const int ng = min(g.G, C - c0);
    for (int i = threadIdx.x; i < g.M * ng; i += kThreads) {
        const int m = i / ng, gi = i - m * ng;
        const size_t o = ((size_t)m * B + b) * C + c0 + gi;
        const float cf = adjoint ? (p.coef[m] != 0.f ? 1.f : 0.f) : p.coef[m];
        s.spec[(size_t)gi * g.M + m] = make_float2(yr[o] * cf, yi[o] * cf);
    }
    __syncthreads();

    for (int gi = 0; gi < ng; ++gi) {
        const float2* spec = s.spec + (size_t)gi * g.M;
        const size_t img_off = ((size_t)b * C + c0 + gi) * g.H * g.W;
        for (int h0 = 0; h0 < g.H; h0 += g.Hc) {
            const int nh = min(g.Hc, g.H - h0);
            // column pass: row[r][ky] = sum_kx2 spec[kx2][ky] e^{+i 2 pi kx (h0+r) / H}
            for (int i = threadIdx.x; i < nh * g.m2; i += kThreads) {
                const int r = i / g.m2, ky = i - r * g.m2;
                const int h = h0 + r;
                float ar = 0.f, ai = 0.f;
                const int hstep = h % g.H;
#pragma unroll 1
                for (int blk = 0; blk < 2; ++blk) {     // kx is consecutive inside each block of rows
                    int t = (int)(((long long)s.kx[blk * g.m1] * h) % g.H);
                    for (int j = 0; j < g.m1; ++j) {
                        const float2 cs = s.tabH[t];
                        const float2 v = spec[(blk * g.m1 + j) * g.m2 + ky];
                        ar = fmaf(v.x, cs.x, ar);
                        ar = fmaf(-v.y, cs.y, ar);
                        ai = fmaf(v.x, cs.y, ai);
                        ai = fmaf(v.y, cs.x, ai);
                        t += hstep;
                        if (t >= g.H) t -= g.H;
                    }
                }
                s.row[r * g.m2 + ky] = make_float2(ar, ai);
            }
            __syncthreads();
            // row pass: y[r][w] = sum_ky Re(row[r][ky] e^{+i 2 pi ky w / W})
            for (int i = threadIdx.x; i < nh * g.W; i += kThreads) {
                const int r = i / g.W, w = i - r * g.W;
                const float2* zr = s.row + r * g.m2;
                float acc = 0.f;
                int t = 0;
                for (int ky = 0; ky < g.m2; ++ky) {
                    const float2 cs = s.tabW[t];
                    const float2 v = zr[ky];
                    acc = fmaf(v.x, cs.x, acc);
                    acc = fmaf(-v.y, cs.y, acc);
                    t += w;
                    if (t >= g.W) t -= g.W;
                }
                const size_t o = img_off + (size_t)(h0 + r) * g.W + w;
                if (addend) acc += addend[o];
                if (pre_act) pre_act[o] = acc;
                y[o] = act_apply(acc, act);
            }
            __syncthreads();
        }
    }
}

}  // namespace

int simt_dft_fwd(const pno_plan* p, const float* x, int B, int C, int grad_scale, float* xr, float* xi, cudaStream_t st) {
    const DftGeom g = make_geom(p, C);
    const size_t smem = dft_smem_bytes(g);
    static bool attr_set = false;   // per-process; the attribute is sticky per function
    if (!attr_set) {
        PNO_CUDA(cudaFuncSetAttribute(k_dft_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (smem > 200 * 1024) PNO_FAIL(PNO_E_UNSUPPORTED, "dft_fwd: %zu bytes of shared memory needed", smem);
    dim3 grid((C + g.G - 1) / g.G, B);
    k_dft_fwd<<<grid, kThreads, smem, st>>>(*p, g, x, B, C, grad_scale, xr, xi);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

int simt_dft_inv(const pno_plan* p, const float* yr, const float* yi, int B, int C, int adjoint, const float* addend,
                 int act, float* y, float* pre_act, cudaStream_t st) {
    const DftGeom g = make_geom(p, C);
    const size_t smem = dft_smem_bytes(g);
    static bool attr_set = false;
    if (!attr_set) {
        PNO_CUDA(cudaFuncSetAttribute(k_dft_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (smem > 200 * 1024) PNO_FAIL(PNO_E_UNSUPPORTED, "dft_inv: %zu bytes of shared memory needed", smem);
    dim3 grid((C + g.G - 1) / g.G, B);
    k_dft_inv<<<grid, kThreads, smem, st>>>(*p, g, yr, yi, B, C, adjoint, addend, act, y, pre_act);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
