// Two-pass pruned transforms on the tensor cores (engines PNO_ENGINE_TC_*), for grids whose twiddle tables do not fit in shared
// memory next to the operands of the fused single-kernel transforms (pno_tc_dft_fwd.cu / pno_tc_dft_inv.cu): 256-row images
// (BASELINE config 5: 256 x 256, 48 modes) and the fp32-parity engine at 128 x 128 with 32 modes (config 4).  Same mathematics
// (reference models.py:89, 93-94, 97), the axes as two GEMM launches with the small intermediate in HBM (it is [.., 2 m2] wide,
// i.e. m2 / (W/2) of the activation -- 19 .. 50 % -- and stays in the 126 MB L2 for the per-GPU batches of these configs):
//
//   forward   x [B,C,H,W] --k_tc_rowgemm (W axis)-->  T1 [(b,c,h)][re ky | im ky]            (columns padded to m2p = m2 up to 16)
//                         --k_tc_local<EPI=1> (H axis: D = C.T1 + (-i S).T1, complex combine)-->  T2 [(b,c)][kx2 (padded to 128)][re | im]
//                         --k_spec_out (transposition on CUDA cores, x coef for the adjoint)-->  xhat[mode][b][c]
//   inverse   yhat[mode][b][c] --k_spec_in (x alive c_ky / (HW))-->  Y2 [(b,c)][kx2 (padded to 32)][re | im]
//                         --k_tc_local<EPI=1> (H axis)-->  Z [(b,c)][h][re | im]
//                         --k_tc_rowgemm (W axis, epilogue: + addend, pre_act, activation)-->  y [B,C,H,W]
//
// Every contraction runs on tcgen05 (single-pass TF32 or 3xTF32); only the two layout transpositions of the (small) spectrum are
// CUDA-core kernels.  Shapes: H % 128 == 0 (H <= 256), W % 32 == 0 (W <= 256), 2*m1 <= 128, m2 <= 64, (B*C) any.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

namespace {

struct Tables2 {
    int m2p, R2p, K1p;
    float *twf_hi, *twf_lo;      // [2 m2p][W]   forward W-axis table (cos | -sin rows)
    float *twi_hi, *twi_lo;      // [W][2 m2p]   inverse W-axis table (its transpose)
    float *cf, *sf;              // [R2p][H]     forward H-axis tables (rows kx2 >= 2 m1 zero)
    float *ci, *si;              // [H][K1p]     inverse H-axis tables (columns of dead / padding kx2 zero)
};

inline float tf32_round(double v) {
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&f, &u, 4);
    return f;
}

int upload(float** dst, const std::vector<float>& v) {
    PNO_CUDA(cudaMalloc(dst, v.size() * sizeof(float)));
    PNO_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return PNO_OK;
}

int pad_to(int v, int m) { return (v + m - 1) / m * m; }

int tables2_get(const pno_plan* cp, Tables2** out) {
    pno_plan* p = const_cast<pno_plan*>(cp);
    if (p->tc_tables2) { *out = static_cast<Tables2*>(p->tc_tables2); return PNO_OK; }
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2;
    const double two_pi = 6.283185307179586476925286766559;
    Tables2* t = new Tables2();
    memset(t, 0, sizeof(*t));
    t->m2p = pad_to(m2, 16);
    t->R2p = pad_to(2 * m1, 128);
    t->K1p = pad_to(2 * m1, 32);
    const int N1 = 2 * t->m2p;
    std::vector<float> fh((size_t)N1 * W, 0.f), fl(fh.size(), 0.f), ih((size_t)W * N1, 0.f), il(ih.size(), 0.f);
    for (int ky = 0; ky < m2; ++ky)
        for (int w = 0; w < W; ++w) {
            const double ang = two_pi * (double)((long long)ky * w % W) / W;
            const double v[2] = {cos(ang), -sin(ang)};
            for (int part = 0; part < 2; ++part) {
                const int j = part * t->m2p + ky;
                const float hi = tf32_round(v[part]);
                const float lo = tf32_round(v[part] - (double)hi);
                fh[(size_t)j * W + w] = hi; fl[(size_t)j * W + w] = lo;
                ih[(size_t)w * N1 + j] = hi; il[(size_t)w * N1 + j] = lo;
            }
        }
    std::vector<float> cf((size_t)t->R2p * H, 0.f), sf(cf.size(), 0.f), ci((size_t)H * t->K1p, 0.f), si(ci.size(), 0.f);
    for (int r = 0; r < 2 * m1; ++r) {
        const int kx = r < m1 ? r : H - 2 * m1 + r;
        const bool alive = r >= m1 || kx < H - m1;                     // models.py:93-94: the second slice overwrites the first
        for (int h = 0; h < H; ++h) {
            const double ang = two_pi * (double)((long long)kx * h % H) / H;
            cf[(size_t)r * H + h] = (float)cos(ang);
            sf[(size_t)r * H + h] = (float)sin(ang);
            if (alive) {
                ci[(size_t)h * t->K1p + r] = (float)cos(ang);
                si[(size_t)h * t->K1p + r] = (float)sin(ang);
            }
        }
    }
    int rc;
    if ((rc = upload(&t->twf_hi, fh)) || (rc = upload(&t->twf_lo, fl)) || (rc = upload(&t->twi_hi, ih)) || (rc = upload(&t->twi_lo, il)) ||
        (rc = upload(&t->cf, cf)) || (rc = upload(&t->sf, sf)) || (rc = upload(&t->ci, ci)) || (rc = upload(&t->si, si)))
        return rc;
    p->tc_tables2 = t;
    *out = t;
    return PNO_OK;
}

constexpr int kImgs = 32;

// T2 [NI][R2p][2 m2p] -> xr / xi [(kx2 m2 + ky)][NI]  (x coef[mode] when scaled).   grid (2 m1, ceil(NI / 32)), 256 threads
__global__ void __launch_bounds__(256) k_spec_out(const float* __restrict__ t2, int NI, int R2p, int m2, int m2p,
                                                  const float* __restrict__ coef, int scaled, float* __restrict__ xr,
                                                  float* __restrict__ xi) {
    extern __shared__ float tile[];                   // [32 imgs][2 m2p + 1]
    const int kx2 = blockIdx.x, img0 = blockIdx.y * kImgs, ld = 2 * m2p + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < kImgs; i += 8) {
        const int img = img0 + i;
        if (img < NI) {
            const float* src = t2 + ((size_t)img * R2p + kx2) * (2 * m2p);
            for (int j = lane; j < 2 * m2p; j += 32) tile[i * ld + j] = src[j];
        }
    }
    __syncthreads();
    const int img = img0 + lane;
    if (img >= NI) return;
    for (int ky = warp; ky < m2; ky += 8) {
        const int mode = kx2 * m2 + ky;
        const float cf = scaled ? coef[mode] : 1.f;
        xr[(size_t)mode * NI + img] = tile[lane * ld + ky] * cf;
        xi[(size_t)mode * NI + img] = tile[lane * ld + m2p + ky] * cf;
    }
}

// yr / yi [(kx2 m2 + ky)][NI] -> Y2 [NI][K1p][2 m2p] (x coef; adjoint: x [coef != 0]); padding rows / columns zero.
// grid (K1p, ceil(NI / 32)), 256 threads
__global__ void __launch_bounds__(256) k_spec_in(const float* __restrict__ yr, const float* __restrict__ yi, int NI, int K1p, int m1,
                                                 int m2, int m2p, const float* __restrict__ coef, int adjoint, float* __restrict__ y2) {
    extern __shared__ float tile[];                   // [32 imgs][2 m2p + 1]
    const int kx2 = blockIdx.x, img0 = blockIdx.y * kImgs, ld = 2 * m2p + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kImgs * ld; i += 256) tile[i] = 0.f;
    __syncthreads();
    if (kx2 < 2 * m1) {
        const int img = img0 + lane;
        if (img < NI)
            for (int ky = warp; ky < m2; ky += 8) {
                const int mode = kx2 * m2 + ky;
                const float c = coef[mode];
                const float cf = adjoint ? (c != 0.f ? 1.f : 0.f) : c;
                tile[lane * ld + ky] = yr[(size_t)mode * NI + img] * cf;
                tile[lane * ld + m2p + ky] = yi[(size_t)mode * NI + img] * cf;
            }
    }
    __syncthreads();
    for (int i = warp; i < kImgs; i += 8) {
        const int img = img0 + i;
        if (img < NI) {
            float* dst = y2 + ((size_t)img * K1p + kx2) * (2 * m2p);
            for (int j = lane; j < 2 * m2p; j += 32) dst[j] = tile[i * ld + j];
        }
    }
}

int check_shape(const pno_plan* p, int B, int C, const char* who) {
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2;
    if (H % 128 || H > 256 || W % 32 || W > 256 || W < 32 || 2 * m1 > 128 || m2 > 64 || m2 < 17)
        PNO_FAIL(PNO_E_UNSUPPORTED, "%s (two-pass) tiles H in {128, 256}, W%%32 == 0, W <= 256, 2*m1 <= 128, 16 < m2 <= 64", who);
    if ((long long)B * C * H > 0x7fffffffLL * 64) PNO_FAIL(PNO_E_UNSUPPORTED, "%s: too many rows", who);
    return PNO_OK;
}

}  // namespace

void tc_plan_release2(pno_plan* p) {
    if (!p->tc_tables2) return;
    Tables2* t = static_cast<Tables2*>(p->tc_tables2);
    float* ptrs[] = {t->twf_hi, t->twf_lo, t->twi_hi, t->twi_lo, t->cf, t->sf, t->ci, t->si};
    for (float* q : ptrs)
        if (q) cudaFree(q);
    delete t;
    p->tc_tables2 = nullptr;
}

int tc_dft2_fwd(const pno_plan* p, int engine, const float* x, int B, int C, int grad_scale, float* xr, float* xi, cudaStream_t st) {
    int rc = check_shape(p, B, C, "tc_dft_fwd");
    if (rc) return rc;
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2, NI = B * C;
    const int m2p = pad_to(m2, 16), N1 = 2 * m2p, R2p = pad_to(2 * m1, 128);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(xi)) & 15)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd needs 16-byte aligned tensors");
    if (g_pno_dry_run) {       // the building blocks validate their own tilings
        float* const d = reinterpret_cast<float*>(uintptr_t(4096));
        if ((rc = tc_rowgemm(engine, d, (long long)NI * H, W, N1, d, d, 0, 0, nullptr, d, nullptr, st))) return rc;
        return tc_local_combine(engine, d, NI, H, R2p, N1, d, d, 1.f, d, st);
    }
    Tables2* t = nullptr;
    if ((rc = tables2_get(p, &t))) return rc;
    float* t1 = pno_scratch(1, (size_t)NI * H * N1 * 4, st);
    float* t2 = pno_scratch(2, (size_t)NI * R2p * N1 * 4, st);
    if (!t1 || !t2) PNO_FAIL(PNO_E_CUDA, "tc_dft_fwd: cannot allocate the two-pass intermediates");
    if ((rc = tc_rowgemm(engine, x, (long long)NI * H, W, N1, t->twf_hi, t->twf_lo, 0, 0, nullptr, t1, nullptr, st))) return rc;
    if ((rc = tc_local_combine(engine, t1, NI, H, R2p, N1, t->cf, t->sf, 1.f, t2, st))) return rc;
    dim3 grid(2 * m1, (NI + kImgs - 1) / kImgs);
    k_spec_out<<<grid, 256, kImgs * (N1 + 1) * 4, st>>>(t2, NI, R2p, m2, m2p, p->coef, grad_scale, xr, xi);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

int tc_dft2_inv(const pno_plan* p, int engine, const float* yr, const float* yi, int B, int C, int adjoint, const float* addend,
                int act, float* y, float* pre_act, cudaStream_t st) {
    int rc = check_shape(p, B, C, "tc_dft_inv");
    if (rc) return rc;
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2, NI = B * C;
    const int m2p = pad_to(m2, 16), N1 = 2 * m2p, K1p = pad_to(2 * m1, 32);
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(addend) | reinterpret_cast<uintptr_t>(pre_act) |
         reinterpret_cast<uintptr_t>(yr) | reinterpret_cast<uintptr_t>(yi)) & 31)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_inv (two-pass) needs 32-byte aligned tensors");
    if (g_pno_dry_run) {
        float* const d = reinterpret_cast<float*>(uintptr_t(4096));
        if ((rc = tc_local_combine(engine, d, NI, K1p, H, N1, d, d, -1.f, d, st))) return rc;
        return tc_rowgemm(engine, d, (long long)NI * H, N1, W, d, d, 1, act, addend, d, pre_act, st);
    }
    Tables2* t = nullptr;
    if ((rc = tables2_get(p, &t))) return rc;
    float* y2 = pno_scratch(1, (size_t)NI * K1p * N1 * 4, st);
    float* z = pno_scratch(2, (size_t)NI * H * N1 * 4, st);
    if (!y2 || !z) PNO_FAIL(PNO_E_CUDA, "tc_dft_inv: cannot allocate the two-pass intermediates");
    dim3 grid(K1p, (NI + kImgs - 1) / kImgs);
    k_spec_in<<<grid, 256, kImgs * (N1 + 1) * 4, st>>>(yr, yi, NI, K1p, m1, m2, m2p, p->coef, adjoint, y2);
    PNO_LAUNCH_CHECK();
    if ((rc = tc_local_combine(engine, y2, NI, K1p, H, N1, t->ci, t->si, -1.f, z, st))) return rc;
    return tc_rowgemm(engine, z, (long long)NI * H, N1, W, t->twi_hi, t->twi_lo, 1, act, addend, y, pre_act, st);
}

// Engine entry points of the transforms: the fused single-kernel form where its tables and operands fit in shared / tensor
// memory, the two-pass form above for the larger grids.
// PNO_DFT_TWOPASS=1 (diagnostics / tests) forces the two-pass form wherever it tiles the shape.
static bool force_two_pass() {
    const char* e = getenv("PNO_DFT_TWOPASS");
    return e && atoi(e) == 1;
}

int tc_dft_fwd(const pno_plan* p, int engine, const float* x, int B, int C, int grad_scale, float* xr, float* xi, cudaStream_t st) {
    if (force_two_pass()) {
        const int r2 = tc_dft2_fwd(p, engine, x, B, C, grad_scale, xr, xi, st);
        if (r2 != PNO_E_UNSUPPORTED) return r2;
    }
    const int rc = tc_dft_fwd_fused(p, engine, x, B, C, grad_scale, xr, xi, st);
    if (rc != PNO_E_UNSUPPORTED) return rc;
    return tc_dft2_fwd(p, engine, x, B, C, grad_scale, xr, xi, st);
}

int tc_dft_inv(const pno_plan* p, int engine, const float* yr, const float* yi, int B, int C, int adjoint, const float* addend,
               int act, float* y, float* pre_act, cudaStream_t st) {
    if (force_two_pass()) {
        const int r2 = tc_dft2_inv(p, engine, yr, yi, B, C, adjoint, addend, act, y, pre_act, st);
        if (r2 != PNO_E_UNSUPPORTED) return r2;
    }
    const int rc = tc_dft_inv_fused(p, engine, yr, yi, B, C, adjoint, addend, act, y, pre_act, st);
    if (rc != PNO_E_UNSUPPORTED) return rc;
    return tc_dft2_inv(p, engine, yr, yi, B, C, adjoint, addend, act, y, pre_act, st);
}
