// Local branch on CUDA cores: 1x1 conv mean (+ 1x1 conv log-variance + reparameterised activation-space noise).
//
// Replaces nn.Conv2d(C,C,1) x2 + ProbabilisticNeuralOperator.reparameterize (reference models.py:288-290, 298-303,
// 309-314, 264-275) with one kernel: both 1x1 convs share one read of x, the noise comes from the Philox stream in the
// epilogue, and the clamp / exp / clamp(min=1e-6) sequence of the reference is reproduced in order.
#include "pno_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int TP = 64;   // pixels per tile
constexpr int TO = 64;   // output channels per tile
constexpr int KC = 16;
constexpr int LD = 64 + 4;

// x [B][Ci][HW], Wm/Wl [Co][Ci], out [B][Co][HW].   grid (ceil(HW/TP), ceil(Co/TO), B)
template <bool kNoise>
__global__ void __launch_bounds__(kThreads) k_local_fwd(const float* __restrict__ x, int Ci, int Co, int HW,
                                                        const float* __restrict__ Wm, const float* __restrict__ bm,
                                                        const float* __restrict__ Wl, const float* __restrict__ bl,
                                                        PhiloxKey key, uint64_t batch_offset, float* __restrict__ out,
                                                        float* __restrict__ dout_dlv) {
    __shared__ __align__(16) float sx[KC][LD], swm[KC][LD], swl[kNoise ? KC : 1][LD];
    const int p_base = blockIdx.x * TP, o_base = blockIdx.y * TO, b = blockIdx.z;
    const int tid = threadIdx.x, to = tid >> 4, tp = tid & 15;
    const float* xb = x + (size_t)b * Ci * HW;
    float am[4][4] = {}, al[4][4] = {};
    for (int i0 = 0; i0 < Ci; i0 += KC) {
        {   // x tile [KC i][TP p]
            const int il = tid >> 4, p4 = (tid & 15) * 4;
            const int i = i0 + il;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = p_base + p4 + j;
                sx[il][p4 + j] = (i < Ci && p < HW) ? xb[(size_t)i * HW + p] : 0.f;
            }
        }
        {   // weight tiles [TO o][KC i] -> sw[i][o]
            const int ol = tid >> 2, i4 = (tid & 3) * 4;
            const int o = o_base + ol;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + i4 + j;
                const bool ok = o < Co && i < Ci;
                swm[i4 + j][ol] = ok ? Wm[(size_t)o * Ci + i] : 0.f;
                if (kNoise) swl[i4 + j][ol] = ok ? Wl[(size_t)o * Ci + i] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float4 xv = *reinterpret_cast<const float4*>(&sx[k][tp * 4]);
            const float4 wm = *reinterpret_cast<const float4*>(&swm[k][to * 4]);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, wa[4] = {wm.x, wm.y, wm.z, wm.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int p = 0; p < 4; ++p) am[q][p] = fmaf(wa[q], xa[p], am[q][p]);
            if (kNoise) {
                const float4 wl = *reinterpret_cast<const float4*>(&swl[k][to * 4]);
                const float wb[4] = {wl.x, wl.y, wl.z, wl.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int p = 0; p < 4; ++p) al[q][p] = fmaf(wb[q], xa[p], al[q][p]);
            }
        }
        __syncthreads();
    }
    const int p0 = p_base + tp * 4;
    if (p0 >= HW) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int o = o_base + to * 4 + q;
        if (o >= Co) continue;
        const size_t off = ((size_t)b * Co + o) * HW + p0;
        const float bias_m = bm ? bm[o] : 0.f;
        float v[4], d[4];
        if (kNoise) {
            const float bias_l = bl ? bl[o] : 0.f;
            const uint64_t e0 = ((batch_offset + b) * (uint64_t)Co + o) * (uint64_t)HW + p0;
            float z[4];
            if ((e0 & 3) == 0) {
                const float4 zz = philox_quad_normals(key, e0 >> 2);
                z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p) z[p] = philox_normal_at(key, e0 + p);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float lv = al[q][p] + bias_l;
                const float lvc = fminf(fmaxf(lv, -10.f), 10.f);
                const float s_raw = expf(0.5f * lvc);
                const float s = fmaxf(s_raw, 1e-6f);
                v[p] = fmaf(z[p], s, am[q][p] + bias_m);
                // d out / d lv: zero where either clamp is active (torch.clamp backward passes at the boundary value)
                const bool live = (lv >= -10.f) && (lv <= 10.f) && (s_raw >= 1e-6f);
                d[p] = live ? 0.5f * z[p] * s : 0.f;
            }
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) v[p] = am[q][p] + bias_m;
        }
        if (p0 + 3 < HW && (HW & 3) == 0) {
            *reinterpret_cast<float4*>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
            if (kNoise && dout_dlv) *reinterpret_cast<float4*>(dout_dlv + off) = make_float4(d[0], d[1], d[2], d[3]);
        } else {
            for (int p = 0; p < 4 && p0 + p < HW; ++p) {
                out[off + p] = v[p];
                if (kNoise && dout_dlv) dout_dlv[off + p] = d[p];
            }
        }
    }
}

}  // namespace

int simt_local_fwd(const float* x, int B, int Ci, int Co, int HW, const float* Wm, const float* bm, const float* Wl,
                   const float* bl, PhiloxKey key, uint64_t batch_offset, float* out, float* dout_dlv, cudaStream_t st) {
    dim3 grid((HW + TP - 1) / TP, (Co + TO - 1) / TO, B);
    if (Wl)
        k_local_fwd<true><<<grid, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
    else
        k_local_fwd<false><<<grid, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
