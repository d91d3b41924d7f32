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
                                                        PhiloxKey key_in, uint64_t batch_offset, float* __restrict__ out,
                                                        float* __restrict__ dout_dlv) {
    const PhiloxKey key = resolve_key(key_in);
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

// ---- reparameterisation epilogue shared by the thin kernels below ------------------------------------------------------
template <bool kDlv = true>
__device__ __forceinline__ void reparam4(const float m[4], const float l[4], const PhiloxKey& key, uint64_t e0, float v[4],
                                         float d[4]) {
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
        const float lvc = fminf(fmaxf(l[p], -10.f), 10.f);
        const float s_raw = fast_ex2(0.72134752044448170f * lvc);      // exp(lv/2), one MUFU
        const float s = fmaxf(s_raw, 1e-6f);
        v[p] = fmaf(z[p], s, m[p]);
        if (kDlv) d[p] = ((l[p] >= -10.f) && (l[p] <= 10.f) && (s_raw >= 1e-6f)) ? 0.5f * z[p] * s : 0.f;
    }
}

// Lift (models.py:288-290): Ci <= 8 input channels.  One thread = 4 consecutive pixels x a chunk of output channels; the
// kernel is a pure streaming write of out (and d out / d lv).   grid (ceil(HW/4/256), ceil(Co/kLiftChunk), B), HW % 4 == 0
constexpr int kLiftChunk = 32;
template <bool kNoise, int CIM>          // CIM: compile-time bound on Ci (4 or 8) so that the FMA loop has no dead iterations
__global__ void __launch_bounds__(kThreads) k_local_thin_in(const float* __restrict__ x, int Ci, int Co, int HW,
                                                            const float* __restrict__ Wm, const float* __restrict__ bm,
                                                            const float* __restrict__ Wl, const float* __restrict__ bl,
                                                            PhiloxKey key_in, uint64_t batch_offset, float* __restrict__ out,
                                                            float* __restrict__ dout_dlv) {
    const PhiloxKey key = resolve_key(key_in);
    __shared__ float swm[kLiftChunk][CIM], swl[kLiftChunk][CIM], sbm[kLiftChunk], sbl[kLiftChunk];
    const int o_base = blockIdx.y * kLiftChunk, b = blockIdx.z;
    for (int e = threadIdx.x; e < kLiftChunk * CIM; e += kThreads) {
        const int ol = e / CIM, i = e % CIM, o = o_base + ol;
        const bool ok = o < Co && i < Ci;
        swm[ol][i] = ok ? Wm[(size_t)o * Ci + i] : 0.f;
        swl[ol][i] = (ok && kNoise) ? Wl[(size_t)o * Ci + i] : 0.f;
        if (i == 0) {
            sbm[ol] = (o < Co && bm) ? bm[o] : 0.f;
            sbl[ol] = (o < Co && kNoise && bl) ? bl[o] : 0.f;
        }
    }
    __syncthreads();
    const int p0 = (blockIdx.x * kThreads + threadIdx.x) * 4;
    if (p0 >= HW) return;
    float xv[CIM][4];
#pragma unroll
    for (int i = 0; i < CIM; ++i) {
        if (i < Ci) {
            const float4 t = *reinterpret_cast<const float4*>(x + ((size_t)b * Ci + i) * HW + p0);
            xv[i][0] = t.x; xv[i][1] = t.y; xv[i][2] = t.z; xv[i][3] = t.w;
        } else {
            xv[i][0] = xv[i][1] = xv[i][2] = xv[i][3] = 0.f;
        }
    }
    for (int ol = 0; ol < kLiftChunk && o_base + ol < Co; ++ol) {
        const int o = o_base + ol;
        float m[4] = {sbm[ol], sbm[ol], sbm[ol], sbm[ol]}, l[4] = {sbl[ol], sbl[ol], sbl[ol], sbl[ol]};
#pragma unroll
        for (int i = 0; i < CIM; ++i) {
            const float wm = swm[ol][i], wl = swl[ol][i];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                m[p] = fmaf(wm, xv[i][p], m[p]);
                if (kNoise) l[p] = fmaf(wl, xv[i][p], l[p]);
            }
        }
        const size_t off = ((size_t)b * Co + o) * HW + p0;
        if (kNoise) {
            float v[4], d[4];
            const uint64_t e0 = ((batch_offset + b) * (uint64_t)Co + o) * (uint64_t)HW + p0;
            if (dout_dlv) {
                reparam4<true>(m, l, key, e0, v, d);
                *reinterpret_cast<float4*>(dout_dlv + off) = make_float4(d[0], d[1], d[2], d[3]);
            } else {
                reparam4<false>(m, l, key, e0, v, d);       // inference: no d out / d log-var bookkeeping
            }
            *reinterpret_cast<float4*>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            *reinterpret_cast<float4*>(out + off) = make_float4(m[0], m[1], m[2], m[3]);
        }
    }
}

// Projection (models.py:309-314): Co <= 4 output channels: a streaming read of x.  Block = 64 pixel quads x 4 channel
// groups (each thread reduces a quarter of the input channels with 8 independent 16-byte loads in flight), partial sums
// are combined through shared memory and the channel-group-0 threads run the reparameterisation epilogue.
// grid (ceil(HW/4/64), B), HW % 4 == 0
constexpr int kProjQuads = 64, kProjGroups = 4;
template <bool kNoise, int CO>          // CO = Co at compile time: no per-channel branches, so the 8 loads of an unrolled step batch
__global__ void __launch_bounds__(kThreads) k_local_thin_out(const float* __restrict__ x, int Ci, int /*Co*/, int HW,
                                                             const float* __restrict__ Wm, const float* __restrict__ bm,
                                                             const float* __restrict__ Wl, const float* __restrict__ bl,
                                                             PhiloxKey key_in, uint64_t batch_offset, float* __restrict__ out,
                                                             float* __restrict__ dout_dlv) {
    const PhiloxKey key = resolve_key(key_in);
    extern __shared__ float sw[];          // [2][Co][Ci] weights, then [groups-1][quads][2*4*4] partial sums
    constexpr int Co = CO;
    float* red = sw + 2 * Co * Ci;
    const int b = blockIdx.y;
    for (int e = threadIdx.x; e < Co * Ci; e += kThreads) {
        sw[e] = Wm[e];
        sw[Co * Ci + e] = kNoise ? Wl[e] : 0.f;
    }
    __syncthreads();
    const int quad = threadIdx.x % kProjQuads, grp = threadIdx.x / kProjQuads;
    const int p0 = (blockIdx.x * kProjQuads + quad) * 4;
    const bool live = p0 < HW;
    float m[CO][4] = {}, l[CO][4] = {};
    if (live) {
        const int per = (Ci + kProjGroups - 1) / kProjGroups;
        const int i0 = grp * per, i1 = min(Ci, i0 + per);
        const float* xb = x + (size_t)b * Ci * HW + p0;
#pragma unroll 8
        for (int i = i0; i < i1; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(xb + (size_t)i * HW));
            const float xv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int o = 0; o < CO; ++o) {
                const float wm = sw[o * Ci + i], wl = kNoise ? sw[(Co + o) * Ci + i] : 0.f;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    m[o][p] = fmaf(wm, xv[p], m[o][p]);
                    if (kNoise) l[o][p] = fmaf(wl, xv[p], l[o][p]);
                }
            }
        }
    }
    if (grp > 0) {
        float* r = red + ((size_t)(grp - 1) * kProjQuads + quad) * 32;
#pragma unroll
        for (int o = 0; o < CO; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                r[o * 4 + p] = m[o][p];
                r[16 + o * 4 + p] = l[o][p];
            }
    }
    __syncthreads();
    if (grp > 0 || !live) return;
    for (int gg = 0; gg < kProjGroups - 1; ++gg) {
        const float* r = red + ((size_t)gg * kProjQuads + quad) * 32;
#pragma unroll
        for (int o = 0; o < CO; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                m[o][p] += r[o * 4 + p];
                l[o][p] += r[16 + o * 4 + p];
            }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        const float bias_m = bm ? bm[o] : 0.f, bias_l = (kNoise && bl) ? bl[o] : 0.f;
        const size_t off = ((size_t)b * Co + o) * HW + p0;
        float mm[4], ll[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) { mm[p] = m[o][p] + bias_m; ll[p] = l[o][p] + bias_l; }
        if (kNoise) {
            float v[4], d[4];
            reparam4(mm, ll, key, ((batch_offset + b) * (uint64_t)Co + o) * (uint64_t)HW + p0, v, d);
            *reinterpret_cast<float4*>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
            if (dout_dlv) *reinterpret_cast<float4*>(dout_dlv + off) = make_float4(d[0], d[1], d[2], d[3]);
        } else {
            *reinterpret_cast<float4*>(out + off) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        }
    }
}


// ---- general-shape backward of the local branch on CUDA cores (fp32 engine, and any shape the tensor-core tilings do not cover) ----
// G[b][r][p] = g[b][r][p] for r < Co, g[b][r-Co][p] * dlv[b][r-Co][p] for r >= Co (the stacked upstream gradient dYm | dYl).
__device__ __forceinline__ float stacked_grad(const float* __restrict__ g, const float* __restrict__ dlv, int Co, int HW, int b,
                                              int r, int p) {
    if (r < Co) return g[((size_t)b * Co + r) * HW + p];
    const size_t o = ((size_t)b * Co + (r - Co)) * HW + p;
    return g[o] * dlv[o];
}

constexpr int BT = 64, BK = 16, BLD = BT + 4;

// dx[b][i][p] = sum_r Wcat[r][i] G[b][r][p]   (Wcat = [Wm; Wl], rows x Ci).   grid (ceil(HW/64), ceil(Ci/64), B)
__global__ void __launch_bounds__(kThreads) k_local_bwd_dx(const float* __restrict__ g, const float* __restrict__ dlv,
                                                           const float* __restrict__ Wm, const float* __restrict__ Wl, int Ci,
                                                           int Co, int HW, int rows, float* __restrict__ dx) {
    __shared__ float sG[BK][BLD], sW[BK][BLD];
    const int p_base = blockIdx.x * BT, i_base = blockIdx.y * BT, b = blockIdx.z;
    const int tid = threadIdx.x, ti = tid >> 4, tp = tid & 15;
    float acc[4][4] = {};
    for (int r0 = 0; r0 < rows; r0 += BK) {
        for (int e = tid; e < BK * BT; e += kThreads) {
            const int k = e / BT, c = e - k * BT;
            const int r = r0 + k, p = p_base + c, i = i_base + c;
            sG[k][c] = (r < rows && p < HW) ? stacked_grad(g, dlv, Co, HW, b, r, p) : 0.f;
            sW[k][c] = (r < rows && i < Ci) ? (r < Co ? Wm[(size_t)r * Ci + i] : Wl[(size_t)(r - Co) * Ci + i]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float wv[4], gv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { wv[a] = sW[k][ti * 4 + a]; gv[a] = sG[k][tp * 4 + a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(wv[a], gv[c], acc[a][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i_base + ti * 4 + a;
        if (i >= Ci) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int p = p_base + tp * 4 + c;
            if (p < HW) dx[((size_t)b * Ci + i) * HW + p] = acc[a][c];
        }
    }
}

// dw[r][i] += sum over this block's slice of (b, p) of G[b][r][p] x[b][i][p];  db[r] += sum G[b][r][p]  (blocks with blockIdx.y == 0).
// grid (ceil(rows/64), ceil(Ci/64), n_split); the slice is a range of 16-pixel chunks of the flattened (b, chunk) index.
__global__ void __launch_bounds__(kThreads) k_local_bwd_dw(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ dlv, int B, int Ci, int Co, int HW, int rows,
                                                           int chunks_per_b, int chunks_per_split, float* __restrict__ dw,
                                                           float* __restrict__ db) {
    __shared__ float sG[BT][BK + 1], sX[BT][BK + 1];
    const int r_base = blockIdx.x * BT, i_base = blockIdx.y * BT;
    const int tid = threadIdx.x, tr = tid >> 4, ti = tid & 15;
    const int total = B * chunks_per_b;
    const int c_lo = blockIdx.z * chunks_per_split, c_hi = min(total, c_lo + chunks_per_split);
    float acc[4][4] = {};
    float bsum[4] = {};
    for (int ch = c_lo; ch < c_hi; ++ch) {
        const int b = ch / chunks_per_b, p0 = (ch - b * chunks_per_b) * BK;
        for (int e = tid; e < BT * BK; e += kThreads) {
            const int row = e / BK, k = e - row * BK;
            const int p = p0 + k, r = r_base + row, i = i_base + row;
            sG[row][k] = (r < rows && p < HW) ? stacked_grad(g, dlv, Co, HW, b, r, p) : 0.f;
            sX[row][k] = (i < Ci && p < HW) ? x[((size_t)b * Ci + i) * HW + p] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float gv[4], xv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { gv[a] = sG[tr * 4 + a][k]; xv[a] = sX[ti * 4 + a][k]; }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                bsum[a] += gv[a];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(gv[a], xv[c], acc[a][c]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int r = r_base + tr * 4 + a;
        if (r >= rows) continue;
        if (dw) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = i_base + ti * 4 + c;
                if (i < Ci) atomicAdd(dw + (size_t)r * Ci + i, acc[a][c]);
            }
        }
        if (db && blockIdx.y == 0 && ti == 0) atomicAdd(db + r, bsum[a]);
    }
}

}  // namespace

int simt_local_fwd(const float* x, int B, int Ci, int Co, int HW, const float* Wm, const float* bm, const float* Wl,
                   const float* bl, PhiloxKey key, uint64_t batch_offset, float* out, float* dout_dlv, cudaStream_t st) {
    const bool aligned = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                             reinterpret_cast<uintptr_t>(dout_dlv)) & 15) == 0;
    if (aligned && Ci <= 8) {            // lift-shaped: streaming write
        dim3 g((HW / 4 + kThreads - 1) / kThreads, (Co + kLiftChunk - 1) / kLiftChunk, B);
        if (Ci <= 4) {
            if (Wl) k_local_thin_in<true, 4><<<g, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
            else k_local_thin_in<false, 4><<<g, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
        } else {
            if (Wl) k_local_thin_in<true, 8><<<g, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
            else k_local_thin_in<false, 8><<<g, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
        }
        PNO_LAUNCH_CHECK();
        return PNO_OK;
    }
    const size_t proj_smem = ((size_t)2 * Co * Ci + (size_t)(kProjGroups - 1) * kProjQuads * 32) * 4;
    if (aligned && Co <= 4 && proj_smem <= 48 * 1024) {   // projection-shaped: streaming read
        dim3 g((HW / 4 + kProjQuads - 1) / kProjQuads, B);
        const size_t smem = proj_smem;
#define PNO_PROJ(NZ, CO_) k_local_thin_out<NZ, CO_><<<g, kThreads, smem, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv)
        if (Wl) { if (Co == 1) PNO_PROJ(true, 1); else if (Co == 2) PNO_PROJ(true, 2); else if (Co == 3) PNO_PROJ(true, 3); else PNO_PROJ(true, 4); }
        else { if (Co == 1) PNO_PROJ(false, 1); else if (Co == 2) PNO_PROJ(false, 2); else if (Co == 3) PNO_PROJ(false, 3); else PNO_PROJ(false, 4); }
#undef PNO_PROJ
        PNO_LAUNCH_CHECK();
        return PNO_OK;
    }
    dim3 grid((HW + TP - 1) / TP, (Co + TO - 1) / TO, B);
    if (Wl)
        k_local_fwd<true><<<grid, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
    else
        k_local_fwd<false><<<grid, kThreads, 0, st>>>(x, Ci, Co, HW, Wm, bm, Wl, bl, key, batch_offset, out, dout_dlv);
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}

// Backward of simt_local_fwd for any shape (see pno_local_bwd in include/pno_b200.h): dx, dw = [dWm; dWl], db = [dbm; dbl].
int simt_local_bwd(const float* x, const float* g, const float* dout_dlv, int B, int Ci, int Co, int HW, const float* Wm,
                   const float* Wl, float* dx, float* dw, float* db, cudaStream_t st) {
    const int noisy = Wl != nullptr;
    const int rows = noisy ? 2 * Co : Co;
    if (noisy && !dout_dlv) PNO_FAIL(PNO_E_INVALID, "pno_local_bwd: d out / d log_var is needed when Wl is given");
    if (B > 65535) PNO_FAIL(PNO_E_INVALID, "pno_local_bwd: batch %d exceeds the grid limit", B);
    if (dx) {
        dim3 grid((HW + BT - 1) / BT, (Ci + BT - 1) / BT, B);
        k_local_bwd_dx<<<grid, kThreads, 0, st>>>(g, dout_dlv, Wm, Wl, Ci, Co, HW, rows, dx);
        PNO_LAUNCH_CHECK();
    }
    if (dw || db) {
        if (dw) PNO_CUDA(cudaMemsetAsync(dw, 0, (size_t)rows * Ci * 4, st));
        if (db) PNO_CUDA(cudaMemsetAsync(db, 0, (size_t)rows * 4, st));
        const int chunks_per_b = (HW + BK - 1) / BK;
        const int total = B * chunks_per_b;
        const int tiles = ((rows + BT - 1) / BT) * ((Ci + BT - 1) / BT);
        int n_split = (4 * 148 + tiles - 1) / tiles;
        if (n_split > total) n_split = total;
        if (n_split > 65535) n_split = 65535;
        const int per = (total + n_split - 1) / n_split;
        n_split = (total + per - 1) / per;
        dim3 grid((rows + BT - 1) / BT, (Ci + BT - 1) / BT, n_split);
        k_local_bwd_dw<<<grid, kThreads, 0, st>>>(x, g, dout_dlv, B, Ci, Co, HW, rows, chunks_per_b, per, dw, db);
        PNO_LAUNCH_CHECK();
    }
    return PNO_OK;
}
