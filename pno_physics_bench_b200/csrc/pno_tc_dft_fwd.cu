// Pruned forward 2-D transform on the tensor cores (engines PNO_ENGINE_TC_*):
//   xhat[kx2*m2+ky][b][c] = sum_{h,w} x[b][c][h][w] e^{-2 pi i (kx h/H + ky w/W)}        (reference models.py:89, 93-94)
// as a chain of two dense twiddle-matrix GEMMs per image with the intermediate kept on chip:
//   stage 1 (rows):    R[(img,h)][c2]   = X[(img,h)][w] * TwT[c2][w]^T     c2 = ky (cos) | m2+ky (-sin);  M=128 rows
//   transposer:        TMEM R -> registers -> (tf32 hi/lo) -> K-major swizzled smem operands Rre[(img,ky)][h], Rim[..][h]
//   stage 2 (columns): Xre[kx2][(img,ky)] = C Rre^T + S Rim^T ;  Xim = C Rim^T - S Rre^T        (A = twiddles, negate-A bit)
// Stage-2 accumulators of a "super tile" (up to 8 consecutive channels) stay in TMEM so that the epilogue writes
// [mode][b][c] with 32-byte contiguous channel runs.
//   warp 0      TMA producer (x row tiles + the matching TwT k-block; C/S tables once)
//   warp 1      MMA issuer (both stages, software pipelined)
//   warps 2-5   tf32 hi/lo splitter of the x tiles (3xTF32 only)
//   warps 6-9   transposer + epilogue
#include <math.h>

#include <vector>

#include "pno_sm100.cuh"
#include "pno_tc.cuh"

using namespace sm100;

namespace {

constexpr int kThreads = 320;

struct FwdGeom {
    int H, W, m1, m2, M;
    int B, C;
    int IPT;        // images per 128-row tile (H <= 128)
    int N1p;        // stage-1 N: 2*m2 padded to 16
    int N2p;        // stage-2 N: IPT*m2 padded to 16
    int R2;         // rows of the C / S tables: 2*m1 padded to 8
    int KB1;        // W / 32
    int KB2;        // H / 32
    int TPS;        // tiles per super tile
    int n_super;    // super tiles in the launch
    int grad_scale;
    int stages;
    // shared-memory byte offsets (from the 1024-aligned base)
    int off_tab, off_ring, off_b2, stage_bytes, x_bytes, tw_bytes, tab_bytes, b2_bytes, smem_bytes;
};

struct TcTables {
    float *twT_hi, *twT_lo;   // [N1p][W]
    float *c_hi, *c_lo, *s_hi, *s_lo;   // [R2][H]
    // inverse-transform tables (pno_tc_dft_inv.cu)
    float *inv_a_hi, *inv_a_lo, *inv_w_hi, *inv_w_lo;
    int N1p, R2;
};

inline float tf32_round_host(double v) {
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&f, &u, 4);
    return f;
}

template <int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_dft_fwd(const FwdGeom g, const float* __restrict__ coef, float* __restrict__ xr, float* __restrict__ xi,
             const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmTwHi,
             const __grid_constant__ CUtensorMap tmTwLo, const __grid_constant__ CUtensorMap tmCHi,
             const __grid_constant__ CUtensorMap tmCLo, const __grid_constant__ CUtensorMap tmSHi,
             const __grid_constant__ CUtensorMap tmSLo) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kMaxStages = 4;
    __shared__ __align__(8) uint64_t bar_raw[kMaxStages], bar_split[kMaxStages], bar_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_tab, bar_d1_full[2], bar_d1_empty[2], bar_b2_full, bar_b2_empty, bar_d2_full,
        bar_d2_empty;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&bar_raw[s], 1);
            mbar_init(&bar_split[s], 128);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_tab, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_d1_full[s], 1);
            mbar_init(&bar_d1_empty[s], 128);
        }
        mbar_init(&bar_b2_full, 128);
        mbar_init(&bar_b2_empty, 1);
        mbar_init(&bar_d2_full, 1);
        mbar_init(&bar_d2_empty, 128);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    // TMEM columns: [0, 2*N1p) stage-1 accumulators (double buffered); then TPS slots of 2*N2p (re | im)
    const int d2_base = 2 * g.N1p;
    // tables: (C hi, C lo, S hi, S lo), each KB2 k-blocks of [R2 x 128 B]
    unsigned char* tab = smem + g.off_tab;
    const int tab_kb = g.R2 * 128;
    const int tab_one = g.KB2 * tab_kb;
    // B2 operands: (Rre hi, Rre lo, Rim hi, Rim lo), each KB2x (<=4) k-blocks of [N2p x 128 B]
    unsigned char* b2 = smem + g.off_b2;
    const int KT = g.H < 128 ? g.H : 128;            // h rows covered by one tile
    const int b2_kb = g.N2p * 128;
    const int b2_one = (KT / 32) * b2_kb;

    const int my_super0 = blockIdx.x;
    const int total_tiles_per_cta_step = g.TPS;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar_tab, 4 * tab_one);
            for (int kb = 0; kb < g.KB2; ++kb) {
                tma_load_2d(tab + 0 * tab_one + kb * tab_kb, &tmCHi, &bar_tab, kb * 32, 0);
                tma_load_2d(tab + 1 * tab_one + kb * tab_kb, &tmCLo, &bar_tab, kb * 32, 0);
                tma_load_2d(tab + 2 * tab_one + kb * tab_kb, &tmSHi, &bar_tab, kb * 32, 0);
                tma_load_2d(tab + 3 * tab_one + kb * tab_kb, &tmSLo, &bar_tab, kb * 32, 0);
            }
            int stage = 0, phase = 0;
            for (int sup = my_super0; sup < g.n_super; sup += gridDim.x) {
                for (int t = 0; t < total_tiles_per_cta_step; ++t) {
                    const int row0 = (sup * g.TPS + t) * 128;          // row of the [B*C*H][W] view
                    for (int kb = 0; kb < g.KB1; ++kb) {
                        mbar_wait(&bar_empty[stage], phase ^ 1);
                        unsigned char* st = smem + g.off_ring + (size_t)stage * g.stage_bytes;
                        mbar_arrive_expect_tx(&bar_raw[stage], g.x_bytes + 2 * g.tw_bytes);
                        tma_load_2d(st, &tmX, &bar_raw[stage], kb * 32, row0);
                        tma_load_2d(st + 2 * g.x_bytes, &tmTwHi, &bar_raw[stage], kb * 32, 0);
                        tma_load_2d(st + 2 * g.x_bytes + g.tw_bytes, &tmTwLo, &bar_raw[stage], kb * 32, 0);
                        if (++stage == g.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc1 = make_idesc(FMT_TF32, 128, g.N1p, MAJOR_K, MAJOR_K);
            const uint32_t idesc2 = make_idesc(FMT_TF32, 128, g.N2p, MAJOR_K, MAJOR_K);
            const uint32_t idesc2n = make_idesc(FMT_TF32, 128, g.N2p, MAJOR_K, MAJOR_K, 1, 0);
            const uint32_t khi = sdesc_base(0, 16, 1024, SWIZZLE_128B).hi;     // all operands: K-major SW128
            int stage = 0, phase = 0;
            int n1 = 0;      // stage-1 tiles issued
            int n2 = 0;      // stage-2 tiles issued
            int nsup = 0;    // super tiles started (stage 2)
            mbar_wait(&bar_tab, 0);

            auto issue_stage1 = [&]() {
                const int buf = n1 & 1;
                mbar_wait(&bar_d1_empty[buf], ((n1 >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d1 = tmem + buf * g.N1p;
                for (int kb = 0; kb < g.KB1; ++kb) {
                    mbar_wait(NSPLIT == 3 ? &bar_split[stage] : &bar_raw[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t x_hi = sdesc_base(smem_u32(smem + g.off_ring + (size_t)stage * g.stage_bytes), 16, 1024, SWIZZLE_128B).lo;
                    const uint32_t x_lo = x_hi + (g.x_bytes >> 4), tw_hi = x_hi + 2 * (g.x_bytes >> 4), tw_lo = tw_hi + (g.tw_bytes >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t accf = (kb | ks) ? 1u : 0u;
                        const uint32_t o = ks * 2;
                        mma_tf32(d1, x_hi + o, khi, tw_hi + o, khi, idesc1, accf);
                        if (NSPLIT == 3) {
                            mma_tf32(d1, x_lo + o, khi, tw_hi + o, khi, idesc1, 1u);
                            mma_tf32(d1, x_hi + o, khi, tw_lo + o, khi, idesc1, 1u);
                        }
                    }
                    mma_commit(&bar_empty[stage]);
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
                mma_commit(&bar_d1_full[buf]);
                ++n1;
            };

            auto issue_stage2 = [&]() {
                const int t = n2 % g.TPS;                       // tile within its super tile
                if (t == 0) {                                   // first write into the D2 slots of a new super tile
                    mbar_wait(&bar_d2_empty, (nsup & 1) ^ 1);
                    ++nsup;
                }
                mbar_wait(&bar_b2_full, n2 & 1);
                tc_fence_after_sync();
                // slot: one per image group (tile if H <= 128, image if H > 128); k-range of this tile inside C/S
                const int tiles_per_slot = g.H > 128 ? g.H / 128 : 1;
                const int slot = t / tiles_per_slot, part = t % tiles_per_slot;
                const uint32_t d_re = tmem + d2_base + slot * 2 * g.N2p, d_im = d_re + g.N2p;
                const uint32_t tb = sdesc_base(smem_u32(tab), 16, 1024, SWIZZLE_128B).lo;
                const uint32_t bb = sdesc_base(smem_u32(b2), 16, 1024, SWIZZLE_128B).lo;
                const uint32_t t1 = tab_one >> 4, bo1 = b2_one >> 4;       // (C hi, C lo, S hi, S lo) / (Rre hi, Rre lo, Rim hi, Rim lo)
                const int nkb = KT / 32;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int tkb = part * 4 + kb;              // k-block inside the full-H tables
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t first = (part | kb | ks) ? 1u : 0u;
                        const uint32_t c = tb + ((tkb * tab_kb) >> 4) + ks * 2, r = bb + ((kb * b2_kb) >> 4) + ks * 2;
                        // Xre += C Rre + S Rim ; Xim += C Rim - S Rre
                        mma_tf32(d_re, c, khi, r, khi, idesc2, first);
                        mma_tf32(d_re, c + 2 * t1, khi, r + 2 * bo1, khi, idesc2, 1u);
                        mma_tf32(d_im, c, khi, r + 2 * bo1, khi, idesc2, first);
                        mma_tf32(d_im, c + 2 * t1, khi, r, khi, idesc2n, 1u);
                        if (NSPLIT == 3) {
                            mma_tf32(d_re, c + t1, khi, r, khi, idesc2, 1u);
                            mma_tf32(d_re, c, khi, r + bo1, khi, idesc2, 1u);
                            mma_tf32(d_re, c + 3 * t1, khi, r + 2 * bo1, khi, idesc2, 1u);
                            mma_tf32(d_re, c + 2 * t1, khi, r + 3 * bo1, khi, idesc2, 1u);
                            mma_tf32(d_im, c + t1, khi, r + 2 * bo1, khi, idesc2, 1u);
                            mma_tf32(d_im, c, khi, r + 3 * bo1, khi, idesc2, 1u);
                            mma_tf32(d_im, c + 3 * t1, khi, r, khi, idesc2n, 1u);
                            mma_tf32(d_im, c + 2 * t1, khi, r + bo1, khi, idesc2n, 1u);
                        }
                    }
                }
                mma_commit(&bar_b2_empty);
                if (t == g.TPS - 1) mma_commit(&bar_d2_full);
                ++n2;
            };

            int my_tiles = 0;
            for (int sup = my_super0; sup < g.n_super; sup += gridDim.x) my_tiles += g.TPS;
            if (my_tiles > 0) issue_stage1();
            for (int t = 0; t < my_tiles; ++t) {
                if (t + 1 < my_tiles) issue_stage1();
                issue_stage2();
            }
        }
    } else if (warp < 6) {
        // ================================ hi/lo splitter of the x tile (3xTF32) ================================
        if (NSPLIT == 3) {
            const int t = threadIdx.x - 64;
            int stage = 0, phase = 0;
            for (int sup = my_super0; sup < g.n_super; sup += gridDim.x) {
                for (int tt = 0; tt < g.TPS * g.KB1; ++tt) {
                    mbar_wait(&bar_raw[stage], phase);
                    float4* hi = reinterpret_cast<float4*>(smem + g.off_ring + (size_t)stage * g.stage_bytes);
                    float4* lo = reinterpret_cast<float4*>(smem + g.off_ring + (size_t)stage * g.stage_bytes + g.x_bytes);
#pragma unroll 4
                    for (int i = t; i < g.x_bytes / 16; i += 128) {
                        const float4 v = hi[i];
                        float4 h;
                        h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u);
                        h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u);
                        h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u);
                        h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u);
                        hi[i] = h;
                        lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&bar_split[stage]);
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================================ transposer + epilogue ================================
        const int q = warp & 3;
        const int row = q * 32 + lane;                       // TMEM lane = row of the 128-row tile
        const int img_l = g.H < 128 ? row / g.H : 0;         // image within the tile
        const int hk = g.H < 128 ? row % g.H : row;          // k index inside this tile's h range
        int n = 0, nsup = 0;
        for (int sup = my_super0; sup < g.n_super; sup += gridDim.x, ++nsup) {
            for (int t = 0; t < g.TPS; ++t, ++n) {
                const int buf = n & 1;
                mbar_wait(&bar_d1_full[buf], (n >> 1) & 1);
                tc_fence_after_sync();
                mbar_wait(&bar_b2_empty, (n & 1) ^ 1);       // stage-2 MMAs of the previous tile have read B2
                const uint32_t t1 = tmem_addr(tmem, q * 32, buf * g.N1p);
                // zero-free transposition: element (n2 = img_l*m2 + ky, k = hk) of Rre / Rim
                for (int c0 = 0; c0 < 2 * g.m2; c0 += 4) {
                    uint32_t r0, r1, r2, r3;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                                 : "r"(t1 + c0));
                    tmem_ld_wait();
                    const float v[4] = {__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3)};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c2 = c0 + j;
                        const int im = c2 >= g.m2;
                        const int ky = c2 - im * g.m2;
                        const uint32_t off = kmajor_sw128_off(img_l * g.m2 + ky, hk, g.N2p);
                        unsigned char* dst = b2 + (im ? 2 : 0) * b2_one + off;
                        if (NSPLIT == 3) {
                            const float h = __uint_as_float((__float_as_uint(v[j]) + 0x1000u) & 0xFFFFE000u);
                            *reinterpret_cast<float*>(dst) = h;
                            *reinterpret_cast<float*>(dst + b2_one) = v[j] - h;
                        } else {
                            *reinterpret_cast<float*>(dst) = v[j];
                        }
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(&bar_d1_empty[buf]);
                fence_proxy_async_smem();
                mbar_arrive(&bar_b2_full);
            }
            // ---- epilogue of the super tile: lanes kx2 < 2*m1 hold Xre / Xim of GS images ----
            mbar_wait(&bar_d2_full, nsup & 1);
            tc_fence_after_sync();
            const int n_slots = g.H > 128 ? g.TPS / (g.H / 128) : g.TPS;
            const int GS = n_slots * g.IPT;                  // images (consecutive channels) in this super tile
            const int img0 = sup * GS;                       // first image = b*C + c0
            const int b = img0 / g.C, c0 = img0 % g.C;
            if (q * 32 < 2 * g.m1) {      // warp-uniform: tcgen05.ld is warp-collective, only the stores are predicated
                const bool active = row < 2 * g.m1;
                for (int part = 0; part < 2; ++part) {
                    float* dst_plane = part ? xi : xr;
                    for (int ky0 = 0; ky0 < g.m2; ky0 += 4) {
                        float vals[4][8];
#pragma unroll
                        for (int gi = 0; gi < 8; ++gi) {
                            if (gi < GS) {
                                const int slot = gi / g.IPT, il = gi % g.IPT;
                                const uint32_t ta = tmem_addr(tmem, q * 32,
                                                              d2_base + slot * 2 * g.N2p + part * g.N2p + il * g.m2 + ky0);
                                uint32_t r0, r1, r2, r3;
                                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                                             : "r"(ta));
                                tmem_ld_wait();
                                vals[0][gi] = __uint_as_float(r0); vals[1][gi] = __uint_as_float(r1);
                                vals[2][gi] = __uint_as_float(r2); vals[3][gi] = __uint_as_float(r3);
                            } else {
                                vals[0][gi] = vals[1][gi] = vals[2][gi] = vals[3][gi] = 0.f;
                            }
                        }
                        if (active) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int mode = row * g.m2 + ky0 + j;
                                const float cf = g.grad_scale ? coef[mode] : 1.f;
                                float* dst = dst_plane + ((size_t)mode * g.B + b) * g.C + c0;
                                if (GS == 8) {
                                    reinterpret_cast<float4*>(dst)[0] = make_float4(vals[j][0] * cf, vals[j][1] * cf, vals[j][2] * cf, vals[j][3] * cf);
                                    reinterpret_cast<float4*>(dst)[1] = make_float4(vals[j][4] * cf, vals[j][5] * cf, vals[j][6] * cf, vals[j][7] * cf);
                                } else {
#pragma unroll
                                    for (int gi = 0; gi < 8; ++gi)
                                        if (gi < GS) dst[gi] = vals[j][gi] * cf;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bar_d2_empty);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

int pad_to(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

// ---- plan-owned tensor-core tables ------------------------------------------------------------------------------------
static int upload(float** dst, const std::vector<float>& v) {
    PNO_CUDA(cudaMalloc(dst, v.size() * sizeof(float)));
    PNO_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return PNO_OK;
}

int tc_tables_get(const pno_plan* cp, void** out) {
    pno_plan* p = const_cast<pno_plan*>(cp);
    if (p->tc_tables) { *out = p->tc_tables; return PNO_OK; }
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2;
    TcTables* t = new TcTables();
    memset(t, 0, sizeof(*t));
    t->N1p = pad_to(2 * m2, 16);
    t->R2 = pad_to(2 * m1, 8);
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> tw_hi((size_t)t->N1p * W, 0.f), tw_lo(tw_hi.size(), 0.f);
    for (int c2 = 0; c2 < 2 * m2; ++c2)
        for (int w = 0; w < W; ++w) {
            const int ky = c2 % m2;
            const double ang = two_pi * (double)((long long)ky * w % W) / W;
            const double v = c2 < m2 ? cos(ang) : -sin(ang);
            const float hi = tf32_round_host(v);
            tw_hi[(size_t)c2 * W + w] = hi;
            tw_lo[(size_t)c2 * W + w] = tf32_round_host(v - (double)hi);
        }
    std::vector<float> c_hi((size_t)t->R2 * H, 0.f), c_lo(c_hi.size(), 0.f), s_hi(c_hi.size(), 0.f), s_lo(c_hi.size(), 0.f);
    for (int r = 0; r < 2 * m1; ++r) {
        const int kx = r < m1 ? r : H - 2 * m1 + r;
        for (int h = 0; h < H; ++h) {
            const double ang = two_pi * (double)((long long)kx * h % H) / H;
            const double c = cos(ang), s = sin(ang);
            const float ch = tf32_round_host(c), sh = tf32_round_host(s);
            c_hi[(size_t)r * H + h] = ch; c_lo[(size_t)r * H + h] = tf32_round_host(c - (double)ch);
            s_hi[(size_t)r * H + h] = sh; s_lo[(size_t)r * H + h] = tf32_round_host(s - (double)sh);
        }
    }
    int rc;
    if ((rc = upload(&t->twT_hi, tw_hi)) || (rc = upload(&t->twT_lo, tw_lo)) || (rc = upload(&t->c_hi, c_hi)) ||
        (rc = upload(&t->c_lo, c_lo)) || (rc = upload(&t->s_hi, s_hi)) || (rc = upload(&t->s_lo, s_lo)))
        return rc;
    p->tc_tables = t;
    *out = t;
    return PNO_OK;
}

void tc_plan_release(pno_plan* p) {
    if (!p->tc_tables) return;
    TcTables* t = static_cast<TcTables*>(p->tc_tables);
    float* ptrs[] = {t->twT_hi, t->twT_lo, t->c_hi, t->c_lo, t->s_hi, t->s_lo, t->inv_a_hi, t->inv_a_lo, t->inv_w_hi, t->inv_w_lo};
    for (float* q : ptrs)
        if (q) cudaFree(q);
    delete t;
    p->tc_tables = nullptr;
}

int tc_dft_fwd(const pno_plan* p, int engine, const float* x, int B, int C, int grad_scale, float* xr, float* xi,
               cudaStream_t st) {
    const int H = p->H, W = p->W, m1 = p->m1, m2 = p->m2;
    if (!(H == 32 || H == 64 || H == 128) || W % 32 || W > 256 || m2 % 4 || 2 * m1 > 128)
        PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd tiles H in {32,64,128}, W%%32 == 0, W <= 256, m2%%4 == 0, 2*m1 <= 128");
    FwdGeom g;
    memset(&g, 0, sizeof(g));
    g.H = H; g.W = W; g.m1 = m1; g.m2 = m2; g.M = p->M; g.B = B; g.C = C; g.grad_scale = grad_scale;
    g.IPT = 128 / H;
    g.N1p = pad_to(2 * m2, 16);
    g.N2p = pad_to(g.IPT * m2, 16);
    g.R2 = pad_to(2 * m1, 8);
    g.KB1 = W / 32;
    g.KB2 = H / 32;
    if (g.N1p > 256 || g.N2p > 256) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd: too many modes for one MMA");
    int slots = (512 - 2 * g.N1p) / (2 * g.N2p);
    int want = 8 / g.IPT;
    if (want < 1) want = 1;
    int ns = 1;
    while (ns * 2 <= slots && ns * 2 <= want) ns *= 2;
    const int GS = ns * g.IPT;
    if (C % GS) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd: channel count %d is not a multiple of %d", C, GS);
    g.TPS = ns;                                   // H <= 128: one tile per slot
    g.n_super = B * C / GS;
    const bool x3 = engine == PNO_ENGINE_TC_TF32X3;
    g.x_bytes = 128 * 128;                        // [128 rows][32 floats]
    g.tw_bytes = g.N1p * 128;
    g.stage_bytes = pad_to(2 * g.x_bytes + 2 * g.tw_bytes, 1024);
    g.tab_bytes = pad_to(4 * g.KB2 * g.R2 * 128 + 128 * 128, 1024);     // + slack: M=128 reads run past R2 rows
    const int KT = H < 128 ? H : 128;
    g.b2_bytes = pad_to(4 * (KT / 32) * g.N2p * 128, 1024);
    g.off_tab = 0;
    g.off_b2 = g.tab_bytes;
    g.off_ring = g.off_b2 + g.b2_bytes;
    g.stages = 4;
    while (g.stages > 2 && g.off_ring + g.stages * g.stage_bytes + 1024 > 220 * 1024) --g.stages;
    g.smem_bytes = g.off_ring + g.stages * g.stage_bytes + 1024;
    if (g.smem_bytes > 225 * 1024) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd: %d bytes of shared memory needed", g.smem_bytes);
    if ((reinterpret_cast<uintptr_t>(x) & 15)) PNO_FAIL(PNO_E_UNSUPPORTED, "tc_dft_fwd needs a 16-byte aligned input");

    void* tv = nullptr;
    int rc = tc_tables_get(p, &tv);
    if (rc) return rc;
    const TcTables* t = static_cast<const TcTables*>(tv);
    CUtensorMap tmX, tmTwHi, tmTwLo, tmCHi, tmCLo, tmSHi, tmSLo;
    {
        const uint64_t dims[2] = {(uint64_t)W, (uint64_t)B * C * H};
        const uint64_t str[1] = {(uint64_t)W * 4};
        const uint32_t box[2] = {32, 128};
        if ((rc = pno_encode_tmap(&tmX, x, 2, dims, str, box, 1))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)W, (uint64_t)g.N1p};
        const uint64_t str[1] = {(uint64_t)W * 4};
        const uint32_t box[2] = {32, (uint32_t)g.N1p};
        if ((rc = pno_encode_tmap(&tmTwHi, t->twT_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmTwLo, t->twT_lo, 2, dims, str, box, 1))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)H, (uint64_t)g.R2};
        const uint64_t str[1] = {(uint64_t)H * 4};
        const uint32_t box[2] = {32, (uint32_t)g.R2};
        if ((rc = pno_encode_tmap(&tmCHi, t->c_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmCLo, t->c_lo, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmSHi, t->s_hi, 2, dims, str, box, 1))) return rc;
        if ((rc = pno_encode_tmap(&tmSLo, t->s_lo, 2, dims, str, box, 1))) return rc;
    }
    int dev = 0, sms = 148;
    PNO_CUDA(cudaGetDevice(&dev));
    PNO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = g.n_super < sms ? g.n_super : sms;
    if (x3) {
        PNO_CUDA(cudaFuncSetAttribute(k_tc_dft_fwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        k_tc_dft_fwd<3><<<grid, kThreads, g.smem_bytes, st>>>(g, p->coef, xr, xi, tmX, tmTwHi, tmTwLo, tmCHi, tmCLo, tmSHi, tmSLo);
    } else {
        PNO_CUDA(cudaFuncSetAttribute(k_tc_dft_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        k_tc_dft_fwd<1><<<grid, kThreads, g.smem_bytes, st>>>(g, p->coef, xr, xi, tmX, tmTwHi, tmTwLo, tmCHi, tmCLo, tmSHi, tmSLo);
    }
    PNO_LAUNCH_CHECK();
    return PNO_OK;
}
