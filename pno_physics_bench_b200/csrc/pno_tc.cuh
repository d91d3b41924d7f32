// Tensor-core (tcgen05 / TMEM / TMA) engines: entry points called from pno_capi.cu.
#pragma once
#include "pno_common.cuh"

void tc_plan_release(pno_plan* p);
void tc_plan_release_inv(pno_plan* p);
void tc_plan_release2(pno_plan* p);
// building blocks of the two-pass transforms (pno_tc_gemm.cu, pno_tc_local.cu) and the transforms themselves (pno_tc_dft2.cu)
int tc_rowgemm(int engine, const float* A, long long rows, int K, int N, const float* t_hi, const float* t_lo, int epi, int act,
               const float* addend, float* out, float* pre_act, cudaStream_t st);
int tc_local_combine(int engine, const float* x, int B, int Ci, int Co, int TN, const float* Wc, const float* Ws, float sign, float* out,
                     cudaStream_t st);
int tc_dft_fwd_fused(const pno_plan* p, int engine, const float* x, int B, int C, int grad_scale, float* xr, float* xi,
                     cudaStream_t st);
int tc_dft_inv_fused(const pno_plan* p, int engine, const float* yr, const float* yi, int B, int C, int adjoint,
                     const float* addend, int act, float* y, float* pre_act, cudaStream_t st);
int tc_dft2_fwd(const pno_plan* p, int engine, const float* x, int B, int C, int grad_scale, float* xr, float* xi, cudaStream_t st);
int tc_dft2_inv(const pno_plan* p, int engine, const float* yr, const float* yi, int B, int C, int adjoint, const float* addend,
                int act, float* y, float* pre_act, cudaStream_t st);
int tc_tables_get(const pno_plan* p, void** tables);
int tc_dft_fwd(const pno_plan* p, int engine, const float* x, int B, int C, int grad_scale, float* xr, float* xi,
               cudaStream_t st);
int tc_dft_inv(const pno_plan* p, int engine, const float* yr, const float* yi, int B, int C, int adjoint,
               const float* addend, int act, float* y, float* pre_act, cudaStream_t st);
// mode_count < 0: all modes of the plan.  Otherwise the launch covers global modes [mode_first, mode_first + mode_count) (inside one
// weight block), the spectra and the weight / gradient arrays hold exactly those modes (mu1 == mu2, lv1 == lv2 = the shard).
int tc_mix_fwd(const pno_plan* p, int engine, const float* xr, const float* xi, int B, int Ci, int Co, const float* mu1,
               const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* yr, float* yi, cudaStream_t st,
               int mode_first = 0, int mode_count = -1);
int tc_mix_bwd_dx(const pno_plan* p, int engine, const float* gr, const float* gi, int B, int Ci, int Co, const float* mu1,
                  const float* mu2, const float* lv1, const float* lv2, PhiloxKey key, float* dxr, float* dxi, cudaStream_t st,
                  int mode_first = 0, int mode_count = -1);
int tc_local_fwd(int engine, const float* x, int B, int Ci, int Co, int HW, const float* Wm, const float* bm,
                 const float* Wl, const float* bl, PhiloxKey key, uint64_t batch_offset, float* out, float* dout_dlv,
                 cudaStream_t st);
// backward of the local branch (pno_tc_local_bwd.cu): dx, stacked weight gradients dw = [dWm; dWl] ([rows][Ci]) and bias sums db
size_t tc_local_bwd_workspace_bytes(int B, int Ci, int Co, int HW, int noisy);
int tc_local_bwd(int engine, const float* x, const float* g, const float* dout_dlv, int B, int Ci, int Co, int HW, const float* Wm,
                 const float* Wl, float* dx, float* dw, float* db, void* workspace, cudaStream_t st);
// weight gradient of the channel mix (pno_tc_mix_dw.cu)
int tc_mix_bwd_dw(const pno_plan* p, int engine, const float* xr, const float* xi, const float* gr, const float* gi, int B, int Ci,
                  int Co, const float* lv1, const float* lv2, PhiloxKey key, float* dmu1, float* dmu2, float* dlv1, float* dlv2,
                  cudaStream_t st, int mode_first = 0, int mode_count = -1);
