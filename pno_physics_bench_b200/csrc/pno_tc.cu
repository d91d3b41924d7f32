// Tensor-core engines (tcgen05 + TMEM).  Shapes a kernel does not tile return PNO_E_UNSUPPORTED with a message;
// the C ABI reports that to the caller (there is no silent fallback for an explicitly requested engine).
#include "pno_tc.cuh"


int tc_mix_fwd(const pno_plan*, int, const float*, const float*, int, int, int, const float*, const float*, const float*,
               const float*, PhiloxKey, float*, float*, cudaStream_t) {
    PNO_FAIL(PNO_E_UNSUPPORTED, "tensor-core mix engine is not built into this version");
}

