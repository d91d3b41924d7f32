// Tensor-core engines (tcgen05 + TMEM).  Shapes a kernel does not tile return PNO_E_UNSUPPORTED with a message;
// the C ABI reports that to the caller (there is no silent fallback for an explicitly requested engine).
#include "pno_tc.cuh"


