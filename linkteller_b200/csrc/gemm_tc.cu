// 3xTF32 tensor-core GEMM (tcgen05) — placeholder until the tcgen05 path lands: reports "not handled"
// so that lt_gemm_f32(algo=0) uses the fp32 SIMT kernel.
#include "common.cuh"

namespace lt {
int gemm_3xtf32(const float*, int64_t, const int32_t*, const float*, int64_t, int, const float*, int, float*, int64_t,
                int32_t, int32_t, int32_t, cudaStream_t, bool* handled) {
    *handled = false;
    return LT_OK;
}
}  // namespace lt
