#include "common.cuh"
namespace kpreg {
size_t kpconv_gemm_tc_weight_bytes(int kd, int n) { return 256; }
int kpconv_gemm_tc_prepare_weights(const float*, int, int, float*, cudaStream_t) { return KPREG_E_INVALID; }
int launch_kpconv_gemm_tc(const float*, const float*, const float*, float*, int64_t, int, int, void*, cudaStream_t) { return KPREG_E_INVALID; }
}
