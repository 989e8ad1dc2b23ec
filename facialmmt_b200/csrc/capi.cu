// extern "C" boundary of libfacialmmt_b200.so (see include/facialmmt_b200.h).
#include "facialmmt_b200.h"

#include <atomic>
#include <cstdio>
#include <string>

#include "gemm.cuh"

namespace fmmt {
thread_local std::string g_last_error;
std::atomic<long long> g_launch_count{0};

int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return FMMT_OK;
  return set_error(FMMT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace fmmt

using namespace fmmt;

extern "C" {

FMMT_API const char* fmmt_last_error(void) { return g_last_error.c_str(); }
FMMT_API const char* fmmt_version(void) { return "facialmmt_b200 0.1 (sm_100a)"; }
FMMT_API int64_t fmmt_launch_count(void) { return g_launch_count.load(); }

FMMT_API int fmmt_op_gemm(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K,
                          const float* bias, int act, const float* residual, int ldr, float* out_f32, int ldo32,
                          void* out_bf16, int ldo16, const int* row_map, int map_period, int block_n, void* stream) {
  if (!A_bf16 || !W_bf16 || (!out_f32 && !out_bf16)) return set_error(FMMT_ERR_INVALID, "fmmt_op_gemm: null pointer");
  GemmArgs a;
  a.A = static_cast<const __nv_bfloat16*>(A_bf16); a.lda = lda;
  a.W = static_cast<const __nv_bfloat16*>(W_bf16); a.ldw = ldw;
  a.M = M; a.N = N; a.K = K;
  a.bias = bias; a.act = act;
  a.residual = residual; a.ldr = ldr;
  a.out_f32 = out_f32; a.ldo32 = ldo32;
  a.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16); a.ldo16 = ldo16;
  a.row_map = row_map; a.map_period = map_period;
  a.block_n = block_n;
  cudaError_t e = launch_gemm(a, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorInvalidValue) return set_error(FMMT_ERR_INVALID, "fmmt_op_gemm: invalid shape/alignment");
  g_launch_count.fetch_add(1);
  return check_cuda(e, "fmmt_op_gemm");
}

}  // extern "C"
