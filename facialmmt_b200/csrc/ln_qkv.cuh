// LayerNorm + window gather fused into the qkv Linear of a Swin block (C = 192 / 384), see ln_qkv.cu.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fmmt {

struct LnQkvArgs {
  const float* x = nullptr;        // fp32 [M, C] residual stream in its current row order
  float* x_raw = nullptr;          // fp32 [M, C]: the gathered rows (this block's window order) = new residual stream; may be
                                   // nullptr when gather == nullptr
  int M = 0, C = 0;
  int T = 0;                       // tokens per frame (period of the gather map)
  const int* gather = nullptr;     // [T] window-order row r of a frame reads row gather[r]; nullptr = identity
  const float* gamma = nullptr;    // norm1
  const float* beta = nullptr;
  float eps = 1e-5f;
  const __nv_bfloat16* w = nullptr;   // qkv.weight bf16 [N, ldw] (nn.Linear layout, K contiguous)
  int ldw = 0;
  const float* bias = nullptr;     // [N]
  int N = 0;                       // output columns (3C), a multiple of 64
  __nv_bfloat16* out = nullptr;    // bf16 [M, ldo]
  int ldo = 0;
  long long* trace = nullptr;      // debug: [8][32] clock64 stamps of CTA 0
  int dbg = 0;                     // debug probes (see ln_qkv.cu); results are wrong when set
};
inline bool ln_qkv_supported(int C, int N) { return (C == 192 || C == 384) && N > 0 && (N % 64) == 0; }
cudaError_t launch_ln_qkv(const LnQkvArgs& a, cudaStream_t stream);
unsigned int* watchdog_addr_ln_qkv();

}  // namespace fmmt
