// Fused Swin MLP half-block for C = 96 (stage 1): x <- x + fc2(GELU(fc1(LayerNorm(x)))) in ONE kernel.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fmmt {

constexpr int MLP96_C = 96;
constexpr int MLP96_H = 384;
// bytes of the pre-swizzled shared-memory image of fc1.weight (384,96) and fc2.weight (96,384), see mlp_fused.cu
constexpr int MLP96_IMG_BYTES = 49152 + 24576 + 73728;

// Builds the image on the host from the reference's fp32 weights (row-major, nn.Linear layout).
void mlp96_pack_weights(const float* fc1_w, const float* fc2_w, __nv_bfloat16* img_host);

struct Mlp96Args {
  float* x = nullptr;          // fp32 [M, 96], updated in place (residual stream)
  int M = 0;
  const float* gamma = nullptr;  // norm2
  const float* beta = nullptr;
  float eps = 1e-5f;
  const __nv_bfloat16* img = nullptr;  // device copy of the packed image (MLP96_IMG_BYTES, 16-byte aligned)
  const float* b1 = nullptr;   // [384]
  const float* b2 = nullptr;   // [96]
};
cudaError_t launch_mlp96(const Mlp96Args& a, cudaStream_t stream);
inline double mlp96_flops(int M) { return 2.0 * 2.0 * M * (double)MLP96_C * MLP96_H; }

// same contract as read_mbar_timeout (gemm.cuh) for the barriers of this translation unit; tags 16..31
unsigned int read_mlp_timeout(bool reset);

// device address of this translation unit's pipeline-watchdog word (ptx.cuh)
unsigned int* watchdog_addr_mlp96();

}  // namespace fmmt
