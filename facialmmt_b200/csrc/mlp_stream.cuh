// Fused Swin MLP half-block with streamed weights (C = 192 / 384): x <- x + fc2(GELU(fc1(LayerNorm(x)))) in ONE kernel.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fmmt {

struct MlpStreamArgs {
  float* x = nullptr;            // fp32 [M, C], updated in place (residual stream)
  int M = 0;
  int C = 0;                     // 192 or 384; hidden = 4C
  const float* gamma = nullptr;  // norm2
  const float* beta = nullptr;
  float eps = 1e-5f;
  const __nv_bfloat16* w1 = nullptr;  // fc1.weight bf16 [4C, ldw1] (nn.Linear layout, K contiguous)
  int ldw1 = 0;
  const float* b1 = nullptr;     // [4C]
  const __nv_bfloat16* w2 = nullptr;  // fc2.weight bf16 [C, ldw2]
  int ldw2 = 0;
  const float* b2 = nullptr;     // [C]
  long long* trace = nullptr;    // optional device buffer [4C/64][8]: clock64 stamps of CTA 0's first tile (debug)
  int copies = 1;                // w1 / w2 hold `copies` identical matrices stacked along the rows; CTA b streams copy
                                 // b % copies: every CTA walks the weights in the same order at the same time, and
                                 // spreading them over several copies spreads those reads over the L2 slices
};
inline bool mlp_stream_supported(int C, int H) { return (C == 192 || C == 384) && H == 4 * C; }
cudaError_t launch_mlp_stream(const MlpStreamArgs& a, cudaStream_t stream);
inline double mlp_stream_flops(int M, int C) { return 2.0 * 2.0 * M * (double)C * 4.0 * C; }

// same contract as read_mbar_timeout (gemm.cuh) for the barriers of this translation unit; tags 40..52
unsigned int read_mlp_stream_timeout(bool reset);

// device address of this translation unit's pipeline-watchdog word (ptx.cuh)
unsigned int* watchdog_addr_mlp_stream();

}  // namespace fmmt
