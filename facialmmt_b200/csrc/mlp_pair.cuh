// CTA-pair (cta_group::2) variant of the streamed fused Swin MLP half-block (C = 192 / 384), see mlp_pair.cu. Same contract and
// argument struct as launch_mlp_stream (mlp_stream.cuh; `copies` and `trace` are ignored).
#pragma once
#include "mlp_stream.cuh"

namespace fmmt {

cudaError_t launch_mlp_pair(const MlpStreamArgs& a, cudaStream_t stream);
unsigned int* watchdog_addr_mlp_pair();

}  // namespace fmmt
