// Fused Swin attention half-block for C = 96 / 3 heads (Swin-tiny stage 1):
//   x_out[r] = x[g(r)] + proj( W-MSA / SW-MSA( LayerNorm1( x[g(r)] ) ) )         Swin_Transformer.py:238-264, 113-143
// in ONE persistent tcgen05 kernel; g = the composed roll + window_partition row gather of this block.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fmmt {

constexpr int ATTN96_C = 96;
constexpr int ATTN96_HEADS = 3;
constexpr int ATTN96_N = 49;        // tokens per 7x7 window
// bytes of the pre-swizzled shared-memory image of qkv.weight (288,96) and proj.weight (96,96), see attn_fused.cu
constexpr int ATTN96_IMG_BYTES = 3 * 288 * 64 + 3 * 96 * 64;   // 73728
constexpr int ATTN96_TAB_FLOATS = 3 * 169;                      // relative_position_bias_table transposed to [head][169], * log2(e)

// Host: builds the weight image from the reference's fp32 qkv.weight (288,96) / proj.weight (96,96) (nn.Linear layout) and
// the bias lookup table from relative_position_bias_table (169,3).
void attn96_pack(const float* qkv_w, const float* proj_w, const float* rel_table, __nv_bfloat16* img_host, float* tab_host);

struct Attn96Args {
  const float* x = nullptr;         // fp32 [M, 96] residual stream in its current row order
  float* x_out = nullptr;           // fp32 [M, 96] in THIS block's window order (must not alias x unless gather == nullptr)
  int M = 0;                        // rows = frames * T, T = tokens per frame (multiple of 98)
  int T = 0;
  const int* gather = nullptr;      // [T]: window-order row r reads row gather[r] of x (per frame); nullptr = identity
  const float* gamma = nullptr;     // norm1
  const float* beta = nullptr;
  float eps = 1e-5f;
  const __nv_bfloat16* img = nullptr;   // ATTN96_IMG_BYTES
  const float* tab = nullptr;       // ATTN96_TAB_FLOATS
  const float* qkv_b = nullptr;     // [288]
  const float* proj_b = nullptr;    // [96]
  const int8_t* rid = nullptr;      // [nW, 49] shift-region ids (SW-MSA) or nullptr (W-MSA)
  const int8_t* wflag = nullptr;    // [nW] 1 where a window spans more than one shift region (needed with rid)
  int nW = 0;                       // windows per frame
  float scale = 0.17677669529663687f;   // head_dim^-0.5 (Swin_Transformer.py:86)
  long long* trace = nullptr;       // optional device buffer [8][32]: clock64 stamps of CTA 0's first 8 tiles (debug)
};
cudaError_t launch_attn96(const Attn96Args& a, cudaStream_t stream);
inline double attn96_flops(int M) {   // qkv + proj Linear layers + QK^T + PV (2*MAC, algorithmic: 49 keys per query)
  return 2.0 * M * 96.0 * (288.0 + 96.0) + 4.0 * M * 49.0 * 96.0;
}
unsigned int* watchdog_addr_attn96();

}  // namespace fmmt
