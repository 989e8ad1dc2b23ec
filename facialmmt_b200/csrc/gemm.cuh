// Host-side interface of the tcgen05 GEMM used by every Linear layer on the path.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fmmt {

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_TANH = 3 };

// out[dest(r), n] = act( sum_k A[r,k] * W[n,k] + bias[n] ) + residual[dest(r), n]
//   A  : bf16 row-major [M, lda]  (K contiguous)      -> "x" of nn.Linear
//   W  : bf16 row-major [N, ldw]  (K contiguous)      -> nn.Linear.weight as stored by the reference
// dest(r): optional row remap so that producers can run in window order and scatter back
//   (Swin window_reverse + roll, Swin_Transformer.py:258-264) or write into a slice of a concat buffer.
struct GemmArgs {
  const __nv_bfloat16* A = nullptr;
  int lda = 0;
  const __nv_bfloat16* W = nullptr;
  int ldw = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;      // [N] fp32
  int act = ACT_NONE;
  const float* residual = nullptr;  // fp32 [*, ldr], indexed by dest row
  int ldr = 0;
  int res_mod = 0;                  // >0: residual row = dest % res_mod (broadcast table, e.g. learned positions)
  float* out_f32 = nullptr;         // fp32 [*, ldo32], indexed by dest row
  int ldo32 = 0;
  __nv_bfloat16* out_bf16 = nullptr;  // bf16 [*, ldo16], indexed by dest row
  int ldo16 = 0;
  const int* row_map = nullptr;     // dest = (r / map_period) * map_period + row_map[r % map_period]
  int map_period = 0;
  int rows_in = 0, rows_out = 0, row_off = 0;  // if rows_in>0: dest = (r/rows_in)*rows_out + row_off + r%rows_in
  const float* ln_gamma = nullptr;  // LayerNorm over the N output columns fused into the epilogue (fp32 output, N <= 256,
  const float* ln_beta = nullptr;   //   N % 32 == 0, no activation / residual): out = LN(A W^T + bias) * gamma + beta
  float ln_eps = 1e-5f;
  int block_n = 0;                  // 0 = choose automatically
  int force_generic = 0;            // 1 = register-path epilogue even where the TMA-epilogue fast path applies
  int two_cta = 0;                  // CTA-pair kernel (cta_group::2, 256-row tiles): 0 = heuristic, 1 = force, -1 = never
};

// Returns cudaSuccess or the launch / tensor-map error. Asynchronous on `stream`.
cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t stream);

// 2-D SWIZZLE_128B tensor map over a row-major matrix (dim0 = columns, dim1 = rows; box_cols * esize must be 128).
// OOB reads give zeros, OOB writes are clipped. Shared with the fused kernels (mlp_fused.cu).
bool make_tmap_2d(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int esize, long long rows, long long cols,
                  long long ld_elems, int box_cols, int box_rows);

// 3-D view of a K-major bf16 matrix with K % 64 == 0: (column within a 64-wide k-block, row, k-block index); one box =
// `box_kb` consecutive k-blocks of `box_rows` rows, landing as consecutive SWIZZLE_128B [box_rows x 128 B] tiles.
bool make_tmap_kblocks_2d(CUtensorMap* tm, const void* base, long long rows, long long K, long long ld_elems, int box_rows,
                          int box_kb);

// Non-zero if a pipeline wait inside a GEMM kernel timed out since the last reset (a protocol bug): bit 31 set,
// bits 24-30 = which barrier, 12-23 = CTA, 0-11 = thread. Synchronises the device.
unsigned int read_mbar_timeout(bool reset);

// Tensor-pipe probe (bench only): cycles per tcgen05.mma (M = 128, N = n, K = 16, bf16) with operands resident in
// shared memory and no other traffic; all SMs run it at once. Synchronous.
double mma_rate_probe(int n, int iters, int mode = 0);

// Feed probe (bench only): TMA bytes per cycle per SM from an L2-resident matrix (out2[0]) with `nstage` boxes of
// 64 x box_rows bf16 in flight on `grid` CTAs; mode 1 runs N = 256 MMAs beside it and reports cycles per MMA (out2[1]).
int feed_probe(int iters, int nstage, int box_rows, int mode, int grid, double* out2);

// Feed probe 2: bytes per cycle per SM with `nthr` independent TMA streams and a row pitch of `pitch_elems` bf16.
double feed_probe2(int iters, int nstage, int box_rows, int pitch_elems, int nthr, int grid);

// Algorithmic work of one launch (2*M*N*K) for roofline accounting.
inline double gemm_flops(const GemmArgs& a) { return 2.0 * a.M * (double)a.N * a.K; }

// device address of this translation unit's pipeline-watchdog word (ptx.cuh)
unsigned int* watchdog_addr_gemm();

}  // namespace fmmt
