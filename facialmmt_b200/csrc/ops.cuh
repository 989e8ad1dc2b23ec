// Launchers of the non-GEMM kernels (kernels.cu, attention.cu). All asynchronous on `stream`.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fmmt {

// ---- attention.cu
cudaError_t launch_window_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, const float* bias,
                                    const int8_t* rid, int num_windows, int nW, int heads, int C, int N, float scale,
                                    cudaStream_t stream);
// fp32-grade mode variants (SIMT fp32 math): fp32 inputs; the output is the split-bf16 operand [hi | lo | hi] of the next
// Linear (row pitch ldo = 3 * width)
cudaError_t launch_window_attention_f32(const float* qkv, __nv_bfloat16* out_split, const float* bias, const int8_t* rid,
                                        int num_windows, int nW, int heads, int C, int N, float scale, cudaStream_t stream);
cudaError_t launch_mha_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                           __nv_bfloat16* out_split, int width, const float* key_mask, float mask_neg, int B, int H, int Lq,
                           int Lk, float scale, cudaStream_t stream);
cudaError_t launch_mha(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v,
                       int ldv, __nv_bfloat16* out, int ldo, const float* key_mask, float mask_neg, int B, int H,
                       int Lq, int Lk, float scale, cudaStream_t stream);

unsigned int* watchdog_addr_attn();   // device address of attention.cu's pipeline-watchdog word (ptx.cuh)

// ---- kernels.cu
// LayerNorm over rows assembled from `nseg` source segments of `cseg` channels each (C = nseg*cseg <= 1536):
//   out row r, segment s comes from source row  (r / map_period) * src_period + map[(r % map_period) * nseg + s]
//   (map == nullptr: identity, nseg must be 1). Biased variance, eps inside the sqrt (nn.LayerNorm and the TF-style
//   LayerNorm of modules/Transformer.py:57-61 are the same formula). Either output may be null.
//   Output rows may be redirected into a slice of a wider buffer: dest = (r / rows_in) * rows_out + row_off + r % rows_in.
struct LnArgs {
  const float* in = nullptr;
  int ld_in = 0;
  int M = 0;            // output rows
  int nseg = 1, cseg = 0;
  const int* map = nullptr;
  int map_period = 0, src_period = 0;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-5f;
  float* out_f32 = nullptr;
  int ld32 = 0;
  __nv_bfloat16* out_bf16 = nullptr;
  int ld16 = 0;
  int rows_in = 0, rows_out = 0, row_off = 0;
  float* out_raw = nullptr;  // optional fp32 copy of the (gathered) un-normalised input rows, row r -> out_raw[r]
  int ld_raw = 0;
  int split = 0;             // fp32-grade mode: out_bf16 rows are [hi | lo | hi] (3C wide; ld16 >= 3C)
};
cudaError_t launch_layernorm(const LnArgs& a, cudaStream_t stream);

// fp32 [M, C] -> bf16 [M, ld_out] (columns >= C untouched; callers zero the padding once)
cudaError_t launch_cast_bf16(const float* in, int ld_in, __nv_bfloat16* out, int ld_out, int M, int C,
                             cudaStream_t stream);

// fp32-grade mode: fp32 [M, C] -> split bf16 [M, 3*ldp] = [hi | lo | hi] (each ldp wide, zero padded past C)
cudaError_t launch_split_bf16(const float* in, int ld_in, __nv_bfloat16* out, int ldp, int M, int C, cudaStream_t stream);

// PatchEmbed im2col: frames fp32 (F,3,H,W) -> bf16 [F*(H/4)*(W/4), 48], k = c*16 + dy*4 + dx (Swin_Transformer.py:419)
//   split = 1 (fp32-grade mode): rows are [hi(48) | lo(48) | hi(48)]
cudaError_t launch_patch_im2col(const float* frames, __nv_bfloat16* out, int F, int H, int W, int split,
                                cudaStream_t stream);

// ---- ingest.cu
// Frame ingest (utils/dataset.py:47-69): uint8 crops (F, H, W, 3) as cv2.imread yields them -> cv2-exact INTER_CUBIC (H < 224)
// / INTER_AREA (H > 224) resize to 224 x 224 -> (v/255 - 0.5)/0.5 -> out_f32 (F,3,224,224) fp32 and / or the PatchEmbed im2col
// rows out_col bf16 [F*3136, 48] (split = 1: [hi|lo|hi], 144 wide). cudaErrorInvalidValue for sizes the reference rejects.
cudaError_t launch_frame_ingest(const uint8_t* crops, int F, int H, int W, float* out_f32, __nv_bfloat16* out_col, int split,
                                cudaStream_t stream);

// Swin-cls tail: feat512 -> Linear(512,64) -> ReLU -> Linear(64,7) [-> softmax((z+g)/tau), sum p^2]
//   (src/models.py:28-32, train.py:183-184). w1t is [feat, hidden] (transposed), w2 is [labels, hidden].
cudaError_t launch_swin_tail(const float* feat, int feat_dim, const float* w1t, const float* b1, int hidden,
                             const float* w2, const float* b2, int labels, const float* gumbel, float tau,
                             float* logits, float* probs, float* importance, int F, cudaStream_t stream);

// Frame filter + segmented compaction (train.py:185-232). frame_off: int32 [U+1] prefix sums of frames per utterance.
cudaError_t launch_filter_pack(const float* vision, const float* vision_mask, const int* frame_off, int total_frames,
                               const float* probs, float threshold, int per_utterance, float* out_v, float* out_mask,
                               int* any_kept_scratch, int U, int Lv, int D, int labels, cudaStream_t stream);

// HF embeddings: word + position + token_type[0] -> LayerNorm (transformers modeling_{bert,roberta}.py Embeddings)
cudaError_t launch_text_embed(const int64_t* ids, int* pos_scratch, int U, int L, int kind_roberta, int pad_id,
                              const float* word, const float* pos, const float* type0, int max_pos, int vocab,
                              const float* gamma, const float* beta, float eps, int D, float* out_f32,
                              __nv_bfloat16* out_bf16, int split, cudaStream_t stream);

// Utterance span extraction (src/models.py:112-150): text [U,L,H] -> out [U,max_len,H] (+0/1 mask)
//   text_row (optional, int32 [U]): utterance u slices row text_row[u] of `text` (dialogues de-duplicated across the batch)
cudaError_t launch_span_extract(const float* text, const int64_t* sep_mask, const int64_t* idx_in_dia, const int* text_row,
                                int U, int L, int H, int max_len, int gap, float* out, float* out_mask, cudaStream_t stream);

// CrossModal embed: out = sqrt(H)*x + sinusoid[pos], pos = t+1 if x[...,0] != 0 else 0 (position_embedding.py:8-27)
cudaError_t launch_cmt_embed(const float* x, int rows_in, int rows_total, int row_off, const float* table, int U,
                             int L, int H, float scale, float* out, cudaStream_t stream);

// AdditiveAttention tail + classifier (modules/Transformer.py:34-43; src/models.py:186-187):
//   score_t = wv . th[t] + bv (th = tanh(P x + Q q) from the GEMM), mask -> -inf, softmax, y = sum a_t x_t, logits = Wc y + bc
//   th (bf16) or th_f32 (fp32-grade mode): exactly one of the two
cudaError_t launch_pool_classify(const float* x, const __nv_bfloat16* th, const float* th_f32, const float* mask,
                                 const float* wv, float bv, const float* wc, const float* bc, int U, int L, int H,
                                 int labels, float* logits, cudaStream_t stream);

// out[q*period + t] = in[q*period + map[t]] for rows of C floats (parity captures of permuted activations)
cudaError_t launch_gather_rows(const float* in, const int* map, int period, int C, int M, float* out,
                               cudaStream_t stream);
// Pipeline watchdog (ptx.cuh): *out = first non-zero of the `n` per-translation-unit words, which are cleared.
cudaError_t launch_collect_status(unsigned int* const* addrs, int n, unsigned int* out, cudaStream_t stream);
cudaError_t launch_cast_i64_f32(const int64_t* in, float* out, int n, cudaStream_t stream);

// Concatenate 0/1 masks along time: out[u] = [a[u] | b[u] | c[u]]
cudaError_t launch_concat_masks(const float* a, int la, const float* b, int lb, const float* c, int lc, float* out,
                                int U, cudaStream_t stream);

}  // namespace fmmt
