/* facialmmt_b200 -- C ABI of the B200-native FacialMMT inference forward path.
 *
 * The reference (NUSTM/FacialMMT) has no FFI: its boundary for this path is the Python nn.Module API of
 * src/models.py (SwinForAffwildClassification :14-37, MultiModalTransformerForClassification :41-188,
 * meld_utt_transformer :192-223) plus the eval glue of train.py:169-234. Each entry point below names the
 * reference call it replaces. All pointers are plain device pointers unless marked HOST; no torch types.
 * Every function returns 0 on success and a negative code on failure; fmmt_last_error() gives the message.
 * Calls are asynchronous on the given cudaStream_t (passed as void*) and never synchronise the device, except
 * fmmt_finalize and the first forward at a new (larger) problem size, which (re)allocates the workspace.
 * A handle belongs to one device and is not re-entrant. The library owns packed weights + workspace only; it never
 * retains caller pointers beyond a call, and writes results into caller-allocated buffers.
 * There is NO CPU fallback: without an sm_100 device every compute entry point fails with FMMT_ERR_CUDA.
 */
#ifndef FACIALMMT_B200_H
#define FACIALMMT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(FMMT_BUILD)
#define FMMT_API __attribute__((visibility("default")))
#else
#define FMMT_API
#endif

#define FMMT_OK 0
#define FMMT_ERR_INVALID (-1) /* bad shape / pointer / alignment (the reference would assert) */
#define FMMT_ERR_CUDA (-2)    /* CUDA runtime / driver error */
#define FMMT_ERR_STATE (-3)   /* handle not finalized, weight missing, wrong model kind, ... */

/* Activation codes for fmmt_op_gemm. */
#define FMMT_ACT_NONE 0
#define FMMT_ACT_GELU 1 /* exact erf GELU (nn.GELU(), F.gelu, modules/Transformer.py:119-124) */
#define FMMT_ACT_RELU 2
#define FMMT_ACT_TANH 3

/* Which reference module a handle implements. */
#define FMMT_MODEL_SWIN_CLS 1   /* src/models.py:14-37   SwinForAffwildClassification */
#define FMMT_MODEL_MULTIMODAL 2 /* src/models.py:41-188  MultiModalTransformerForClassification */
#define FMMT_MODEL_UNIMODAL 3   /* src/models.py:192-223 meld_utt_transformer */

#define FMMT_TEXT_ROBERTA 0
#define FMMT_TEXT_BERT 1

/* Arithmetic mode of a handle (north_star: logits within 1e-2 of the fp32 reference in bf16 mode, 1e-3 in fp32 mode).
 * BF16: bf16 operands, fp32 accumulate, fp32 residual streams / LayerNorm / softmax (the speed build).
 * FP32: fp32-grade: every Linear runs on the same tcgen05 kernels with split-bf16 x3 operands (A = A_hi + A_lo,
 *       W = W_hi + W_lo; A_hi W_hi + A_lo W_hi + A_hi W_lo, 16 mantissa bits per operand, fp32 accumulate), attention
 *       cores and all element-wise math in fp32. About 4x slower; the reference's own arithmetic is fp32
 *       (src/models.py:95-188). */
#define FMMT_PRECISION_BF16 0
#define FMMT_PRECISION_FP32 1

typedef struct fmmt_handle fmmt_handle;

/* Dimensions only (mirrors swin_conf.yaml, main.py:62-83 and the HF roberta-large / bert-large configs). */
typedef struct fmmt_config {
  int32_t model; /* FMMT_MODEL_* */
  /* Swin-cls */
  int32_t img_size, patch_size, in_chans, embed_dim, num_stages;
  int32_t depths[4], num_heads[4];
  int32_t window_size, mlp_ratio; /* mlp_ratio as integer (4) */
  int32_t feat_dim, head_hidden, num_labels;
  int32_t swin_chunk;      /* frames per pass through stages 1-2 (L2-resident working set); 0 = default */
  int32_t swin_chunk_late; /* frames per pass through stages 3-4; 0 = default */
  /* text encoder */
  int32_t text_kind, vocab_size, text_hidden, text_layers, text_heads, text_ffn, max_pos, type_vocab, pad_id;
  float text_eps;
  /* fusion */
  int32_t hidden, heads, ffn, audio_dim, vision_dim, audio_layers, vision_layers;
  int32_t cmt_layers_ta, cmt_heads_ta, cmt_layers_tav, cmt_heads_tav;
  int32_t text_len, audio_len, vision_len;
  float eps;
  int32_t precision; /* FMMT_PRECISION_* */
} fmmt_config;

FMMT_API const char* fmmt_last_error(void);
FMMT_API const char* fmmt_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
FMMT_API int64_t fmmt_launch_count(void);

/* ---- lifecycle (replaces nn.Module construction + load_state_dict / torch.load, train.py:428-432) ---- */
FMMT_API int fmmt_create(const fmmt_config* cfg, fmmt_handle** out);
FMMT_API void fmmt_destroy(fmmt_handle* h);
/* Stage one tensor of the reference state_dict under its reference key name. `data` is a HOST pointer to contiguous
 * fp32; shape/ndim as in the state_dict. Keys the path does not use are stored and ignored; integer buffers such as
 * relative_position_index / num_batches_tracked / position_ids are recomputed and need not be passed. A key that
 * fmmt_finalize needs and does not find makes it fail with FMMT_ERR_STATE naming the key. */
FMMT_API int fmmt_load_weight(fmmt_handle* h, const char* ref_key, const float* data, const int64_t* shape, int ndim);
/* Pack to device: bf16 K-major weights, fused QKV, BatchNorm folded into Linear(37632,512), relative-position bias
 * expanded to (heads,49,49), shift-region ids, window/merge gather maps, sinusoid table. Synchronous. */
FMMT_API int fmmt_finalize(fmmt_handle* h);

/* ---- model-level forwards ---- */

/* SwinForAffwildClassification.forward (src/models.py:26-37) on `n_frames` frames, fp32 NCHW (F,3,224,224).
 * gumbel: (F,labels) explicit noise g so that probs = softmax((logits+g)/tau) == F.gumbel_softmax (NULL: g = 0).
 * Outputs (any may be NULL): logits (F,labels) raw; probs (F,labels); importance (F) = sum_c p_c^2 (train.py:183-184);
 * feat (F,feat_dim) = backbone output after BatchNorm. */
FMMT_API int fmmt_swin_forward(fmmt_handle* h, const float* frames, int n_frames, const float* gumbel, float tau,
                               float* logits, float* probs, float* importance, float* feat, void* stream);

/* Same forward fed with DECODED uint8 face crops instead of the fp32 tensor: crops (F, crop_h, crop_w, 3) uint8 on the device,
 * exactly the bytes cv2.imread returns (B,G,R). Replaces utils/dataset.py:47-69 from_image_to_embedding_no_IncepRes + the
 * host->device copy of its 602 KB/frame fp32 result: cv2.resize to 224x224 (INTER_CUBIC if crop_h < 224, INTER_AREA if
 * crop_h > 224; decided on the height only, like the reference), ToTensor, Normalize(0.5, 0.5), bit-exact with OpenCV's own
 * (non-IPP) arithmetic, fused with PatchEmbed's unfold. All crops of one call share one size (group ragged crops by size).
 * FMMT_ERR_INVALID for sizes the reference itself fails on (crop_h == 224 with crop_w != 224) or > 10x area shrinks. */
FMMT_API int fmmt_swin_forward_u8(fmmt_handle* h, const uint8_t* crops, int n_frames, int crop_h, int crop_w,
                                  const float* gumbel, float tau, float* logits, float* probs, float* importance, float* feat,
                                  void* stream);

/* Frame filter + compaction of the eval glue (train.py:185-232). frame_off: device int32 [U+1], prefix sums of the
 * per-utterance frame counts into `probs` (total_frames rows). per_utterance=1: each utterance decides the
 * "no frame passes" fallback on its own (== the reference at its batch size 1); 0: literal whole-batch decision.
 * out_v: (U,Lv,D+labels), out_mask: (U,Lv). scratch: device int32[1] (needed when per_utterance=0). */
FMMT_API int fmmt_filter_pack(const float* vision, const float* vision_mask, const int32_t* frame_off, int total_frames,
                              const float* probs, float threshold, int per_utterance, float* out_v, float* out_mask,
                              int32_t* scratch, int U, int Lv, int D, int labels, void* stream);

/* MultiModalTransformerForClassification.forward (src/models.py:95-188). ids/mask/sep_mask: int64 (U,L);
 * audio fp32 (U,La,audio_dim); audio_mask fp32 (U,La); vision fp32 (U,Lv,vision_dim+labels); vision_mask fp32 (U,Lv);
 * idx_in_dia int64 (U); logits fp32 (U,labels). */
FMMT_API int fmmt_multimodal_forward(fmmt_handle* h, const int64_t* ids, const int64_t* mask, const int64_t* sep_mask,
                                     const float* audio, const float* audio_mask, const float* vision,
                                     const float* vision_mask, const int64_t* idx_in_dia, int U, int L, float* logits,
                                     void* stream);

/* Same forward with the dialogues of the batch de-duplicated: MELD encodes every utterance with its whole
 * dialogue (src/meld_bert_extraText.py:65-130), so consecutive utterances of a batch carry IDENTICAL ids/mask rows. ids/mask
 * hold the n_dialogues distinct rows, int64 (n_dialogues, L); dialogue_of_utt: device int32 (U), the row of utterance u;
 * sep_mask stays per utterance (U, L). The text encoder then runs n_dialogues rows instead of U: identical results (rows are
 * independent in eval), ~10x fewer text FLOPs on MELD. */
FMMT_API int fmmt_multimodal_forward_dedup(fmmt_handle* h, const int64_t* ids, const int64_t* mask, int n_dialogues,
                                           const int32_t* dialogue_of_utt, const int64_t* sep_mask, const float* audio,
                                           const float* audio_mask, const float* vision, const float* vision_mask,
                                           const int64_t* idx_in_dia, int U, int L, float* logits, void* stream);

/* meld_utt_transformer.forward (src/models.py:209-223): inputs fp32 (U,Lv,vision_dim), utt_mask fp32 (U,Lv). */
FMMT_API int fmmt_unimodal_forward(fmmt_handle* h, const float* inputs, const float* utt_mask, int U, float* logits,
                                   void* stream);

/* Error surfacing for the asynchronous forwards (the reference raises Python exceptions synchronously). Every kernel
 * pipeline wait is bounded by a watchdog; a timed-out wait terminates the kernel with invalid results instead of hanging
 * the GPU. fmmt_check synchronises the stream of the handle's last forward and returns FMMT_ERR_CUDA if the watchdog fired
 * in any forward since the previous check (fmmt_last_error names barrier / CTA / thread), else FMMT_OK. The next
 * fmmt_*_forward on the handle reports the same condition if fmmt_check was not called. Callers check before they use
 * logits (evaluate.py, bench.py and smoke() do). */
FMMT_API int fmmt_check(fmmt_handle* h);

/* CUDA-graph replay (off by default). When enabled, the second forward of a handle with IDENTICAL arguments (same pointers,
 * sizes, scalars and stream) is captured into a CUDA graph and every later identical call replays it with a single launch
 * (results are the same kernels on the same buffers; the data the pointers hold may change between calls). Meant for
 * steady-state loops over pre-allocated buffers, e.g. the reference's default trg_batch_size = 1 eval loop (main.py:56),
 * which is launch-bound. Captures / profiling bypass the graph; at most 16 argument sets are cached per handle. */
FMMT_API int fmmt_set_graph(fmmt_handle* h, int enable);

/* Stage-wise parity hook: during the next forwards, copy the named fp32 intermediate into `dst` (device, `count`
 * floats). name == NULL clears all captures. Names: swin.patch_embed, swin.layer<l>.block<b>, swin.feat,
 * mm.text768, mm.text, mm.audio, mm.vision, mm.ta, mm.fused. */
FMMT_API int fmmt_set_capture(fmmt_handle* h, const char* name, float* dst, int64_t count);
/* Per-kernel timing with CUDA events on the launch stream (two records per launch while enabled). fmmt_profile_read
 * synchronises and writes a JSON object {"<kernel key>": {"ms","flops","bytes","launches"}, ...} aggregated since
 * profiling was enabled; returns the number of bytes needed (including the terminating NUL). */
FMMT_API int fmmt_set_profile(fmmt_handle* h, int enable);
FMMT_API int64_t fmmt_profile_read(fmmt_handle* h, char* buf, int64_t buf_len);
/* Pipeline watchdog word for the operator-level entry points (fmmt_op_*), which have no handle: non-zero if an mbarrier
 * wait inside a kernel timed out since the last reset. Bit 31 set, bits 24-30 barrier id, 12-23 CTA, 0-11 thread.
 * Synchronises the device. Model-level forwards report through fmmt_check instead. */
FMMT_API uint32_t fmmt_debug_timeout(int reset);
/* Layout probe (tests): one accumulator group D[128 x ncols] = sum_k A_k B_k on tcgen05 with caller-built shared-memory
 * operand images (device pointers; copied to 1024-byte aligned shared memory) and descriptor templates (all descriptor
 * fields except the start address); a_off/b_off: start offsets inside the images; a_step/b_step: descriptor address
 * increments (16-byte units) per k-step; idesc: instruction descriptor. out: device fp32 [128, ncols]. Synchronous.
 * Pins the operand layouts of the fused attention kernel (tests/test_umma_layouts_gpu.py). */
FMMT_API int fmmt_debug_umma(const void* a_img, int a_bytes, const void* b_img, int b_bytes, uint64_t adesc_tpl,
                             uint64_t bdesc_tpl, uint32_t a_off, uint32_t b_off, uint32_t idesc, int ksteps, int a_step,
                             int b_step, int ncols, float* out);
/* Bench probe: cycles per tcgen05.mma (M=128, N=n, K=16, bf16) issued back to back on resident shared-memory tiles, all
 * SMs at once (the tensor-pipe floor the GEMM roofline fractions are read against). Bits 16.. of n select a variant:
 * 1 = tcgen05.commit after every 4 MMAs, 2 = tcgen05.fence before every 4, 4 = alternate two operand tile sets. Synchronous. */
FMMT_API double fmmt_debug_mma_cycles(int n, int iters);
/* Bench probe: TMA feed rate from L2 (out2[0] = bytes per cycle per SM) with `nstage` 64 x box_rows bf16 boxes in flight on
 * `grid` CTAs; mode 1 keeps the tensor pipe busy beside it and reports cycles per N=256 MMA in out2[1]. Synchronous. */
FMMT_API int fmmt_debug_feed(int iters, int nstage, int box_rows, int mode, int grid, double* out2);
/* Bench probe 2: same with `nthr` (1..4) independent issuing threads and a matrix row pitch of `pitch_elems` bf16
 * (64 = boxes contiguous in memory); returns bytes per cycle per SM. Synchronous. */
FMMT_API double fmmt_debug_feed2(int iters, int nstage, int box_rows, int pitch_elems, int nthr, int grid);
/* Algorithmic FLOPs (2*MAC of every GEMM/attention launched) accumulated by the handle since the last reset. */
FMMT_API double fmmt_flops(fmmt_handle* h, int reset);
/* Bytes of device memory held by the handle (weights + workspace). */
FMMT_API int64_t fmmt_device_bytes(fmmt_handle* h);

/* ---- operator-level entry points (one CUDA kernel each); used by the per-kernel parity tests ---- */

/* nn.Linear: out[dest(r),:] = act(A[r,:] @ W^T + bias) + residual[dest(r),:].  A [M,lda] bf16, W [N,ldw] bf16 (the
 * nn.Linear.weight layout), fp32 accumulate on tcgen05 tensor cores. row_map (device int32, length map_period) is
 * optional: dest(r) = (r / map_period) * map_period + row_map[r % map_period]. Either output may be NULL. */
FMMT_API int fmmt_op_gemm(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K,
                          const float* bias, int act, const float* residual, int ldr, float* out_f32, int ldo32,
                          void* out_bf16, int ldo16, const int* row_map, int map_period, int block_n, void* stream);

/* Linear + LayerNorm over the N output columns in ONE kernel (the LayerNorm runs in the GEMM epilogue, where a thread owns a
 * whole accumulator row): out = LayerNorm(A W^T + bias) * gamma + beta, fp32 [M, ldo]. N <= 256, N % 32 == 0. Used for
 * PatchEmbed: Conv2d(3,96,k4,s4) as a K = 48 GEMM followed by norm_layer(96) (Swin_Transformer.py:402-412). */
FMMT_API int fmmt_op_gemm_ln(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K, const float* bias,
                             const float* gamma, const float* beta, float eps, float* out_f32, int ldo, void* stream);

/* LayerNorm over rows gathered from `nseg` segments (see csrc/ops.cuh LnArgs); biased variance, eps inside sqrt. */
FMMT_API int fmmt_op_layernorm(const float* in, int ld_in, int M, int nseg, int cseg, const int* map, int map_period,
                               int src_period, const float* gamma, const float* beta, float eps, float* out_f32,
                               int ld32, void* out_bf16, int ld16, void* stream);

/* Swin window attention core (Swin_Transformer.py:119-141): qkv bf16 [num_windows*N, 3C] in window order,
 * bias fp32 [heads,N,N], rid int8 [nW,N] shift-region ids or NULL, out bf16 [num_windows*N, C]. head_dim = 32. */
FMMT_API int fmmt_op_window_attention(const void* qkv_bf16, void* out_bf16, const float* bias, const int8_t* rid,
                                      int num_windows, int nW, int heads, int C, int N, float scale, void* stream);

/* Fused Swin MLP half-block for C = 96, hidden = 384 (Swin_Transformer.py:24-30 Mlp.forward inside :264-268):
 *   x <- x + fc2(GELU_erf(fc1(LayerNorm(x; gamma, beta, eps))))        x: fp32 [M, 96] on the device, updated IN PLACE.
 * fmmt_op_swin_mlp_pack turns the reference's fc1.weight (384,96) / fc2.weight (96,384) (HOST fp32, nn.Linear layout)
 * into the 147456-byte bf16 shared-memory image the kernel keeps resident (img_dev: device buffer of that size). */
#define FMMT_MLP96_IMG_BYTES 147456
FMMT_API int fmmt_op_swin_mlp_pack(const float* fc1_w_host, const float* fc2_w_host, void* img_dev);
FMMT_API int fmmt_op_swin_mlp(float* x, int M, const float* gamma, const float* beta, float eps, const void* img_dev,
                              const float* b1, const float* b2, void* stream);

/* Fused Swin ATTENTION half-block for C = 96 / 3 heads / 7x7 windows (Swin_Transformer.py:238-264 up to the first residual,
 * :113-143 WindowAttention.forward), one tcgen05 kernel:
 *   x_out[r] = x[g(r)] + proj(softmax(scale * q k^T + rel_bias (+ shift mask)) v),   q,k,v = qkv(LayerNorm(x[g(r)]))
 * x, x_out: fp32 [M, 96] on the device, M = frames * T, T % 98 == 0; rows of x_out are in window order (49 per window);
 * gather: device int32 [T], window-order row r of a frame reads row gather[r] of that frame in x (roll + window_partition), or
 * NULL (identity; then x_out may alias x). rid: device int8 [nW, 49] shift-region ids, wflag: device int8 [nW] = window has
 * more than one region (both NULL for W-MSA). fmmt_op_swin_attn_pack turns qkv.weight (288,96), proj.weight (96,96) and
 * relative_position_bias_table (169,3) (HOST fp32) into the 73728-byte weight image and the 507-float bias table. */
#define FMMT_ATTN96_IMG_BYTES 73728
#define FMMT_ATTN96_TAB_FLOATS 507
FMMT_API int fmmt_op_swin_attn_pack(const float* qkv_w_host, const float* proj_w_host, const float* rel_table_host,
                                    void* img_dev, float* tab_dev);
FMMT_API int fmmt_op_swin_attn(const float* x, float* x_out, int M, int T, const int* gather, const float* gamma,
                               const float* beta, float eps, const void* img_dev, const float* tab_dev, const float* qkv_b,
                               const float* proj_b, const int8_t* rid, const int8_t* wflag, int nW, void* stream);

/* Same half-block for C = 192 / 384 (hidden 4C): the weights are streamed from L2 per 128-row tile instead of being
 * resident. w1 = fc1.weight as bf16 [copies * 4C, ldw1], w2 = fc2.weight as bf16 [copies * C, ldw2] (nn.Linear layout,
 * device pointers; `copies` >= 1 identical matrices stacked along the rows, CTA b reads copy b % copies so that the
 * lock-step weight walk of all CTAs spreads over the L2 slices). */
FMMT_API int fmmt_op_swin_mlp_stream(float* x, int M, int C, const float* gamma, const float* beta, float eps,
                                     const void* w1_bf16, int ldw1, const float* b1, const void* w2_bf16, int ldw2,
                                     const float* b2, int copies, void* stream);

/* The same half-block on CTA PAIRS (tcgen05 cta_group::2, clusters of two CTAs on 256 rows: every CTA holds its 128 rows of
 * the activations and half of every weight chunk, so the weight bytes streamed per SM halve). Same arguments and results. */
FMMT_API int fmmt_op_swin_mlp_pair(float* x, int M, int C, const float* gamma, const float* beta, float eps,
                                   const void* w1_bf16, int ldw1, const float* b1, const void* w2_bf16, int ldw2,
                                   const float* b2, void* stream);

/* norm1 + roll + window_partition + WindowAttention's qkv Linear of a Swin block with C = 192 / 384 as ONE tcgen05 kernel
 * (Swin_Transformer.py:238-247 and :119):  out[r] = LayerNorm(x[g(r)]) @ W^T + bias (bf16 [M, ldo]),  x_raw[r] = x[g(r)] (fp32, the
 * block's residual stream in window order; may be NULL when gather is NULL). gather: device int32 [T] or NULL (identity),
 * M % T == 0 when given. w_bf16 = qkv.weight as bf16 [N, ldw] (nn.Linear layout), N % 64 == 0. flags: 0. */
FMMT_API int fmmt_op_ln_qkv(const float* x, float* x_raw, int M, int C, int T, const int* gather, const float* gamma,
                            const float* beta, float eps, const void* w_bf16, int ldw, const float* bias, int N,
                            void* out_bf16, int ldo, int flags, void* stream);

/* Multi-head attention core, head_dim 64: softmax(scale * q k^T + (1 - key_mask) * mask_neg) v.
 * q rows (b*Lq+i), k/v rows (b*Lk+j), head h at columns [64h, 64h+64). key_mask fp32 (B,Lk) of 0/1 or NULL. */
FMMT_API int fmmt_op_mha(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                         const float* key_mask, float mask_neg, int B, int H, int Lq, int Lk, float scale, void* stream);

/* The frame ingest alone (parity tests): crops uint8 (F, crop_h, crop_w, 3) -> out fp32 (F, 3, 224, 224) as the reference's
 * DataLoader yields it (utils/dataset.py:47-69). */
FMMT_API int fmmt_op_frame_ingest(const uint8_t* crops, int n_frames, int crop_h, int crop_w, float* out_f32, void* stream);

/* Target-utterance span extraction (src/models.py:112-150): text fp32 (U,L,H), sep_mask int64 (U,L), idx_in_dia int64 (U)
 * -> out fp32 (U,max_len,H) zero-filled past the span, out_mask fp32 (U,max_len) of 0/1. text_kind FMMT_TEXT_ROBERTA: the
 * span of utterance p > 0 starts 2 after the previous separator (<s> A </s></s> B </s>), FMMT_TEXT_BERT: 1 after. */
FMMT_API int fmmt_op_span_extract(const float* text, const int64_t* sep_mask, const int64_t* idx_in_dia, int U, int L,
                                  int H, int max_len, int text_kind, float* out, float* out_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FACIALMMT_B200_H */
