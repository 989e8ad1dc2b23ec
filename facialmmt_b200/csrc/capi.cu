// extern "C" boundary of libfacialmmt_b200.so (see include/facialmmt_b200.h).
#include "facialmmt_b200.h"

#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "engine.cuh"

namespace fmmt {
long long launch_count();


static int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return FMMT_OK;
  if (e == cudaErrorInvalidValue) return set_error(FMMT_ERR_INVALID, std::string(what) + ": invalid shape/alignment");
  return set_error(FMMT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace fmmt

namespace fmmt {
int umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, unsigned long long adesc_tpl,
               unsigned long long bdesc_tpl, unsigned int a_off, unsigned int b_off, unsigned int idesc, int ksteps,
               int a_step, int b_step, int ncols, float* out);
}

using namespace fmmt;

struct fmmt_handle {
  Engine* eng;
};

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

FMMT_API const char* fmmt_last_error(void) { return g_last_error.c_str(); }
FMMT_API const char* fmmt_version(void) { return "facialmmt_b200 0.1 (sm_100a)"; }
FMMT_API int64_t fmmt_launch_count(void) { return launch_count(); }

FMMT_API int fmmt_create(const fmmt_config* cfg, fmmt_handle** out) {
  if (!cfg || !out) return set_error(FMMT_ERR_INVALID, "fmmt_create: null argument");
  if (cfg->model < FMMT_MODEL_SWIN_CLS || cfg->model > FMMT_MODEL_UNIMODAL)
    return set_error(FMMT_ERR_INVALID, "fmmt_create: unknown model kind");
  fmmt_handle* h = new (std::nothrow) fmmt_handle;
  if (!h) return set_error(FMMT_ERR_STATE, "out of host memory");
  h->eng = new (std::nothrow) Engine(*cfg);
  if (!h->eng) {
    delete h;
    return set_error(FMMT_ERR_STATE, "out of host memory");
  }
  *out = h;
  return FMMT_OK;
}

FMMT_API void fmmt_destroy(fmmt_handle* h) {
  if (!h) return;
  delete h->eng;
  delete h;
}

FMMT_API int fmmt_load_weight(fmmt_handle* h, const char* ref_key, const float* data, const int64_t* shape, int ndim) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->load_weight(ref_key, data, shape, ndim);
}

FMMT_API int fmmt_finalize(fmmt_handle* h) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->finalize();
}

FMMT_API int fmmt_swin_forward(fmmt_handle* h, const float* frames, int n_frames, const float* gumbel, float tau,
                               float* logits, float* probs, float* importance, float* feat, void* stream) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->swin_forward(frames, n_frames, gumbel, tau, logits, probs, importance, feat, S(stream));
}

FMMT_API int fmmt_swin_forward_u8(fmmt_handle* h, const uint8_t* crops, int n_frames, int crop_h, int crop_w,
                                  const float* gumbel, float tau, float* logits, float* probs, float* importance, float* feat,
                                  void* stream) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->swin_forward_u8(crops, n_frames, crop_h, crop_w, gumbel, tau, logits, probs, importance, feat, S(stream));
}

FMMT_API int fmmt_op_frame_ingest(const uint8_t* crops, int n_frames, int crop_h, int crop_w, float* out_f32, void* stream) {
  if (!crops || !out_f32) return set_error(FMMT_ERR_INVALID, "fmmt_op_frame_ingest: null pointer");
  count_launch();
  return check_cuda(launch_frame_ingest(crops, n_frames, crop_h, crop_w, out_f32, nullptr, 0, S(stream)), "fmmt_op_frame_ingest");
}

FMMT_API int fmmt_filter_pack(const float* vision, const float* vision_mask, const int32_t* frame_off, int total_frames,
                              const float* probs, float threshold, int per_utterance, float* out_v, float* out_mask,
                              int32_t* scratch, int U, int Lv, int D, int labels, void* stream) {
  if (!vision || !vision_mask || !frame_off || !probs || !out_v || !out_mask)
    return set_error(FMMT_ERR_INVALID, "fmmt_filter_pack: null pointer");
  count_launch(per_utterance ? 1 : 2);
  return check_cuda(launch_filter_pack(vision, vision_mask, frame_off, total_frames, probs, threshold, per_utterance, out_v,
                                       out_mask, scratch, U, Lv, D, labels, S(stream)),
                    "fmmt_filter_pack");
}

FMMT_API int fmmt_multimodal_forward(fmmt_handle* h, const int64_t* ids, const int64_t* mask, const int64_t* sep_mask,
                                     const float* audio, const float* audio_mask, const float* vision,
                                     const float* vision_mask, const int64_t* idx_in_dia, int U, int L, float* logits,
                                     void* stream) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->multimodal_forward(ids, mask, sep_mask, audio, audio_mask, vision, vision_mask, idx_in_dia, U, L, logits,
                                    S(stream));
}

FMMT_API int fmmt_multimodal_forward_dedup(fmmt_handle* h, const int64_t* ids, const int64_t* mask, int n_dialogues,
                                           const int32_t* dialogue_of_utt, const int64_t* sep_mask, const float* audio,
                                           const float* audio_mask, const float* vision, const float* vision_mask,
                                           const int64_t* idx_in_dia, int U, int L, float* logits, void* stream) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  if (!dialogue_of_utt || n_dialogues <= 0 || n_dialogues > U)
    return set_error(FMMT_ERR_INVALID, "fmmt_multimodal_forward_dedup: dialogue_of_utt / n_dialogues");
  return h->eng->multimodal_forward(ids, mask, sep_mask, audio, audio_mask, vision, vision_mask, idx_in_dia, U, L, logits,
                                    S(stream), dialogue_of_utt, n_dialogues);
}

FMMT_API int fmmt_unimodal_forward(fmmt_handle* h, const float* inputs, const float* utt_mask, int U, float* logits,
                                   void* stream) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->unimodal_forward(inputs, utt_mask, U, logits, S(stream));
}

FMMT_API int fmmt_check(fmmt_handle* h) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->check();
}

FMMT_API int fmmt_set_graph(fmmt_handle* h, int enable) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  h->eng->set_graph(enable != 0);
  return FMMT_OK;
}

FMMT_API int fmmt_set_capture(fmmt_handle* h, const char* name, float* dst, int64_t count) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  return h->eng->set_capture(name, dst, count);
}

FMMT_API int fmmt_set_profile(fmmt_handle* h, int enable) {
  if (!h) return set_error(FMMT_ERR_INVALID, "null handle");
  h->eng->set_profile(enable != 0);
  return FMMT_OK;
}
FMMT_API int64_t fmmt_profile_read(fmmt_handle* h, char* buf, int64_t buf_len) {
  if (!h) return 0;
  const std::string js = h->eng->profile_json();
  if (buf && buf_len > 0) {
    const size_t n = js.size() < static_cast<size_t>(buf_len - 1) ? js.size() : static_cast<size_t>(buf_len - 1);
    memcpy(buf, js.data(), n);
    buf[n] = 0;
  }
  return static_cast<int64_t>(js.size() + 1);
}

FMMT_API uint32_t fmmt_debug_timeout(int reset) {
  unsigned int* addrs[7] = {watchdog_addr_gemm(), watchdog_addr_mlp96(), watchdog_addr_mlp_stream(), watchdog_addr_attn(),
                            watchdog_addr_attn96(), watchdog_addr_ln_qkv(), watchdog_addr_mlp_pair()};
  cudaDeviceSynchronize();
  uint32_t first = 0;
  for (unsigned int* a : addrs) {
    if (a == nullptr) continue;
    unsigned int v = 0;
    if (cudaMemcpy(&v, a, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) continue;
    if (v != 0 && first == 0) first = v;
    if (reset && v != 0) {
      const unsigned int z = 0;
      cudaMemcpy(a, &z, sizeof(z), cudaMemcpyHostToDevice);
    }
  }
  return first;
}

FMMT_API int fmmt_debug_umma(const void* a_img, int a_bytes, const void* b_img, int b_bytes, uint64_t adesc_tpl,
                             uint64_t bdesc_tpl, uint32_t a_off, uint32_t b_off, uint32_t idesc, int ksteps, int a_step,
                             int b_step, int ncols, float* out) {
  if (!a_img || !b_img || !out) return set_error(FMMT_ERR_INVALID, "fmmt_debug_umma: null pointer");
  const int rc = umma_probe(a_img, a_bytes, b_img, b_bytes, adesc_tpl, bdesc_tpl, a_off, b_off, idesc, ksteps, a_step, b_step,
                            ncols, out);
  if (rc != 0) return set_error(rc == -3 ? FMMT_ERR_CUDA : FMMT_ERR_INVALID, "fmmt_debug_umma failed");
  return FMMT_OK;
}

FMMT_API double fmmt_debug_mma_cycles(int n, int iters) { return mma_rate_probe(n & 0xFFFF, iters, n >> 16); }

FMMT_API int fmmt_debug_feed(int iters, int nstage, int box_rows, int mode, int grid, double* out2) {
  return feed_probe(iters, nstage, box_rows, mode, grid, out2);
}

FMMT_API double fmmt_debug_feed2(int iters, int nstage, int box_rows, int pitch_elems, int nthr, int grid) {
  return feed_probe2(iters, nstage, box_rows, pitch_elems, nthr, grid);
}

FMMT_API double fmmt_flops(fmmt_handle* h, int reset) { return h ? h->eng->flops(reset != 0) : 0.0; }
FMMT_API int64_t fmmt_device_bytes(fmmt_handle* h) { return h ? h->eng->device_bytes() : 0; }

// ------------------------------------------------------------------------------------------------ operator level
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
  // block_n: 0 auto; > 0 fixed tile width (single-CTA kernels); < 0 register-path epilogue (-1 auto, else -block_n);
  // >= 1000 CTA-pair kernel (1000 auto, else block_n - 1000); 999 = single-CTA kernels only
  if (block_n >= 1000) { a.two_cta = 1; a.block_n = block_n - 1000; }
  else if (block_n == 999) { a.two_cta = -1; a.block_n = 0; }
  else {
    a.two_cta = block_n != 0 ? -1 : 0;
    a.block_n = block_n < 0 ? (block_n == -1 ? 0 : -block_n) : block_n;
    a.force_generic = block_n < 0;
  }
  count_launch();
  return check_cuda(launch_gemm(a, S(stream)), "fmmt_op_gemm");
}

FMMT_API int fmmt_op_gemm_ln(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K, const float* bias,
                             const float* gamma, const float* beta, float eps, float* out_f32, int ldo, void* stream) {
  if (!A_bf16 || !W_bf16 || !gamma || !beta || !out_f32) return set_error(FMMT_ERR_INVALID, "fmmt_op_gemm_ln: null pointer");
  GemmArgs a;
  a.A = static_cast<const __nv_bfloat16*>(A_bf16); a.lda = lda;
  a.W = static_cast<const __nv_bfloat16*>(W_bf16); a.ldw = ldw;
  a.M = M; a.N = N; a.K = K; a.bias = bias;
  a.ln_gamma = gamma; a.ln_beta = beta; a.ln_eps = eps;
  a.out_f32 = out_f32; a.ldo32 = ldo;
  count_launch();
  return check_cuda(launch_gemm(a, S(stream)), "fmmt_op_gemm_ln");
}

FMMT_API int fmmt_op_layernorm(const float* in, int ld_in, int M, int nseg, int cseg, const int* map, int map_period,
                               int src_period, const float* gamma, const float* beta, float eps, float* out_f32,
                               int ld32, void* out_bf16, int ld16, void* stream) {
  if (!in || !gamma || !beta || (!out_f32 && !out_bf16)) return set_error(FMMT_ERR_INVALID, "fmmt_op_layernorm: null pointer");
  LnArgs a;
  a.in = in; a.ld_in = ld_in; a.M = M; a.nseg = nseg; a.cseg = cseg;
  a.map = map; a.map_period = map_period; a.src_period = src_period;
  a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.out_f32 = out_f32; a.ld32 = ld32;
  a.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16); a.ld16 = ld16;
  count_launch();
  return check_cuda(launch_layernorm(a, S(stream)), "fmmt_op_layernorm");
}

FMMT_API int fmmt_op_swin_mlp_pack(const float* fc1_w_host, const float* fc2_w_host, void* img_dev) {
  if (!fc1_w_host || !fc2_w_host || !img_dev) return set_error(FMMT_ERR_INVALID, "fmmt_op_swin_mlp_pack: null pointer");
  std::vector<__nv_bfloat16> img(MLP96_IMG_BYTES / sizeof(__nv_bfloat16));
  mlp96_pack_weights(fc1_w_host, fc2_w_host, img.data());
  return check_cuda(cudaMemcpy(img_dev, img.data(), MLP96_IMG_BYTES, cudaMemcpyHostToDevice), "fmmt_op_swin_mlp_pack");
}

FMMT_API int fmmt_op_swin_mlp(float* x, int M, const float* gamma, const float* beta, float eps, const void* img_dev,
                              const float* b1, const float* b2, void* stream) {
  if (!x || !gamma || !beta || !img_dev || !b1 || !b2) return set_error(FMMT_ERR_INVALID, "fmmt_op_swin_mlp: null pointer");
  Mlp96Args a;
  a.x = x; a.M = M; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.img = static_cast<const __nv_bfloat16*>(img_dev); a.b1 = b1; a.b2 = b2;
  count_launch();
  return check_cuda(launch_mlp96(a, S(stream)), "fmmt_op_swin_mlp");
}

FMMT_API int fmmt_op_swin_attn_pack(const float* qkv_w_host, const float* proj_w_host, const float* rel_table_host,
                                    void* img_dev, float* tab_dev) {
  if (!qkv_w_host || !proj_w_host || !rel_table_host || !img_dev || !tab_dev)
    return set_error(FMMT_ERR_INVALID, "fmmt_op_swin_attn_pack: null pointer");
  std::vector<__nv_bfloat16> img(ATTN96_IMG_BYTES / sizeof(__nv_bfloat16));
  std::vector<float> tab(ATTN96_TAB_FLOATS);
  attn96_pack(qkv_w_host, proj_w_host, rel_table_host, img.data(), tab.data());
  int rc = check_cuda(cudaMemcpy(img_dev, img.data(), ATTN96_IMG_BYTES, cudaMemcpyHostToDevice), "fmmt_op_swin_attn_pack");
  if (rc != FMMT_OK) return rc;
  return check_cuda(cudaMemcpy(tab_dev, tab.data(), ATTN96_TAB_FLOATS * sizeof(float), cudaMemcpyHostToDevice),
                    "fmmt_op_swin_attn_pack");
}

FMMT_API int fmmt_op_swin_attn(const float* x, float* x_out, int M, int T, const int* gather, const float* gamma,
                               const float* beta, float eps, const void* img_dev, const float* tab_dev, const float* qkv_b,
                               const float* proj_b, const int8_t* rid, const int8_t* wflag, int nW, void* stream) {
  if (!x || !x_out || !gamma || !beta || !img_dev || !tab_dev || !qkv_b || !proj_b)
    return set_error(FMMT_ERR_INVALID, "fmmt_op_swin_attn: null pointer");
  Attn96Args a;
  a.x = x; a.x_out = x_out; a.M = M; a.T = T; a.gather = gather; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.img = static_cast<const __nv_bfloat16*>(img_dev); a.tab = tab_dev; a.qkv_b = qkv_b; a.proj_b = proj_b;
  a.rid = rid; a.wflag = wflag; a.nW = nW;
  // debug hook: nW < 0 -> `stream` carries a device trace buffer [8][32] of clock64 stamps (default stream is used)
  if (nW < 0) { a.nW = -nW; a.trace = static_cast<long long*>(stream); count_launch(); return check_cuda(launch_attn96(a, nullptr), "fmmt_op_swin_attn"); }
  count_launch();
  return check_cuda(launch_attn96(a, S(stream)), "fmmt_op_swin_attn");
}

FMMT_API int fmmt_op_swin_mlp_stream(float* x, int M, int C, const float* gamma, const float* beta, float eps,
                                     const void* w1_bf16, int ldw1, const float* b1, const void* w2_bf16, int ldw2,
                                     const float* b2, int copies, void* stream) {
  if (!x || !gamma || !beta || !w1_bf16 || !w2_bf16 || !b1 || !b2)
    return set_error(FMMT_ERR_INVALID, "fmmt_op_swin_mlp_stream: null pointer");
  MlpStreamArgs a;
  a.x = x; a.M = M; a.C = C; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.w1 = static_cast<const __nv_bfloat16*>(w1_bf16); a.ldw1 = ldw1; a.b1 = b1;
  a.w2 = static_cast<const __nv_bfloat16*>(w2_bf16); a.ldw2 = ldw2; a.b2 = b2;
  a.copies = copies < 0 ? 1 : copies;
  // debug hook: copies < 0 -> `stream` argument carries a device trace buffer instead (default stream is used)
  if (copies < 0) { a.trace = static_cast<long long*>(stream); count_launch(); return check_cuda(launch_mlp_stream(a, nullptr), "fmmt_op_swin_mlp_stream"); }
  count_launch();
  return check_cuda(launch_mlp_stream(a, S(stream)), "fmmt_op_swin_mlp_stream");
}

FMMT_API int fmmt_op_swin_mlp_pair(float* x, int M, int C, const float* gamma, const float* beta, float eps,
                                   const void* w1_bf16, int ldw1, const float* b1, const void* w2_bf16, int ldw2,
                                   const float* b2, void* stream) {
  if (!x || !gamma || !beta || !w1_bf16 || !w2_bf16 || !b1 || !b2)
    return set_error(FMMT_ERR_INVALID, "fmmt_op_swin_mlp_pair: null pointer");
  MlpStreamArgs a;
  a.x = x; a.M = M; a.C = C; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.w1 = static_cast<const __nv_bfloat16*>(w1_bf16); a.ldw1 = ldw1; a.b1 = b1;
  a.w2 = static_cast<const __nv_bfloat16*>(w2_bf16); a.ldw2 = ldw2; a.b2 = b2;
  count_launch();
  return check_cuda(launch_mlp_pair(a, S(stream)), "fmmt_op_swin_mlp_pair");
}

FMMT_API int fmmt_op_ln_qkv(const float* x, float* x_raw, int M, int C, int T, const int* gather, const float* gamma,
                            const float* beta, float eps, const void* w_bf16, int ldw, const float* bias, int N,
                            void* out_bf16, int ldo, int flags, void* stream) {
  if (!x || !gamma || !beta || !w_bf16 || !bias || !out_bf16) return set_error(FMMT_ERR_INVALID, "fmmt_op_ln_qkv: null pointer");
  LnQkvArgs a;
  a.x = x; a.x_raw = x_raw; a.M = M; a.C = C; a.T = T; a.gather = gather; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.w = static_cast<const __nv_bfloat16*>(w_bf16); a.ldw = ldw; a.bias = bias; a.N = N;
  a.out = static_cast<__nv_bfloat16*>(out_bf16); a.ldo = ldo;
  count_launch();
  a.dbg = (flags >> 1) & 3;
  // debug hook: flags & 1 -> `stream` carries a device trace buffer [8][32] of clock64 stamps (default stream is used)
  if (flags & 1) { a.trace = static_cast<long long*>(stream); return check_cuda(launch_ln_qkv(a, nullptr), "fmmt_op_ln_qkv"); }
  return check_cuda(launch_ln_qkv(a, S(stream)), "fmmt_op_ln_qkv");
}

FMMT_API int fmmt_op_window_attention(const void* qkv_bf16, void* out_bf16, const float* bias, const int8_t* rid,
                                      int num_windows, int nW, int heads, int C, int N, float scale, void* stream) {
  if (!qkv_bf16 || !out_bf16 || !bias) return set_error(FMMT_ERR_INVALID, "fmmt_op_window_attention: null pointer");
  count_launch();
  return check_cuda(launch_window_attention(static_cast<const __nv_bfloat16*>(qkv_bf16),
                                            static_cast<__nv_bfloat16*>(out_bf16), bias, rid, num_windows, nW, heads, C,
                                            N, scale, S(stream)),
                    "fmmt_op_window_attention");
}

FMMT_API int fmmt_op_mha(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                         const float* key_mask, float mask_neg, int B, int H, int Lq, int Lk, float scale, void* stream) {
  if (!q || !k || !v || !out) return set_error(FMMT_ERR_INVALID, "fmmt_op_mha: null pointer");
  count_launch();
  return check_cuda(launch_mha(static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
                               static_cast<const __nv_bfloat16*>(v), ldv, static_cast<__nv_bfloat16*>(out), ldo, key_mask,
                               mask_neg, B, H, Lq, Lk, scale, S(stream)),
                    "fmmt_op_mha");
}

FMMT_API int fmmt_op_span_extract(const float* text, const int64_t* sep_mask, const int64_t* idx_in_dia, int U, int L,
                                  int H, int max_len, int text_kind, float* out, float* out_mask, void* stream) {
  if (!text || !sep_mask || !idx_in_dia || !out || !out_mask) return set_error(FMMT_ERR_INVALID, "fmmt_op_span_extract: null pointer");
  if (text_kind != FMMT_TEXT_ROBERTA && text_kind != FMMT_TEXT_BERT) return set_error(FMMT_ERR_INVALID, "fmmt_op_span_extract: text_kind");
  count_launch();
  return check_cuda(launch_span_extract(text, sep_mask, idx_in_dia, nullptr, U, L, H, max_len, text_kind == FMMT_TEXT_ROBERTA ? 2 : 1,
                                        out, out_mask, S(stream)),
                    "fmmt_op_span_extract");
}

}  // extern "C"
