// Engine: owns packed device weights + a workspace arena and sequences the kernels of one forward on a stream.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "facialmmt_b200.h"
#include "attn_fused.cuh"
#include "gemm.cuh"
#include "ln_qkv.cuh"
#include "mlp_fused.cuh"
#include "mlp_pair.cuh"
#include "mlp_stream.cuh"
#include "ops.cuh"

namespace fmmt {

typedef __nv_bfloat16 bf16;

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

// Linear weight on device: bf16 [N, ld] (K contiguous, ld = K rounded up to 8, zero padded) + fp32 bias.
// fp32-grade mode (fmmt_config.precision = 1): bf16 [N, 3*ld] = [hi | hi | lo] per group of `ld` columns, so that with the
// activation stored as [hi | lo | hi] the K' = 3K product is A_hi W_hi + A_lo W_hi + A_hi W_lo (16 mantissa bits per operand).
struct Lin {
  bf16* w = nullptr;
  float* b = nullptr;
  int N = 0, K = 0, ld = 0;
};
struct Norm {
  float* g = nullptr;
  float* b = nullptr;
  int C = 0;
};

struct SwinBlockW {
  Norm ln1, ln2;
  Lin qkv, proj, fc1, fc2;
  bf16* mlp_img = nullptr;    // C == 96: fc1/fc2 pre-swizzled for the fused MLP kernel (mlp_fused.cu), else nullptr
  bf16* attn_img = nullptr;   // C == 96: qkv/proj pre-swizzled for the fused attention half-block (attn_fused.cu), else nullptr
  float* attn_tab = nullptr;  //          relative-position bias table [heads][169] * log2(e)
  float* bias_exp = nullptr;  // [heads, N, N]
  int shift = 0;
  // The residual stream is kept in the window order of the most recent attention block, so that every GEMM output
  // is written in place order (TMA-storable) and the only permutations are row gathers inside LayerNorm:
  int* gather = nullptr;      // [T] window-order row r of THIS block <- row gather[r] of the current stream order
  int* to_natural = nullptr;  // [T] natural token t sits at row to_natural[t] after this block (parity captures)
  bool identity = false;      // gather is the identity (stage with a single window): no copy needed
};
struct SwinStageW {
  int R = 0, C = 0, heads = 0, ws = 0, N = 0, nW = 0;
  int8_t* rid = nullptr;                 // [nW, N] region ids for shifted blocks
  int8_t* wflag = nullptr;               // [nW] 1 where a window spans more than one shift region
  std::vector<SwinBlockW> blocks;
  bool has_merge = false;
  int* merge_map = nullptr;  // [T/4, 4] rows of the current stream order (after the stage's last block)
  Norm merge_ln;
  Lin merge;
};
struct SwinW {
  Lin patch;  // [C0, 48]
  Norm patch_ln;
  std::vector<SwinStageW> stages;
  Norm head_ln;
  int* final_gather = nullptr;  // natural token -> stream row after the last block (nullptr: already natural)
  Lin head;  // [feat, R*R*C] with BatchNorm folded in
  float* w1t = nullptr;  // [feat, hidden]
  float* b1 = nullptr;
  float* w2 = nullptr;   // [labels, hidden]
  float* b2 = nullptr;
};

// Post-LN encoder layer (HF BERT/RoBERTa layer and MELDTrans TransformerEnoderLayer share this structure).
struct EncLayerW {
  Lin qkv, o, fc1, fc2;
  Norm ln1, ln2;
};
struct CmtLayerW {
  Lin q, kv, o, fc1, fc2;
  Norm ln0, ln1;
};
struct CmtW {
  std::vector<CmtLayerW> layers;
  Norm final_ln;
  int heads = 12;
};
struct MeldEncW {
  Lin in;                 // audio_linear / vision_linear / modality_linear
  float* pos = nullptr;   // [max_len, H]
  int max_len = 0;
  std::vector<EncLayerW> layers;
};
struct PoolW {
  Lin P;        // bias = P.bias + Q(query_vector)
  float* wv = nullptr;
  float bv = 0.f;
  float* wc = nullptr;  // [labels, H]
  float* bc = nullptr;
};
struct TextW {
  float* word = nullptr;
  float* pos = nullptr;
  float* type0 = nullptr;
  Norm emb_ln;
  std::vector<EncLayerW> layers;
  Lin out;  // text_linear
};

class Arena {
 public:
  void begin(bool dry, char* base, size_t cap) { dry_ = dry; base_ = base; cap_ = cap; off_ = 0; peak_ = 0; }
  template <typename T>
  T* alloc(size_t n) {
    off_ = (off_ + 255) & ~static_cast<size_t>(255);
    char* p = base_ + off_;
    off_ += n * sizeof(T);
    if (off_ > peak_) peak_ = off_;
    return reinterpret_cast<T*>(p);
  }
  size_t mark() const { return off_; }
  void release(size_t m) { off_ = m; }
  size_t peak() const { return peak_; }
  bool dry() const { return dry_; }
  size_t cap() const { return cap_; }

 private:
  bool dry_ = true;
  char* base_ = nullptr;
  size_t cap_ = 0, off_ = 0, peak_ = 0;
};

class Engine {
 public:
  explicit Engine(const fmmt_config& cfg);
  ~Engine();
  int load_weight(const char* key, const float* data, const int64_t* shape, int ndim);
  int finalize();
  int swin_forward(const float* frames, int F, const float* gumbel, float tau, float* logits, float* probs,
                   float* importance, float* feat, cudaStream_t st);
  int swin_forward_u8(const uint8_t* crops, int F, int crop_h, int crop_w, const float* gumbel, float tau, float* logits,
                      float* probs, float* importance, float* feat, cudaStream_t st);
  int multimodal_forward(const int64_t* ids, const int64_t* mask, const int64_t* sep, const float* audio,
                         const float* audio_mask, const float* vision, const float* vision_mask, const int64_t* idx,
                         int U, int L, float* logits, cudaStream_t st, const int32_t* text_row = nullptr, int n_text = 0);
  int unimodal_forward(const float* inputs, const float* mask, int U, float* logits, cudaStream_t st);
  int set_capture(const char* name, float* dst, int64_t count);
  // Per-kernel CUDA-event timing (on the launch stream) for roofline accounting; adds two event records per launch.
  void set_profile(bool on);
  std::string profile_json();  // synchronises; aggregates by kernel key since set_profile(true)
  // CUDA-graph replay of repeated identical forwards (same pointers / sizes / stream): see Engine::run.
  void set_graph(bool on);
  // Synchronises the stream of the last forward and reports a pipeline-watchdog event (ptx.cuh) of any forward since
  // the previous check: FMMT_OK or FMMT_ERR_CUDA (message via fmmt_last_error). Clears the condition.
  int check();
  double flops(bool reset) { double f = flops_; if (reset) flops_ = 0; return f; }
  int64_t device_bytes() const { return static_cast<int64_t>(weight_bytes_ + ws_cap_); }
  const std::string& error() const { return err_; }

 private:
  // ---- weight packing
  const HostTensor* find(const std::string& key);
  const HostTensor& need(const std::string& key);
  template <typename T> T* dev_alloc(size_t n);
  float* up_f32(const float* src, size_t n);
  bf16* up_bf16(const float* src, int rows, int cols, int ld);
  Lin make_lin(const float* w, const float* b, int N, int K, int group = 0);
  Lin lin(const std::string& prefix, bool bias = true);
  Norm norm(const std::string& prefix);
  EncLayerW enc_layer(const std::string& qkv_prefix, const std::string& o_prefix, const std::string& ln1_prefix,
                      const std::string& fc1_prefix, const std::string& fc2_prefix, const std::string& ln2_prefix,
                      int H);
  void pack_swin();
  void pack_text();
  void pack_meld(MeldEncW& m, const std::string& lin_prefix, const std::string& enc_prefix, int layers, int in_dim,
                 int max_len);
  void pack_cmt(CmtW& c, const std::string& prefix, int layers, int heads);
  void pack_pool(const std::string& prefix, const std::string& cls_prefix);
  void build_sinusoid(int max_len);

  // ---- op wrappers (no-ops while sizing the workspace)
  void gemm(GemmArgs a);
  void gemm_lin(const bf16* A, int lda, int M, const Lin& l, GemmArgs ep);
  // Linear whose output is the A operand of a later Linear: bf16 [M, N] (bf16 mode) or split bf16 [M, 3N] through an
  // fp32 scratch (fp32-grade mode)
  void lin_to_operand(const bf16* A, int lda, int M, const Lin& l, int act, bf16* out16);
  // Linear whose output feeds an attention core: bf16 [M, N] or fp32 [M, N] (fp32-grade mode); `out` is sized for either
  void lin_to_attn(const bf16* A, int lda, int M, const Lin& l, void* out);
  void attn_mha(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, bf16* out, int width,
                const float* key_mask, float mask_neg, int B, int H, int Lq, int Lk, const std::string& key);
  void split16(const float* in, int ld_in, bf16* out, int ldp, int M, int C);
  void ln(LnArgs a);
  void mlp96(float* x, int M, const SwinBlockW& bw);
  void attn96(const float* x, float* x_out, int M, const SwinStageW& sw, const SwinBlockW& bw);
  void mlp_stream(float* x, int M, int C, const SwinBlockW& bw);
  void ck(cudaError_t e, const char* what);
  void capture(const std::string& name, const float* src, size_t count, size_t dst_off = 0);

  // ---- forward bodies (run twice the first time a size is seen: dry sizing pass, then real)
  // frames: fp32 (F,3,img,img) as the reference DataLoader yields them, or uint8 crops (F,h,w,3) ingested on the device
  struct FrameSrc { const float* f32 = nullptr; const uint8_t* u8 = nullptr; int h = 0, w = 0; };
  void swin_body(const FrameSrc& frames, int F, const float* gumbel, float tau, float* logits, float* probs,
                 float* importance, float* feat);
  void swin_early(const FrameSrc& frames, int f0, int nf, float* x2_out);
  void swin_late(float* x2, int f0, int nf, bf16* feat_ln);
  void swin_block(const SwinStageW& sw, const SwinBlockW& bw, float*& x, float*& xalt, int nf, bf16* h, void* qkv,
                  bf16* a, bf16* hid);
  void capture_block(const std::string& name, const SwinStageW& sw, const SwinBlockW& bw, const float* x, int nf, int f0);
  void enc_layers(const std::vector<EncLayerW>& layers, float* x32, bf16* x16, int U, int L, int H, int heads, int ffn,
                  const float* mask01, float mask_neg, float eps);
  void meld_encoder(const MeldEncW& m, const float* in, int in_dim, int U, int L, const float* mask01, float* x32,
                    bf16* x16);
  void cmt_encoder(const CmtW& c, const float* xq, int Lq, int q_total, int q_off, const float* xkv, int Lk, int kv_total,
                   int kv_off, int U, float* out32, bf16* out16, int out_total, int out_off, bool keep_scratch = false);
  void pool_head(const float* x32, const bf16* x16, const float* mask01, int U, int L, float* logits);
  void multimodal_body(const int64_t* ids, const int64_t* mask, const int64_t* sep, const float* audio,
                       const float* audio_mask, const float* vision, const float* vision_mask, const int64_t* idx, int U,
                       int L, float* logits, const int32_t* text_row, int n_text);
  void unimodal_body(const float* inputs, const float* mask, int U, float* logits);
  template <typename Fn> int run(Fn&& body, cudaStream_t st, const std::vector<unsigned long long>& key);
  struct GraphEntry {
    std::vector<unsigned long long> key;
    cudaGraphExec_t exec = nullptr;
    int seen = 0;
    int launches = 0;
    double flops = 0;
  };
  std::unordered_map<unsigned long long, GraphEntry> graphs_;
  bool graph_on_ = false;
  cudaStream_t graph_stream_ = nullptr;   // private capture stream
  // Side streams for the independent branches of the fusion forward (audio / vision encoders beside the text encoder, the two
  // directions of a CrossmodalTransformer pair): fork_to(k) makes side stream k wait for everything issued on the forward's
  // stream so far and redirects the following launches to it; branch_done(k, main) switches back, join_wait(k) makes the main
  // stream wait for the branch. Inside a graph capture the same calls produce parallel branches of the graph. Scratch memory of
  // concurrent branches must not be released until after the join (cmt_encoder's keep_scratch).
  cudaStream_t side_[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork_ = nullptr, ev_join_[2] = {nullptr, nullptr};
  bool branches_ = true;                  // FMMT_NO_BRANCHES=1 runs the fusion forward on one stream
  bool pending_[2] = {false, false};
  cudaStream_t fork_to(int k);
  void branch_done(int k, cudaStream_t main_stream);
  void join_wait(int k);
  void drop_graphs();

  fmmt_config cfg_;
  bool precise_ = false;   // fp32-grade mode (cfg.precision == 1)
  int kw_ = 1;             // operand width multiplier: 3 in fp32-grade mode (split-bf16 x3), else 1
  size_t attn_esize() const { return precise_ ? sizeof(float) : sizeof(bf16); }
  bool finalized_ = false;
  std::unordered_map<std::string, HostTensor> host_;
  std::vector<void*> dev_ptrs_;
  size_t weight_bytes_ = 0;
  SwinW swin_;
  TextW text_;
  MeldEncW audio_, vision_;
  CmtW cmt_ta_, cmt_tav_;
  PoolW pool_;
  float* sinusoid_ = nullptr;  // [max_len+1, H]
  int sinusoid_len_ = 0;

  int device_ = -1;                       // the device that owns the weights and the workspace (set by finalize)
  unsigned int* status_dev_ = nullptr;    // device word written by launch_collect_status at the end of every forward
  unsigned int* status_host_ = nullptr;   // pinned copy (stream-ordered D2H at the end of every forward)
  int consume_status(const char* where);  // FMMT_ERR_CUDA if the pinned word is set (and clears it)
  Arena arena_;
  char* ws_ = nullptr;
  size_t ws_cap_ = 0;
  cudaStream_t st_ = nullptr;
  cudaError_t first_err_ = cudaSuccess;
  std::string err_;
  double flops_ = 0;
  struct ProfRec { std::string key; double flops; double bytes; cudaEvent_t e0, e1; };
  bool prof_ = false;
  std::vector<ProfRec> prof_recs_;
  std::vector<cudaEvent_t> ev_pool_;
  cudaEvent_t prof_begin(const std::string& key, double flops, double bytes);
  void prof_end(cudaEvent_t e1);
  cudaEvent_t get_event();
  struct Cap { float* dst; int64_t count; };
  std::map<std::string, Cap> caps_;
};

extern thread_local std::string g_last_error;
int set_error(int code, const std::string& msg);
void count_launch(int n = 1);

}  // namespace fmmt
