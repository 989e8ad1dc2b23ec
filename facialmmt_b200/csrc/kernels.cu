// HBM-bound kernels of the path: LayerNorm (+ window / patch-merge gathers), casts, patch im2col, Swin-cls tail,
// frame filter + compaction, text embeddings, span extraction, cross-modal embedding, additive-attention pooling.
// All are vectorised (16-byte accesses where the layout allows), warp-shuffle reductions, fp32 math.
#include <cooperative_groups.h>
#include "ops.cuh"
#include "ptx.cuh"

namespace fmmt {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
constexpr int LN_NV = 12;  // float4 per lane -> C <= 1536

struct LnParams {
  const float* in; int ld_in; int M; int nseg; int cseg;
  const int* map; int map_period; int src_period;
  const float* gamma; const float* beta; float eps;
  float* out_f32; int ld32; __nv_bfloat16* out_bf16; int ld16;
  int rows_in, rows_out, row_off;
  float* out_raw; int ld_raw;
  int split;   // fp32-grade mode: the bf16 output row is [hi | lo | hi] (3C wide), hi = bf16(y), lo = bf16(y - hi)
};

// Split-bf16 operand of the fp32-grade mode: with weights stored as [hi | hi | lo] the K' = 3K product is
// A_hi W_hi + A_lo W_hi + A_hi W_lo, i.e. both operands carry 16 mantissa bits (the dropped lo*lo term is 2^-18 relative).
__device__ __forceinline__ void store_split4(__nv_bfloat16* row, int C, int col, float4 y) {
  const __nv_bfloat16 h0 = __float2bfloat16(y.x), h1 = __float2bfloat16(y.y), h2 = __float2bfloat16(y.z), h3 = __float2bfloat16(y.w);
  const uint2 hi = make_uint2(pack_bf16(__bfloat162float(h0), __bfloat162float(h1)), pack_bf16(__bfloat162float(h2), __bfloat162float(h3)));
  const uint2 lo = make_uint2(pack_bf16(y.x - __bfloat162float(h0), y.y - __bfloat162float(h1)),
                              pack_bf16(y.z - __bfloat162float(h2), y.w - __bfloat162float(h3)));
  *reinterpret_cast<uint2*>(row + col) = hi;
  *reinterpret_cast<uint2*>(row + C + col) = lo;
  *reinterpret_cast<uint2*>(row + 2 * C + col) = hi;
}

__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams a) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= a.M) return;
  const int vps = a.cseg >> 2;       // float4 per segment
  const int nvec = vps * a.nseg;
  const int C = a.cseg * a.nseg;
  long long base_q = 0;
  int rin = r;
  if (a.map != nullptr) {
    const int q = r / a.map_period;
    rin = r - q * a.map_period;
    base_q = static_cast<long long>(q) * a.src_period;
  }
  float4 v[LN_NV];
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < LN_NV; ++t) {
    const int i = lane + 32 * t;
    v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nvec) {
      const int seg = i / vps;
      const int off = i - seg * vps;
      const long long src = (a.map != nullptr) ? base_q + a.map[rin * a.nseg + seg] : r;
      v[t] = *reinterpret_cast<const float4*>(a.in + src * a.ld_in + off * 4);
      s += v[t].x + v[t].y + v[t].z + v[t].w;
    }
  }
  const float mean = warp_sum(s) / C;
  float q2 = 0.f;
#pragma unroll
  for (int t = 0; t < LN_NV; ++t) {
    if (lane + 32 * t < nvec) {
      const float dx = v[t].x - mean, dy = v[t].y - mean, dz = v[t].z - mean, dw = v[t].w - mean;
      q2 += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q2) / C + a.eps);
  long long dest = r;
  if (a.rows_in > 0) {
    const int q = r / a.rows_in;
    dest = static_cast<long long>(q) * a.rows_out + a.row_off + (r - q * a.rows_in);
  }
#pragma unroll
  for (int t = 0; t < LN_NV; ++t) {
    const int i = lane + 32 * t;
    if (i < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta) + i);
      float4 y;
      y.x = (v[t].x - mean) * rstd * g.x + b.x;
      y.y = (v[t].y - mean) * rstd * g.y + b.y;
      y.z = (v[t].z - mean) * rstd * g.z + b.z;
      y.w = (v[t].w - mean) * rstd * g.w + b.w;
      if (a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + dest * a.ld32 + i * 4) = y;
      if (a.out_bf16 != nullptr) {
        if (a.split) store_split4(a.out_bf16 + dest * a.ld16, C, i * 4, y);
        else *reinterpret_cast<uint2*>(a.out_bf16 + dest * a.ld16 + i * 4) = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      }
      if (a.out_raw != nullptr) *reinterpret_cast<float4*>(a.out_raw + static_cast<long long>(r) * a.ld_raw + i * 4) = v[t];
    }
  }
}

// LPR lanes cooperate on one row (32/LPR rows per warp), VPL float4 per lane: C == LPR * VPL * 4 exactly.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const LnParams a) {
  constexpr int RPW = 32 / LPR;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (static_cast<long long>(gwarp) * RPW >= a.M) return;  // warp-uniform
  const int sub = lane / LPR, sl = lane % LPR;
  long long r = static_cast<long long>(gwarp) * RPW + sub;
  const bool ok = r < a.M;
  if (!ok) r = a.M - 1;  // keep the shuffles uniform; stores are predicated
  const int vps = a.cseg >> 2;
  constexpr int C = LPR * VPL * 4;
  long long base_q = 0;
  int rin = static_cast<int>(r);
  if (a.map != nullptr) {
    const int q = static_cast<int>(r / a.map_period);
    rin = static_cast<int>(r - static_cast<long long>(q) * a.map_period);
    base_q = static_cast<long long>(q) * a.src_period;
  }
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const int i = sl + LPR * t;
    int seg = 0, off = i;
    if (a.nseg > 1) { seg = i / vps; off = i - seg * vps; }
    const long long src = (a.map != nullptr) ? base_q + __ldg(a.map + rin * a.nseg + seg) : r;
    v[t] = *reinterpret_cast<const float4*>(a.in + src * a.ld_in + off * 4);
    s += v[t].x + v[t].y + v[t].z + v[t].w;
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / C);
  float q2 = 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const float dx = v[t].x - mean, dy = v[t].y - mean, dz = v[t].z - mean, dw = v[t].w - mean;
    q2 += dx * dx + dy * dy + dz * dz + dw * dw;
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
  const float rstd = 1.0f / sqrtf(q2 * (1.0f / C) + a.eps);
  if (!ok) return;
  long long dest = r;
  if (a.rows_in > 0) {
    const long long q = r / a.rows_in;
    dest = q * a.rows_out + a.row_off + (r - q * a.rows_in);
  }
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const int i = sl + LPR * t;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta) + i);
    float4 y;
    y.x = (v[t].x - mean) * rstd * g.x + b.x;
    y.y = (v[t].y - mean) * rstd * g.y + b.y;
    y.z = (v[t].z - mean) * rstd * g.z + b.z;
    y.w = (v[t].w - mean) * rstd * g.w + b.w;
    if (a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + dest * a.ld32 + i * 4) = y;
    if (a.out_bf16 != nullptr) {
      if (a.split) store_split4(a.out_bf16 + dest * a.ld16, C, i * 4, y);
      else *reinterpret_cast<uint2*>(a.out_bf16 + dest * a.ld16 + i * 4) = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
    }
    if (a.out_raw != nullptr) *reinterpret_cast<float4*>(a.out_raw + r * a.ld_raw + i * 4) = v[t];
  }
}

template <int LPR, int VPL>
static cudaError_t launch_ln_vec(const LnParams& p, cudaStream_t stream) {
  constexpr int RPW = 32 / LPR;
  const long long warps = (static_cast<long long>(p.M) + RPW - 1) / RPW;
  const int grid = static_cast<int>((warps + 7) / 8);
  layernorm_vec_kernel<LPR, VPL><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ cast
__global__ void cast_bf16_kernel(const float* __restrict__ in, int ld_in, __nv_bfloat16* __restrict__ out, int ld_out,
                                 int M, int C) {
  const long long total = static_cast<long long>(M) * ld_out;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / ld_out;
    const int col = static_cast<int>(idx - row * ld_out);
    out[idx] = __float2bfloat16(col < C ? in[row * ld_in + col] : 0.f);
  }
}

// fp32 [M, C] (row pitch ld_in) -> split bf16 [M, 3 * ldp]: [hi (ldp) | lo (ldp) | hi (ldp)], columns >= C zero
__global__ void split_bf16_kernel(const float* __restrict__ in, int ld_in, __nv_bfloat16* __restrict__ out, int ldp, int M,
                                  int C) {
  const long long total = static_cast<long long>(M) * ldp;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / ldp;
    const int col = static_cast<int>(idx - row * ldp);
    const float x = col < C ? in[row * ld_in + col] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(x);
    const __nv_bfloat16 lo = __float2bfloat16(x - __bfloat162float(hi));
    __nv_bfloat16* o = out + row * (3LL * ldp);
    o[col] = hi;
    o[ldp + col] = lo;
    o[2 * ldp + col] = hi;
  }
}

// ------------------------------------------------------------------------------------------------ patch im2col
__global__ void __launch_bounds__(256) patch_im2col_kernel(const float* __restrict__ frames,
                                                           __nv_bfloat16* __restrict__ out, int H, int W, int split) {
  extern __shared__ float s_img[];  // [3][4][W]
  const int PH = H >> 2, PW = W >> 2;
  const int f = blockIdx.x / PH, py = blockIdx.x - f * PH;
  const int wv = W >> 2;  // float4 per image row
  for (int idx = threadIdx.x; idx < 12 * wv; idx += blockDim.x) {
    const int cr = idx / wv, x4 = idx - cr * wv;
    const int c = cr >> 2, dy = cr & 3;
    const float4 v = *reinterpret_cast<const float4*>(
        frames + ((static_cast<size_t>(f) * 3 + c) * H + py * 4 + dy) * W + x4 * 4);
    *reinterpret_cast<float4*>(s_img + cr * W + x4 * 4) = v;
  }
  __syncthreads();
  const int pitch = split ? 144 : 48;
  __nv_bfloat16* dst = out + (static_cast<size_t>(f) * PH + py) * PW * pitch;
  for (int idx = threadIdx.x; idx < PW * 6; idx += blockDim.x) {
    const int px = idx / 6, j = idx - px * 6;
    const int c = j >> 1, dy0 = (j & 1) * 2;
    const float* r0 = s_img + (c * 4 + dy0) * W + px * 4;
    const float* r1 = r0 + W;
    if (!split) {
      *reinterpret_cast<uint4*>(dst + px * 48 + j * 8) =
          make_uint4(pack_bf16(r0[0], r0[1]), pack_bf16(r0[2], r0[3]), pack_bf16(r1[0], r1[1]), pack_bf16(r1[2], r1[3]));
    } else {   // fp32-grade mode: [hi(48) | lo(48) | hi(48)]
      store_split4(dst + px * 144, 48, j * 8, make_float4(r0[0], r0[1], r0[2], r0[3]));
      store_split4(dst + px * 144, 48, j * 8 + 4, make_float4(r1[0], r1[1], r1[2], r1[3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------ Swin-cls tail
__global__ void swin_tail_kernel(const float* __restrict__ feat, int feat_dim, const float* __restrict__ w1t,
                                 const float* __restrict__ b1, int hidden, const float* __restrict__ w2,
                                 const float* __restrict__ b2, int labels, const float* __restrict__ gumbel, float tau,
                                 float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ importance) {
  extern __shared__ float sm[];  // feat[feat_dim] | hid[hidden] | z[labels]
  float* s_feat = sm;
  float* s_hid = sm + feat_dim;
  float* s_z = s_hid + hidden;
  const int f = blockIdx.x;
  for (int i = threadIdx.x; i < feat_dim; i += blockDim.x) s_feat[i] = feat[static_cast<size_t>(f) * feat_dim + i];
  __syncthreads();
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
    float acc = b1[j];
    for (int k = 0; k < feat_dim; ++k) acc = fmaf(s_feat[k], w1t[k * hidden + j], acc);
    s_hid[j] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < labels; c += blockDim.x) {
    float acc = b2[c];
    for (int j = 0; j < hidden; ++j) acc = fmaf(s_hid[j], w2[c * hidden + j], acc);
    s_z[c] = acc;
    if (logits != nullptr) logits[static_cast<size_t>(f) * labels + c] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0 && (probs != nullptr || importance != nullptr)) {
    float mx = -INFINITY;
    for (int c = 0; c < labels; ++c) {
      const float g = gumbel != nullptr ? gumbel[static_cast<size_t>(f) * labels + c] : 0.f;
      s_z[c] = (s_z[c] + g) / tau;
      mx = fmaxf(mx, s_z[c]);
    }
    float sum = 0.f;
    for (int c = 0; c < labels; ++c) {
      s_z[c] = expf(s_z[c] - mx);
      sum += s_z[c];
    }
    float imp = 0.f;
    for (int c = 0; c < labels; ++c) {
      const float pr = s_z[c] / sum;
      if (probs != nullptr) probs[static_cast<size_t>(f) * labels + c] = pr;
      imp = fmaf(pr, pr, imp);
    }
    if (importance != nullptr) importance[f] = imp;
  }
}

// ------------------------------------------------------------------------------------------------ filter + pack
__global__ void filter_any_kernel(const float* __restrict__ probs, int total_frames, int labels, float threshold,
                                  int* __restrict__ any_kept) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= total_frames) return;
  float imp = 0.f;
  for (int c = 0; c < labels; ++c) {
    const float pr = probs[static_cast<size_t>(f) * labels + c];
    imp = fmaf(pr, pr, imp);
  }
  if (imp > threshold) atomicOr(any_kept, 1);
}

__global__ void __launch_bounds__(256) filter_pack_kernel(const float* __restrict__ vision,
                                                          const float* __restrict__ vision_mask,
                                                          const int* __restrict__ frame_off,
                                                          const float* __restrict__ probs, float threshold,
                                                          int per_utterance, const int* __restrict__ any_kept,
                                                          float* __restrict__ out_v, float* __restrict__ out_mask, int Lv,
                                                          int D, int labels) {
  extern __shared__ int s_int[];  // src[Lv] (source frame of packed slot k, or -1)
  __shared__ int s_k;
  const int u = blockIdx.x;
  const int off = frame_off[u];
  int n = frame_off[u + 1] - off;
  if (n > Lv) n = Lv;
  if (threadIdx.x == 0) {
    int k = 0;
    for (int j = 0; j < n; ++j) {
      float imp = 0.f;
      for (int c = 0; c < labels; ++c) {
        const float pr = probs[static_cast<size_t>(off + j) * labels + c];
        imp = fmaf(pr, pr, imp);   // same accumulation order as swin_tail_kernel / filter_any_kernel
      }
      if (imp > threshold) s_int[k++] = j;
    }
    s_k = k;
  }
  __syncthreads();
  const int k = s_k;
  const bool fallback = per_utterance ? (k == 0) : (*any_kept == 0);
  const int W = D + labels;
  float* ov = out_v + static_cast<size_t>(u) * Lv * W;
  const float* vin = vision + static_cast<size_t>(u) * Lv * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int slot = warp; slot < Lv; slot += nwarps) {
    int src = -1;        // vision row to copy
    int psrc = -1;       // probability row to append
    float mval = 0.f;
    if (fallback) {
      src = slot;
      psrc = slot < n ? slot : -1;
      mval = vision_mask[static_cast<size_t>(u) * Lv + slot];
    } else if (slot < k) {
      src = s_int[slot];
      psrc = src;
      mval = 1.f;
    }
    for (int c = lane; c < D; c += 32) ov[static_cast<size_t>(slot) * W + c] = src >= 0 ? vin[static_cast<size_t>(src) * D + c] : 0.f;
    for (int c = lane; c < labels; c += 32)
      ov[static_cast<size_t>(slot) * W + D + c] = psrc >= 0 ? probs[static_cast<size_t>(off + psrc) * labels + c] : 0.f;
    if (lane == 0) out_mask[static_cast<size_t>(u) * Lv + slot] = mval;
  }
}

// ------------------------------------------------------------------------------------------------ text embeddings
__global__ void text_posids_kernel(const int64_t* __restrict__ ids, int L, int kind_roberta, int pad_id, int max_pos,
                                   int* __restrict__ pos) {
  const int u = blockIdx.x;
  if (threadIdx.x != 0) return;
  int run = 0;
  for (int t = 0; t < L; ++t) {
    int pi = t;
    if (kind_roberta) {
      const int ne = ids[static_cast<size_t>(u) * L + t] != pad_id ? 1 : 0;
      run += ne;
      pi = run * ne + pad_id;
    }
    pos[static_cast<size_t>(u) * L + t] = pi < max_pos ? pi : max_pos - 1;
  }
}

__global__ void __launch_bounds__(256) text_embed_ln_kernel(const int64_t* __restrict__ ids, const int* __restrict__ pos,
                                                            int rows, const float* __restrict__ word,
                                                            const float* __restrict__ pemb, const float* __restrict__ type0,
                                                            int vocab, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, int D,
                                                            float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                                                            int split) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  long long id = ids[r];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const int nvec = D >> 2;
  const float4* w4 = reinterpret_cast<const float4*>(word + id * D);
  const float4* p4 = reinterpret_cast<const float4*>(pemb + static_cast<long long>(pos[r]) * D);
  const float4* t4 = reinterpret_cast<const float4*>(type0);
  float4 v[LN_NV];
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < LN_NV; ++t) {
    const int i = lane + 32 * t;
    v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nvec) {
      const float4 a = __ldg(w4 + i), b = __ldg(p4 + i), c = __ldg(t4 + i);
      v[t] = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
      s += v[t].x + v[t].y + v[t].z + v[t].w;
    }
  }
  const float mean = warp_sum(s) / D;
  float q2 = 0.f;
#pragma unroll
  for (int t = 0; t < LN_NV; ++t) {
    if (lane + 32 * t < nvec) {
      const float dx = v[t].x - mean, dy = v[t].y - mean, dz = v[t].z - mean, dw = v[t].w - mean;
      q2 += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q2) / D + eps);
#pragma unroll
  for (int t = 0; t < LN_NV; ++t) {
    const int i = lane + 32 * t;
    if (i < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i);
      float4 y;
      y.x = (v[t].x - mean) * rstd * g.x + b.x;
      y.y = (v[t].y - mean) * rstd * g.y + b.y;
      y.z = (v[t].z - mean) * rstd * g.z + b.z;
      y.w = (v[t].w - mean) * rstd * g.w + b.w;
      if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(r) * D + i * 4) = y;
      if (out_bf16 != nullptr) {
        if (split) store_split4(out_bf16 + static_cast<size_t>(r) * 3 * D, D, i * 4, y);
        else *reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(r) * D + i * 4) =
            make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ span extraction
__global__ void __launch_bounds__(256) span_extract_kernel(const float* __restrict__ text,
                                                           const int64_t* __restrict__ sep_mask,
                                                           const int64_t* __restrict__ idx_in_dia,
                                                           const int* __restrict__ text_row, int L, int H,
                                                           int max_len, int gap, float* __restrict__ out,
                                                           float* __restrict__ out_mask) {
  __shared__ int s_start, s_n;
  const int u = blockIdx.x;
  const int tr = text_row != nullptr ? text_row[u] : u;   // de-duplicated dialogues: row of `text` that holds u's dialogue
  if (threadIdx.x == 0) {
    const long long p = idx_in_dia[u];
    int start = 0, n = 0, seen = 0, prev = -1;
    for (int t = 0; t < L; ++t) {
      if (sep_mask[static_cast<size_t>(u) * L + t] == 1) {
        if (seen == p) {
          if (p == 0) { start = 1; n = t - 1; }
          else { start = prev + gap; n = t - prev - gap; }
          break;
        }
        prev = t;
        ++seen;
      }
    }
    if (n < 0) n = 0;
    if (n > max_len) n = max_len;
    s_start = start;
    s_n = n;
  }
  __syncthreads();
  const int start = s_start, n = s_n;
  const int hv = H >> 2;
  for (int idx = threadIdx.x; idx < max_len * hv; idx += blockDim.x) {
    const int row = idx / hv, c4 = idx - row * hv;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) v = *reinterpret_cast<const float4*>(text + (static_cast<size_t>(tr) * L + start + row) * H + c4 * 4);
    *reinterpret_cast<float4*>(out + (static_cast<size_t>(u) * max_len + row) * H + c4 * 4) = v;
  }
  for (int row = threadIdx.x; row < max_len; row += blockDim.x)
    out_mask[static_cast<size_t>(u) * max_len + row] = row < n ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------ cross-modal embed
__global__ void cmt_embed_kernel(const float* __restrict__ x, int rows_total, int row_off, const float* __restrict__ table,
                                 int L, int H, float scale, float* __restrict__ out, long long total4) {
  const int hv = H >> 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total4;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / hv;
    const int c4 = static_cast<int>(idx - r * hv);
    const int u = static_cast<int>(r / L), t = static_cast<int>(r - static_cast<long long>(u) * L);
    const float* src = x + (static_cast<size_t>(u) * rows_total + row_off + t) * H;
    const float first = __ldg(src);
    const int pos = first != 0.f ? t + 1 : 0;
    const float4 v = *reinterpret_cast<const float4*>(src + c4 * 4);
    const float4 pe = __ldg(reinterpret_cast<const float4*>(table + static_cast<size_t>(pos) * H) + c4);
    *reinterpret_cast<float4*>(out + r * H + c4 * 4) =
        make_float4(fmaf(scale, v.x, pe.x), fmaf(scale, v.y, pe.y), fmaf(scale, v.z, pe.z), fmaf(scale, v.w, pe.w));
  }
}

// ------------------------------------------------------------------------------------------------ pooling + classifier
__device__ __forceinline__ float2 ld2f(const __nv_bfloat16* p) {
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
  return make_float2(__bfloat162float(v.x), __bfloat162float(v.y));
}
__device__ __forceinline__ float2 ld2f(const float* p) { return *reinterpret_cast<const float2*>(p); }

// AdditiveAttention pooling + classifier (src/models.py:171-188), one CLUSTER of POOL_SLICES CTAs per utterance:
//   1. the L score rows are split over the CTAs (four rows in flight per warp); every score is written into the score array
//      of EVERY CTA of the cluster through distributed shared memory;
//   2. every CTA runs the softmax over the L scores (cheap) and pools ITS slice of H / POOL_SLICES columns: thread = (float4
//      column, one of eight row groups), four rows in flight, fixed-order reduction over the row groups in shared memory;
//   3. the slices are written into CTA 0's y vector (distributed shared memory), CTA 0 applies the classifier.
// Deterministic: no atomics, fixed summation order.
constexpr int POOL_SLICES = 6;
template <typename T>
__device__ __forceinline__ T* cluster_map(T* p, unsigned int rank) { return cooperative_groups::this_cluster().map_shared_rank(p, rank); }
__device__ __forceinline__ void pool_cluster_sync() { cooperative_groups::this_cluster().sync(); }

template <typename TH>
__device__ __forceinline__ float score_row(const TH* __restrict__ row, const float* __restrict__ wv, int H, int lane) {
  float acc = 0.f;
  for (int c = lane * 2; c < H; c += 64) {
    const float2 v = ld2f(row + c);
    acc = fmaf(v.x, wv[c], acc);
    acc = fmaf(v.y, wv[c + 1], acc);
  }
  return acc;
}

template <typename TH>
__global__ void __launch_bounds__(256) pool_classify_kernel(const float* __restrict__ x,
                                                            const TH* __restrict__ th,
                                                            const float* __restrict__ mask, const float* __restrict__ wv,
                                                            float bv, const float* __restrict__ wc,
                                                            const float* __restrict__ bc, int L, int H, int labels,
                                                            float* __restrict__ logits) {
  extern __shared__ __align__(16) float sm[];  // partial[8][128] | y[H] (complete only in CTA 0 of the cluster) | scores[L]
  float* s_part = sm;
  float* s_y = sm + 8 * 128;
  float* s_sc = s_y + H;
  __shared__ float s_red[8];
  const int u = blockIdx.x;
  const unsigned int slice = blockIdx.y;   // == rank in the cluster (cluster dims 1 x POOL_SLICES x 1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  pool_cluster_sync();   // every CTA of the cluster is running before its shared memory is written remotely
  // ---- 1. scores of a contiguous block of rows per CTA (four per warp at a time), broadcast to the cluster
  {
    const int per = (L + POOL_SLICES - 1) / POOL_SLICES;      // rows of this CTA: t = slice * per + i
    const int tb = static_cast<int>(slice) * per, te = min(L, tb + per);
    for (int t0 = tb + warp * 4; t0 < te; t0 += nwarps * 4) {
      float acc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = min(t0 + q, te - 1);
        acc[q] = score_row(th + (static_cast<size_t>(u) * L + t) * H, wv, H, lane);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
      if (lane < POOL_SLICES) {
        float* dst = cluster_map(s_sc, static_cast<unsigned int>(lane));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int t = t0 + q;
          if (t < te) dst[t] = mask[static_cast<size_t>(u) * L + t] == 0.f ? -INFINITY : acc[q] + bv;
        }
      }
    }
  }
  pool_cluster_sync();
  // ---- 2. softmax over the L scores
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < L; t += blockDim.x) mx = fmaxf(mx, s_sc[t]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
  for (int w = 1; w < nwarps; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const float e = expf(s_sc[t] - mx);
    s_sc[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < nwarps; ++w) sum += s_red[w];
  const float inv = 1.f / sum;
  // ---- 3. pooled slice: columns [c0, c0 + cw), cw = H / POOL_SLICES (a multiple of 4, <= 128)
  const int cw = H / POOL_SLICES;
  const int c0 = static_cast<int>(slice) * cw;
  {
    const int c4 = threadIdx.x & 31;       // float4 column of the slice
    const int lg = threadIdx.x >> 5;       // row group: t = lg, lg + 8, ...
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * c4 < cw) {
      const float* xc = x + static_cast<size_t>(u) * L * H + c0 + 4 * c4;
      for (int t0 = lg; t0 < L; t0 += 32) {
        float4 v[4];
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int t = t0 + 8 * q;
          w[q] = t < L ? s_sc[t] : 0.f;
          v[q] = t < L ? *reinterpret_cast<const float4*>(xc + static_cast<size_t>(t) * H) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a.x = fmaf(w[q], v[q].x, a.x); a.y = fmaf(w[q], v[q].y, a.y);
          a.z = fmaf(w[q], v[q].z, a.z); a.w = fmaf(w[q], v[q].w, a.w);
        }
      }
      *reinterpret_cast<float4*>(s_part + lg * 128 + 4 * c4) = a;
    }
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < cw) {
      float acc = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) acc += s_part[g * 128 + threadIdx.x];
      cluster_map(s_y, 0u)[c0 + threadIdx.x] = acc * inv;           // into CTA 0's y
    }
  }
  pool_cluster_sync();
  if (slice != 0) return;
  for (int k = warp; k < labels; k += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < H; c += 32) acc = fmaf(s_y[c], wc[k * H + c], acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[static_cast<size_t>(u) * labels + k] = acc + bc[k];
  }
}

__global__ void concat_masks_kernel(const float* a, int la, const float* b, int lb, const float* c, int lc, float* out,
                                    int U) {
  const int Lt = la + lb + lc;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < U * Lt; idx += gridDim.x * blockDim.x) {
    const int u = idx / Lt, t = idx - u * Lt;
    float v;
    if (t < la) v = a[u * la + t];
    else if (t < la + lb) v = b[u * lb + (t - la)];
    else v = c[u * lc + (t - la - lb)];
    out[idx] = v;
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ in, const int* __restrict__ map, int period, int C4,
                                   long long total4, float* __restrict__ out) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total4;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / C4;
    const int c = static_cast<int>(idx - r * C4);
    const long long q = r / period;
    const long long src = q * period + map[r - q * period];
    reinterpret_cast<float4*>(out)[idx] = reinterpret_cast<const float4*>(in)[src * C4 + c];
  }
}

__global__ void cast_i64_f32_kernel(const int64_t* __restrict__ in, float* __restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = static_cast<float>(in[i]);
}

// Pipeline-watchdog hand-off (ptx.cuh): OR every translation unit's word into *out and clear the words.
struct WatchdogAddrs { unsigned int* a[8]; int n; };
__global__ void collect_status_kernel(const WatchdogAddrs w, unsigned int* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned int v = 0;
  for (int i = 0; i < w.n; ++i)
    if (w.a[i] != nullptr) {
      const unsigned int x = atomicExch(w.a[i], 0u);
      if (v == 0) v = x;
    }
  *out = v;
}

inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

cudaError_t launch_layernorm(const LnArgs& a, cudaStream_t stream) {
  const int C = a.nseg * a.cseg;
  if (a.M <= 0 || a.cseg <= 0 || (a.cseg % 4) != 0 || C > LN_NV * 128) return cudaErrorInvalidValue;
  if (a.map == nullptr && a.nseg != 1) return cudaErrorInvalidValue;
  if ((a.ld_in % 4) != 0 || (a.out_f32 && (a.ld32 % 4) != 0) || (a.out_bf16 && (a.ld16 % 4) != 0))
    return cudaErrorInvalidValue;
  LnParams p{a.in, a.ld_in, a.M, a.nseg, a.cseg, a.map, a.map_period, a.src_period, a.gamma, a.beta, a.eps,
             a.out_f32, a.ld32, a.out_bf16, a.ld16, a.rows_in, a.rows_out, a.row_off, a.out_raw, a.ld_raw, a.split};
  switch (C) {  // specialised row shapes of the path; anything else takes the generic one-row-per-warp kernel
    case 96: return launch_ln_vec<8, 3>(p, stream);
    case 192: return launch_ln_vec<16, 3>(p, stream);
    case 384: return launch_ln_vec<32, 3>(p, stream);
    case 768: return launch_ln_vec<32, 6>(p, stream);
    case 1024: return launch_ln_vec<32, 8>(p, stream);
    case 1536: return launch_ln_vec<32, 12>(p, stream);
    default: break;
  }
  const int grid = (a.M + 7) / 8;
  layernorm_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* in, int ld_in, __nv_bfloat16* out, int ld_out, int M, int C,
                             cudaStream_t stream) {
  if (M <= 0 || C <= 0 || ld_out < C) return cudaErrorInvalidValue;
  cast_bf16_kernel<<<grid_for(static_cast<long long>(M) * ld_out, 256), 256, 0, stream>>>(in, ld_in, out, ld_out, M, C);
  return cudaGetLastError();
}

cudaError_t launch_split_bf16(const float* in, int ld_in, __nv_bfloat16* out, int ldp, int M, int C, cudaStream_t stream) {
  if (M <= 0 || C <= 0 || ldp < C) return cudaErrorInvalidValue;
  split_bf16_kernel<<<grid_for(static_cast<long long>(M) * ldp, 256), 256, 0, stream>>>(in, ld_in, out, ldp, M, C);
  return cudaGetLastError();
}

cudaError_t launch_patch_im2col(const float* frames, __nv_bfloat16* out, int F, int H, int W, int split,
                                cudaStream_t stream) {
  if (F <= 0 || (H % 4) != 0 || (W % 4) != 0 || W > 1024) return cudaErrorInvalidValue;
  patch_im2col_kernel<<<F * (H / 4), 256, 12 * W * sizeof(float), stream>>>(frames, out, H, W, split);
  return cudaGetLastError();
}

cudaError_t launch_swin_tail(const float* feat, int feat_dim, const float* w1t, const float* b1, int hidden,
                             const float* w2, const float* b2, int labels, const float* gumbel, float tau,
                             float* logits, float* probs, float* importance, int F, cudaStream_t stream) {
  if (F <= 0 || hidden > 1024 || labels > 64 || tau == 0.f) return cudaErrorInvalidValue;
  const int block = ((hidden + 31) / 32) * 32;
  const size_t smem = (feat_dim + hidden + labels) * sizeof(float);
  swin_tail_kernel<<<F, block, smem, stream>>>(feat, feat_dim, w1t, b1, hidden, w2, b2, labels, gumbel, tau, logits,
                                              probs, importance);
  return cudaGetLastError();
}

cudaError_t launch_filter_pack(const float* vision, const float* vision_mask, const int* frame_off, int total_frames,
                               const float* probs, float threshold, int per_utterance, float* out_v, float* out_mask,
                               int* any_kept_scratch, int U, int Lv, int D, int labels, cudaStream_t stream) {
  if (U <= 0 || Lv <= 0 || Lv > 4096 || total_frames < 0) return cudaErrorInvalidValue;
  if (!per_utterance) {
    // literal batch semantics (train.py:187,223): the fallback is taken only if NO frame of the whole batch passes
    if (any_kept_scratch == nullptr) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(any_kept_scratch, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    if (total_frames > 0)
      filter_any_kernel<<<(total_frames + 255) / 256, 256, 0, stream>>>(probs, total_frames, labels, threshold,
                                                                       any_kept_scratch);
  }
  filter_pack_kernel<<<U, 256, Lv * sizeof(int), stream>>>(vision, vision_mask, frame_off, probs, threshold,
                                                           per_utterance, any_kept_scratch, out_v, out_mask, Lv, D, labels);
  return cudaGetLastError();
}

cudaError_t launch_text_embed(const int64_t* ids, int* pos_scratch, int U, int L, int kind_roberta, int pad_id,
                              const float* word, const float* pos, const float* type0, int max_pos, int vocab,
                              const float* gamma, const float* beta, float eps, int D, float* out_f32,
                              __nv_bfloat16* out_bf16, int split, cudaStream_t stream) {
  if (U <= 0 || L <= 0 || (D % 4) != 0 || D > LN_NV * 128) return cudaErrorInvalidValue;
  text_posids_kernel<<<U, 32, 0, stream>>>(ids, L, kind_roberta, pad_id, max_pos, pos_scratch);
  const int rows = U * L;
  text_embed_ln_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, pos_scratch, rows, word, pos, type0, vocab, gamma, beta,
                                                          eps, D, out_f32, out_bf16, split);
  return cudaGetLastError();
}

cudaError_t launch_span_extract(const float* text, const int64_t* sep_mask, const int64_t* idx_in_dia, const int* text_row,
                                int U, int L, int H, int max_len, int gap, float* out, float* out_mask, cudaStream_t stream) {
  if (U <= 0 || (H % 4) != 0) return cudaErrorInvalidValue;
  span_extract_kernel<<<U, 256, 0, stream>>>(text, sep_mask, idx_in_dia, text_row, L, H, max_len, gap, out, out_mask);
  return cudaGetLastError();
}

cudaError_t launch_cmt_embed(const float* x, int rows_in, int rows_total, int row_off, const float* table, int U,
                             int L, int H, float scale, float* out, cudaStream_t stream) {
  if (U <= 0 || L <= 0 || rows_in != L || (H % 4) != 0) return cudaErrorInvalidValue;
  const long long total4 = static_cast<long long>(U) * L * (H / 4);
  cmt_embed_kernel<<<grid_for(total4, 256), 256, 0, stream>>>(x, rows_total, row_off, table, L, H, scale, out, total4);
  return cudaGetLastError();
}

cudaError_t launch_pool_classify(const float* x, const __nv_bfloat16* th, const float* th_f32, const float* mask,
                                 const float* wv, float bv, const float* wc, const float* bc, int U, int L, int H,
                                 int labels, float* logits, cudaStream_t stream) {
  if (U <= 0 || L <= 0 || (H % 64) != 0 || ((th == nullptr) == (th_f32 == nullptr))) return cudaErrorInvalidValue;
  const size_t smem = (L + H + 8 * 128) * sizeof(float);
  if (smem > 48 * 1024 || (H % (POOL_SLICES * 4)) != 0 || H / POOL_SLICES > 128) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(U, POOL_SLICES, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = POOL_SLICES; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (th != nullptr) return cudaLaunchKernelEx(&cfg, pool_classify_kernel<__nv_bfloat16>, x, th, mask, wv, bv, wc, bc, L, H, labels, logits);
  return cudaLaunchKernelEx(&cfg, pool_classify_kernel<float>, x, th_f32, mask, wv, bv, wc, bc, L, H, labels, logits);
}

cudaError_t launch_gather_rows(const float* in, const int* map, int period, int C, int M, float* out,
                               cudaStream_t stream) {
  if (M <= 0 || (C % 4) != 0 || period <= 0) return cudaErrorInvalidValue;
  const long long total4 = static_cast<long long>(M) * (C / 4);
  gather_rows_kernel<<<grid_for(total4, 256), 256, 0, stream>>>(in, map, period, C / 4, total4, out);
  return cudaGetLastError();
}

cudaError_t launch_collect_status(unsigned int* const* addrs, int n, unsigned int* out, cudaStream_t stream) {
  if (n < 0 || n > 8 || out == nullptr) return cudaErrorInvalidValue;
  WatchdogAddrs w{};
  w.n = n;
  for (int i = 0; i < n; ++i) w.a[i] = addrs[i];
  collect_status_kernel<<<1, 32, 0, stream>>>(w, out);
  return cudaGetLastError();
}

cudaError_t launch_cast_i64_f32(const int64_t* in, float* out, int n, cudaStream_t stream) {
  if (n <= 0) return cudaErrorInvalidValue;
  cast_i64_f32_kernel<<<grid_for(n, 256), 256, 0, stream>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t launch_concat_masks(const float* a, int la, const float* b, int lb, const float* c, int lc, float* out,
                                int U, cudaStream_t stream) {
  const int total = U * (la + lb + lc);
  concat_masks_kernel<<<grid_for(total, 256), 256, 0, stream>>>(a, la, b, lb, c, lc, out, U);
  return cudaGetLastError();
}

}  // namespace fmmt
