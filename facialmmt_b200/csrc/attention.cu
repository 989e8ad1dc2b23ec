// Attention cores of the path (bf16 operands, fp32 softmax, warp-shuffle row reductions):
//   * window_attention_kernel : Swin W-MSA / SW-MSA core, N = ws*ws (49) tokens, head_dim 32,
//       softmax(q*scale @ k^T + rel_bias[h] (+ shift mask)) @ v      (Swin_Transformer.py:119-141)
//   * mha_flash_kernel        : head_dim 64 attention with online softmax over 64-key tiles, optional additive key
//       mask (1-m)*neg: HF text encoder, MELDTrans SelfAttention (Transformer.py:87-116), fairseq-style
//       MultiheadAttention core (multihead_attention.py:109-130, no masks).
// Tensor-core path here is mma.sync.m16n8k16 (tiles are 49x49x32 / 64x64x64: too small to amortise a
// TMEM round trip); the large GEMMs around them run on tcgen05 (gemm.cu).
#include "ops.cuh"
#include "ptx.cuh"

namespace fmmt {

namespace {

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_s(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans_s(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ float ex2_fast(float x) {   // 2^x, one MUFU; arguments are <= 0 here, -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------ Swin window attention
// Persistent per head: a CTA keeps the relative-position bias of its head as MMA-fragment registers and streams
// windows through a cp.async double buffer (q,k,v of one (window, head) = 3 x 49 x 64 B).
constexpr int WA_D = 32;        // head_dim (all four Swin-tiny stages)
constexpr int WA_PITCH = 40;    // smem row pitch in bf16 (80 B): conflict-free ldmatrix
constexpr int WA_ROWS = 64;     // 49 tokens padded to 4 m-tiles
constexpr int WA_NT = 7;        // key n-tiles (56 >= 49)
constexpr float LOG2E = 1.4426950408889634f;

struct WinAttnParams {
  const __nv_bfloat16* qkv;  // [num_windows*N, 3C], window order; q | k | v blocks of C, head h at h*32
  __nv_bfloat16* out;        // [num_windows*N, C]
  const float* bias;         // [heads, N, N] expanded relative-position bias
  const int8_t* rid;         // [nW, N] shift-region ids or nullptr (no mask)
  int num_windows, nW, heads, C, N;
  float scale;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128, 6) window_attention_kernel(const WinAttnParams p) {
  __shared__ __align__(16) __nv_bfloat16 sbuf[2][3][WA_ROWS * WA_PITCH];
  __shared__ int8_t sRid[2][WA_ROWS];

  const int h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N;
  const int ld = 3 * p.C;
  const bool use_mask = p.rid != nullptr;

  // zero the padding rows (>= N) of both buffers once; they are never overwritten
  for (int idx = tid; idx < 2 * 3 * WA_ROWS * 4; idx += 128) {
    const int row = (idx >> 2) % WA_ROWS;
    if (row >= N) {
      const int bm = idx / (WA_ROWS * 4);
      *reinterpret_cast<uint4*>(&sbuf[bm / 3][bm % 3][row * WA_PITCH + (idx & 3) * 8]) = make_uint4(0, 0, 0, 0);
    }
  }
  if (tid < 2 * WA_ROWS) sRid[tid / WA_ROWS][tid % WA_ROWS] = 0;

  const int r0 = warp * 16 + (lane >> 2);  // this thread's rows: r0 and r0 + 8
  const int cq = (lane & 3) * 2;
  // bias of this head in accumulator-fragment order, pre-multiplied by log2(e); key columns >= N -> -inf
  float bfr[WA_NT][4];
  {
    const float* bias_h = p.bias + static_cast<size_t>(h) * N * N;
#pragma unroll
    for (int j = 0; j < WA_NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = r0 + (e >> 1) * 8, col = j * 8 + cq + (e & 1);
        float v = -INFINITY;
        if (col < N) v = row < N ? __ldg(bias_h + row * N + col) * LOG2E : 0.f;
        bfr[j][e] = v;
      }
  }
  const float sscale = p.scale * LOG2E;

  // Loop-invariant addressing. Loads: a (window, head) is 3 matrices x N rows x 4 chunks of 16 B; chunk ids tid and
  // tid + 128 cover rows 0..63 (rows >= N are skipped). Stores: rows r0 / r0 + 8, 4 bytes per 8-column n-tile.
  const int lrow0 = tid >> 2, lrow1 = lrow0 + 32, lch = tid & 3;
  const bool lv0 = lrow0 < N, lv1 = lrow1 < N;
  const __nv_bfloat16* gsrc = p.qkv + static_cast<size_t>(lrow0) * ld + h * WA_D + lch * 8;
  const size_t g_row32 = static_cast<size_t>(32) * ld;
  const size_t g_win = static_cast<size_t>(N) * ld;
  const uint32_t s_dst0 = smem_u32(&sbuf[0][0][lrow0 * WA_PITCH + lch * 8]);
  constexpr uint32_t S_MAT = WA_ROWS * WA_PITCH * 2, S_BUF = 3 * S_MAT, S_ROW32 = 32 * WA_PITCH * 2;
  auto issue = [&](int wb, int b) {
    const __nv_bfloat16* g = gsrc + static_cast<size_t>(wb) * g_win;
    const uint32_t d = s_dst0 + static_cast<uint32_t>(b) * S_BUF;
#pragma unroll
    for (int mat = 0; mat < 3; ++mat) {
      if (lv0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + mat * S_MAT), "l"(g + mat * p.C) : "memory");
      if (lv1)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + mat * S_MAT + S_ROW32), "l"(g + mat * p.C + g_row32)
                     : "memory");
    }
    cp_async_commit();
  };
  __nv_bfloat16* gout = p.out + static_cast<size_t>(r0) * p.C + h * WA_D + cq;
  const size_t o_win = static_cast<size_t>(N) * p.C;
  const size_t o_row8 = static_cast<size_t>(8) * p.C;
  const bool sv0 = r0 < N, sv1 = r0 + 8 < N;
  // ldmatrix source offsets (bytes inside one matrix)
  const uint32_t q_off = ((warp * 16 + (lane & 15)) * WA_PITCH + (lane >> 4) * 8) * 2;
  const uint32_t k_off = (((lane & 7) + (lane >> 4) * 8) * WA_PITCH + ((lane >> 3) & 1) * 8) * 2;
  const uint32_t v_off = (((lane & 7) + ((lane >> 3) & 1) * 8) * WA_PITCH + (lane >> 4) * 8) * 2;
  const uint32_t s_base = smem_u32(&sbuf[0][0][0]);

  int wb = blockIdx.x;
  int b = 0;
  if (wb < p.num_windows) issue(wb, 0);
  for (; wb < p.num_windows; wb += gridDim.x, b ^= 1) {
    const int nxt = wb + gridDim.x;
    if (nxt < p.num_windows) {
      issue(nxt, b ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    int differs = 0;
    if (use_mask && tid < N) {
      const int8_t* rw = p.rid + (wb % p.nW) * N;
      const int8_t mine = rw[tid];
      sRid[b][tid] = mine;
      differs = mine != rw[0];
    }
    // one barrier: the tile is visible, and (masked blocks) whether this window has more than one shift region
    const bool masked = __syncthreads_or(differs) != 0;
    const uint32_t sQ = s_base + static_cast<uint32_t>(b) * S_BUF, sK = sQ + S_MAT, sV = sK + S_MAT;

    // S = Q K^T  (16 x 56 per warp)
    float s[WA_NT][4];
#pragma unroll
    for (int j = 0; j < WA_NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < WA_D / 16; ++kk) {
      uint32_t a[4];
      ldmatrix_x4_s(a, sQ + q_off + kk * 32);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // key n-tile pairs (0,1) (2,3) (4,5) (6,-)
        uint32_t bb[4];
        ldmatrix_x4_s(bb, sK + k_off + jp * (16 * WA_PITCH * 2) + kk * 32);
        mma_bf16(s[2 * jp], a, bb[0], bb[1]);
        if (jp < 3) mma_bf16(s[jp < 3 ? 2 * jp + 1 : 0], a, bb[2], bb[3]);
      }
    }
    // log2-domain logits: s*scale*log2e + bias*log2e (+ mask); row max; exp2
#pragma unroll
    for (int j = 0; j < WA_NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = fmaf(s[j][e], sscale, bfr[j][e]);
    if (masked) {   // uniform over the CTA: only the last window row / column of a shifted block get here
      const int8_t rr0 = sRid[b][r0], rr1 = sRid[b][(r0 + 8) & 63];
#pragma unroll
      for (int j = 0; j < WA_NT; ++j) {
        const int8_t c0 = sRid[b][j * 8 + cq], c1 = sRid[b][j * 8 + cq + 1];
        s[j][0] += (rr0 != c0 ? -100.f * LOG2E : 0.f);
        s[j][1] += (rr0 != c1 ? -100.f * LOG2E : 0.f);
        s[j][2] += (rr1 != c0 ? -100.f * LOG2E : 0.f);
        s[j][3] += (rr1 != c1 ? -100.f * LOG2E : 0.f);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < WA_NT; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    }
#pragma unroll
    for (int j = 0; j < WA_NT; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = ex2_fast(s[j][e] - mx[e >> 1]);
        s[j][e] = v;
        sum[e >> 1] += v;
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 1);
      sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 2);
    }

    // O = P V  (16 x 32 per warp), P re-used from the S accumulators as bf16 A fragments
    float o[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16-key tiles; tile 3 = keys 48..63 (n-tile 6 + zeros)
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      if (kk < 3) {
        a[2] = pack_bf16(s[kk < 3 ? 2 * kk + 1 : 0][0], s[kk < 3 ? 2 * kk + 1 : 0][1]);
        a[3] = pack_bf16(s[kk < 3 ? 2 * kk + 1 : 0][2], s[kk < 3 ? 2 * kk + 1 : 0][3]);
      } else {
        a[2] = 0u;
        a[3] = 0u;
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {  // dim n-tile pairs (0,1), (2,3)
        uint32_t bb[4];
        ldmatrix_x4_trans_s(bb, sV + v_off + kk * (16 * WA_PITCH * 2) + np * 32);
        mma_bf16(o[2 * np], a, bb[0], bb[1]);
        mma_bf16(o[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    const float inv0 = __fdividef(1.f, sum[0]), inv1 = __fdividef(1.f, sum[1]);
    __nv_bfloat16* go = gout + static_cast<size_t>(wb) * o_win;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (sv0) *reinterpret_cast<uint32_t*>(go + j * 8) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
      if (sv1) *reinterpret_cast<uint32_t*>(go + o_row8 + j * 8) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
    }
    __syncthreads();  // every warp is done with buffer b before the next iteration refills it
  }
}

// ------------------------------------------------------------------------------------------ general MHA (d = 64)
constexpr int FA_D = 64;
constexpr int FA_PITCH = 72;  // 144 B rows: conflict-free ldmatrix
constexpr int FA_BQ = 64;
constexpr int FA_BK = 64;

struct MhaParams {
  const __nv_bfloat16* q; int ldq;   // row (b*Lq + i), col h*64 + d
  const __nv_bfloat16* k; int ldk;   // row (b*Lk + j)
  const __nv_bfloat16* v; int ldv;
  __nv_bfloat16* out; int ldo;       // row (b*Lq + i), col h*64 + d
  const float* key_mask;             // [B, Lk] of 0/1 or nullptr
  float mask_neg;                    // additive value for masked keys: (1 - m) * mask_neg
  int B, H, Lq, Lk;
  float scale;
};

__global__ void __launch_bounds__(128) mha_flash_kernel(const MhaParams p) {
  __shared__ __align__(16) __nv_bfloat16 sQ[FA_BQ * FA_PITCH];
  __shared__ __align__(16) __nv_bfloat16 sK[FA_BK * FA_PITCH];
  __shared__ __align__(16) __nv_bfloat16 sV[FA_BK * FA_PITCH];
  __shared__ float sMask[FA_BK];

  const int q0 = blockIdx.x * FA_BQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < FA_BQ * 8; idx += 128) {
    const int row = idx >> 3, ch = idx & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + row < p.Lq)
      v = *reinterpret_cast<const uint4*>(p.q + (static_cast<size_t>(b) * p.Lq + q0 + row) * p.ldq + h * FA_D + ch * 8);
    *reinterpret_cast<uint4*>(sQ + row * FA_PITCH + ch * 8) = v;
  }
  __syncthreads();
  uint32_t qf[FA_D / 16][4];
#pragma unroll
  for (int kk = 0; kk < FA_D / 16; ++kk)
    ldmatrix_x4(qf[kk], sQ + (warp * 16 + (lane & 15)) * FA_PITCH + kk * 16 + (lane >> 4) * 8);

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const int cq = (lane & 3) * 2;

  for (int k0 = 0; k0 < p.Lk; k0 += FA_BK) {
    __syncthreads();  // previous tile fully consumed
    for (int idx = tid; idx < 2 * FA_BK * 8; idx += 128) {
      const int mat = idx / (FA_BK * 8);
      const int rem = idx - mat * (FA_BK * 8);
      const int row = rem >> 3, ch = rem & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (k0 + row < p.Lk) {
        const __nv_bfloat16* src = mat == 0 ? p.k + (static_cast<size_t>(b) * p.Lk + k0 + row) * p.ldk
                                            : p.v + (static_cast<size_t>(b) * p.Lk + k0 + row) * p.ldv;
        v = *reinterpret_cast<const uint4*>(src + h * FA_D + ch * 8);
      }
      *reinterpret_cast<uint4*>((mat == 0 ? sK : sV) + row * FA_PITCH + ch * 8) = v;
    }
    if (tid < FA_BK) {
      float add = 0.f;
      if (k0 + tid >= p.Lk) add = -INFINITY;
      else if (p.key_mask != nullptr) add = (1.0f - p.key_mask[static_cast<size_t>(b) * p.Lk + k0 + tid]) * p.mask_neg;
      sMask[tid] = add;
    }
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < FA_D / 16; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t bb[4];
        ldmatrix_x4(bb, sK + (jp * 16 + (lane & 7) + (lane >> 4) * 8) * FA_PITCH + kk * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16(s[2 * jp], qf[kk], bb[0], bb[1]);
        mma_bf16(s[2 * jp + 1], qf[kk], bb[2], bb[3]);
      }
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = s[j][e] * p.scale + sMask[j * 8 + cq + (e & 1)];
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float alpha[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
      alpha[hh] = (m_run[hh] == -INFINITY) ? 0.f : __expf(m_run[hh] - mx[hh]);
      m_run[hh] = mx[hh];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = __expf(s[j][e] - mx[e >> 1]);
        s[j][e] = v;
        rs[e >> 1] += v;
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 1);
      rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 2);
      l_run[hh] = l_run[hh] * alpha[hh] + rs[hh];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
      o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
    }
#pragma unroll
    for (int kk = 0; kk < FA_BK / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        ldmatrix_x4_trans(bb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * FA_PITCH + np * 16 + (lane >> 4) * 8);
        mma_bf16(o[2 * np], a, bb[0], bb[1]);
        mma_bf16(o[2 * np + 1], a, bb[2], bb[3]);
      }
    }
  }
  const int r0 = q0 + warp * 16 + (lane >> 2);
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = h * FA_D + j * 8 + cq;
    if (r0 < p.Lq)
      *reinterpret_cast<uint32_t*>(p.out + (static_cast<size_t>(b) * p.Lq + r0) * p.ldo + col) =
          pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
    if (r0 + 8 < p.Lq)
      *reinterpret_cast<uint32_t*>(p.out + (static_cast<size_t>(b) * p.Lq + r0 + 8) * p.ldo + col) =
          pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------ fp32-grade mode (SIMT)
// Plain fp32 math for the 1e-3 parity mode (the bf16 tensor-core cores above are the speed build). One thread per query
// row; K and V of the (window | key tile, head) staged in shared memory as fp32. Outputs are written as the split-bf16
// operand [hi | lo | hi] of the following Linear (see kernels.cu store_split4).
__device__ __forceinline__ void store_split1(__nv_bfloat16* row, int width, int col, float y) {
  const __nv_bfloat16 hi = __float2bfloat16(y);
  const __nv_bfloat16 lo = __float2bfloat16(y - __bfloat162float(hi));
  row[col] = hi;
  row[width + col] = lo;
  row[2 * width + col] = hi;
}

__global__ void __launch_bounds__(64) window_attention_f32_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                                  const float* __restrict__ bias,
                                                                  const int8_t* __restrict__ rid, int nW, int C, int N,
                                                                  float scale) {
  __shared__ float sK[49][WA_D + 1];
  __shared__ float sV[49][WA_D + 1];
  __shared__ int8_t sR[64];
  const int w = blockIdx.x, h = blockIdx.y, t = threadIdx.x;
  const size_t ld = static_cast<size_t>(3) * C;
  const float* base = qkv + static_cast<size_t>(w) * N * ld + h * WA_D;
  for (int idx = t; idx < N * WA_D; idx += 64) {
    const int r = idx / WA_D, d = idx - r * WA_D;
    sK[r][d] = base[r * ld + C + d];
    sV[r][d] = base[r * ld + 2 * C + d];
  }
  if (t < N) sR[t] = rid != nullptr ? rid[(w % nW) * N + t] : 0;
  __syncthreads();
  if (t >= N) return;
  float q[WA_D];
#pragma unroll
  for (int d = 0; d < WA_D; ++d) q[d] = base[t * ld + d] * scale;       // q * scale before QK^T (Swin_Transformer.py:123)
  float sc[49];
  float mx = -INFINITY;
  const float* bh = bias + (static_cast<size_t>(h) * N + t) * N;
  const int8_t myr = sR[t];
  for (int j = 0; j < N; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < WA_D; ++d) a = fmaf(q[d], sK[j][d], a);
    a += bh[j];
    if (rid != nullptr && sR[j] != myr) a += -100.0f;
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
  for (int j = 0; j < N; ++j) {
    sc[j] = expf(sc[j] - mx);
    sum += sc[j];
  }
  const float inv = 1.0f / sum;
  float o[WA_D];
#pragma unroll
  for (int d = 0; d < WA_D; ++d) o[d] = 0.f;
  for (int j = 0; j < N; ++j) {
    const float pj = sc[j] * inv;
#pragma unroll
    for (int d = 0; d < WA_D; ++d) o[d] = fmaf(pj, sV[j][d], o[d]);
  }
  __nv_bfloat16* orow = out + (static_cast<size_t>(w) * N + t) * ld;     // split row: 3C wide
#pragma unroll
  for (int d = 0; d < WA_D; ++d) store_split1(orow, C, h * WA_D + d, o[d]);
}

constexpr int F32_BK = 32;
__global__ void __launch_bounds__(64) mha_f32_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                     const float* __restrict__ v, int ldv, __nv_bfloat16* __restrict__ out,
                                                     int width, const float* __restrict__ key_mask, float mask_neg, int Lq,
                                                     int Lk, float scale) {
  __shared__ float sK[F32_BK][FA_D + 1];
  __shared__ float sV[F32_BK][FA_D + 1];
  __shared__ float sM[F32_BK];
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const int qi = q0 + t;
  const bool active = qi < Lq;
  float qr[FA_D];
#pragma unroll
  for (int d = 0; d < FA_D; ++d) qr[d] = active ? q[(static_cast<size_t>(b) * Lq + qi) * ldq + h * FA_D + d] * scale : 0.f;
  float o[FA_D];
#pragma unroll
  for (int d = 0; d < FA_D; ++d) o[d] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  for (int k0 = 0; k0 < Lk; k0 += F32_BK) {
    __syncthreads();
    for (int idx = t; idx < F32_BK * FA_D; idx += 64) {
      const int r = idx / FA_D, d = idx - r * FA_D;
      const bool ok = k0 + r < Lk;
      sK[r][d] = ok ? k[(static_cast<size_t>(b) * Lk + k0 + r) * ldk + h * FA_D + d] : 0.f;
      sV[r][d] = ok ? v[(static_cast<size_t>(b) * Lk + k0 + r) * ldv + h * FA_D + d] : 0.f;
    }
    if (t < F32_BK) {
      float add = 0.f;
      if (k0 + t >= Lk) add = -INFINITY;
      else if (key_mask != nullptr) add = (1.0f - key_mask[static_cast<size_t>(b) * Lk + k0 + t]) * mask_neg;
      sM[t] = add;
    }
    __syncthreads();
    float sc[F32_BK];
    float mx = m_run;
#pragma unroll 4
    for (int j = 0; j < F32_BK; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < FA_D; ++d) a = fmaf(qr[d], sK[j][d], a);
      a += sM[j];
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
    // a fully masked history keeps mx = -inf only if every key so far is padding (-inf); additive finite masks never do
    const float alpha = (m_run == -INFINITY) ? 0.f : expf(m_run - mx);
    m_run = mx;
    l_run *= alpha;
#pragma unroll
    for (int d = 0; d < FA_D; ++d) o[d] *= alpha;
#pragma unroll 4
    for (int j = 0; j < F32_BK; ++j) {
      const float pj = (mx == -INFINITY) ? 0.f : expf(sc[j] - mx);
      l_run += pj;
#pragma unroll
      for (int d = 0; d < FA_D; ++d) o[d] = fmaf(pj, sV[j][d], o[d]);
    }
  }
  if (!active) return;
  const float inv = 1.0f / l_run;
  __nv_bfloat16* orow = out + (static_cast<size_t>(b) * Lq + qi) * (3 * static_cast<size_t>(width));
#pragma unroll
  for (int d = 0; d < FA_D; ++d) store_split1(orow, width, h * FA_D + d, o[d] * inv);
}

}  // namespace

FMMT_DEFINE_WATCHDOG_ADDR(watchdog_addr_attn)

cudaError_t launch_window_attention_f32(const float* qkv, __nv_bfloat16* out_split, const float* bias, const int8_t* rid,
                                        int num_windows, int nW, int heads, int C, int N, float scale, cudaStream_t stream) {
  if (C != heads * WA_D || N > 49 || N < 1 || num_windows <= 0) return cudaErrorInvalidValue;
  dim3 grid(num_windows, heads);
  window_attention_f32_kernel<<<grid, 64, 0, stream>>>(qkv, out_split, bias, rid, nW, C, N, scale);
  return cudaGetLastError();
}

cudaError_t launch_mha_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                           __nv_bfloat16* out_split, int width, const float* key_mask, float mask_neg, int B, int H, int Lq,
                           int Lk, float scale, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0 || width < H * FA_D) return cudaErrorInvalidValue;
  dim3 grid((Lq + 63) / 64, H, B);
  mha_f32_kernel<<<grid, 64, 0, stream>>>(q, ldq, k, ldk, v, ldv, out_split, width, key_mask, mask_neg, Lq, Lk, scale);
  return cudaGetLastError();
}

cudaError_t launch_window_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, const float* bias,
                                    const int8_t* rid, int num_windows, int nW, int heads, int C, int N, float scale,
                                    cudaStream_t stream) {
  if (C != heads * WA_D || N > 49 || N < 1 || num_windows <= 0 || (C % 8) != 0) return cudaErrorInvalidValue;
  WinAttnParams p{qkv, out, bias, rid, num_windows, nW, heads, C, N, scale};
  // persistent over windows: exactly one wave of resident CTAs (a partial second wave would run alone at the end),
  // split across the heads (grid.y)
  static int resident = 0;   // CTAs that fit on the device at once
  if (resident == 0) {
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_attention_kernel, 128, 0) != cudaSuccess || per_sm < 1)
      per_sm = 4;
    resident = sms * per_sm;
  }
  int gx = resident / heads;
  if (gx < 1) gx = 1;
  if (gx > num_windows) gx = num_windows;
  dim3 grid(gx, heads);
  window_attention_kernel<<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_mha(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v,
                       int ldv, __nv_bfloat16* out, int ldo, const float* key_mask, float mask_neg, int B, int H,
                       int Lq, int Lk, float scale, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) return cudaErrorInvalidValue;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 2)) return cudaErrorInvalidValue;
  MhaParams p{q, ldq, k, ldk, v, ldv, out, ldo, key_mask, mask_neg, B, H, Lq, Lk, scale};
  dim3 grid((Lq + FA_BQ - 1) / FA_BQ, H, B);
  mha_flash_kernel<<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace fmmt
