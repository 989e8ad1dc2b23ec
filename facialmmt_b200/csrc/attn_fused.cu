// Fused Swin attention half-block, C = 96 / 3 heads / 7x7 windows (Swin-tiny stage 1):
//     x_out[r] = x[g(r)] + proj( softmax( (q * scale) k^T + rel_bias (+ shift mask) ) v ),   q,k,v = qkv(LayerNorm1(x[g(r)]))
// (Swin_Transformer.py:238-264 SwinTransformerBlock.forward up to the first residual; :113-143 WindowAttention.forward;
// g(r) = the composed torch.roll + window_partition gather of this block, :244,43-44) as ONE persistent kernel. Un-fused this
// half-block was four launches (LayerNorm+gather, qkv GEMM, mma.sync window attention, proj GEMM + residual) moving
// 2 x [M,3C] bf16 + 2 x [M,C] bf16 + 4 x [M,C] fp32 through HBM; fused it reads x once and writes x_out once, and the
// 128 x 288 qkv tile, the scores and the probabilities never leave the SM.
//
// Tile = 2 windows = 98 tokens, padded to 128 accumulator rows (window w at rows 64w .. 64w+48; TMEM lane == row).
//   * qkv.weight + proj.weight (72 KB bf16) stay resident in shared memory (pre-swizzled on the host, 3 bulk copies);
//   * 8 LayerNorm warps gather the tile's rows of x (coalesced, 8 lanes per row), write them to x_out (the residual),
//     normalise and write the bf16 A operand as three 32-channel k-blocks (K-major, SWIZZLE_64B);
//   * one thread issues every tcgen05.mma: qkv (N = 192 + 96, K = 96) -> TMEM; per head S = Q_h K_h^T (M = 128, N = 128,
//     K = 32: both windows in one instruction, each row reads its own window's 64-column half); O_h = P_h V_h as two
//     N = 32, K = 64 instructions (one per window) with V_h read in place as an MN-major operand; proj (N = 96, K = 96);
//   * 12 compute warps in three groups, ONE GROUP PER HEAD, drain q_h / k_h / v_h (+bias, q * scale * log2 e) into the
//     head's Q/K/V tiles, run the softmax from TMEM (relative-position bias from a 2 KB [head][169] table, shift mask from
//     region ids, exp2, fp32), write P_h as a bf16 K-major tile, scale O_h by 1/sum into the tile that held Q_h, and hand
//     32 columns of the proj accumulator (+bias) to TMA reduce-add (cp.reduce.async.bulk.tensor .add), which applies the
//     residual in the memory system.
// TMEM columns: [0,288) qkv accumulator, re-used as three 128-column S_h buffers [0,384) whose first 64 columns are re-used
// again for O_h (2 x 32, one per window); [416,512) proj accumulator.
#include "attn_fused.cuh"

#include <mutex>
#include <vector>

#include "gemm.cuh"
#include "ptx.cuh"

namespace fmmt {

namespace {

constexpr int C = ATTN96_C;
constexpr int NTOK = ATTN96_N;
constexpr int TILE_TOK = 2 * NTOK;              // 98 valid rows per tile
constexpr int MMA_WARP = 0;
constexpr int LN_WARP0 = 1, LN_WARPS = 8;       // warps 1..8
constexpr int CW0 = LN_WARP0 + LN_WARPS;        // compute warps 9..20: group h = warps 9 + 4h .. 12 + 4h = head h
constexpr int THREADS = (CW0 + 12) * 32;        // 672

constexpr int WQKV_KB = 288 * 64;               // bytes of one 32-channel k-block of qkv.weight
constexpr int WPROJ_KB = 96 * 64;
constexpr int OFF_W = 0;
constexpr int OFF_WPROJ = 3 * WQKV_KB;          // 55296
constexpr int OFF_A = ATTN96_IMG_BYTES;         // 73728: three [128 x 64 B] k-blocks
constexpr int OFF_QKV = OFF_A + 3 * 8192;       // 98304: nine [128 x 64 B] tiles Q0 Q1 Q2 K0 K1 K2 V0 V1 V2 (O_h re-uses Q_h)
constexpr int OFF_P = OFF_QKV + 9 * 8192;       // 172032: three [128 x 128 B] probability tiles P_h; P_h doubles as group h's
                                                //         [128 x 32] fp32 output slab once its P V products are done
constexpr int OFF_TAB = OFF_P + 3 * 16384;      // 221184: bias table [3][169] fp32 (pre-multiplied by log2 e)
constexpr int OFF_VEC = OFF_TAB + 2048;         // qkv bias [288] | proj bias [96] | gamma [96] | beta [96]
constexpr int OFF_RID = OFF_VEC + 576 * 4;      // region id per tile row [128] int8
constexpr int SMEM_BYTES = OFF_RID + 128;       // 225664
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "shared memory budget");

constexpr uint32_t TM_S = 0;        // S_h at 128 h; O_h (window w) at 128 h + 32 w once the softmax has read S_h
constexpr uint32_t TM_V = 192;      // v part of the qkv accumulator
constexpr uint32_t TM_PROJ = 416;
constexpr float LOG2E = 1.4426950408889634f;

struct Attn96Params {
  const float* x; float* x_out;
  int M, T, num_tiles;
  const int* gather;
  float eps;
  const __nv_bfloat16* img;
  const float* tab; const float* qkv_b; const float* proj_b; const float* gamma; const float* beta;
  const int8_t* rid; const int8_t* wflag;
  int nW;
  float qscale;      // scale * log2(e)
  long long* trace;  // optional [8 tiles][32] clock64 stamps of CTA 0 (debug)
};

#define TRACE(slot)                                                                          \
  do {                                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0 && i < 8) p.trace[i * 32 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(THREADS, 1)
swin_attn96_fused_kernel(const __grid_constant__ CUtensorMap tmOut, const Attn96Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_bar, a_full, a_empty, qkv_full, qkv_ready, o_smem_full, proj_full, proj_drained;
  __shared__ uint64_t s_full[3], p_full[3], o_full[3], o_drained[3];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  float* s_tab = reinterpret_cast<float*>(smem + OFF_TAB);
  float* s_vec = reinterpret_cast<float*>(smem + OFF_VEC);
  int8_t* s_rid = reinterpret_cast<int8_t*>(smem + OFF_RID);

  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    mbar_init(&a_full, LN_WARPS * 32);
    mbar_init(&a_empty, 1);
    mbar_init(&qkv_full, 1);
    mbar_init(&qkv_ready, 384);
    mbar_init(&o_smem_full, 384);
    mbar_init(&proj_full, 1);
    mbar_init(&proj_drained, 384);
    for (int h = 0; h < 3; ++h) {
      mbar_init(&s_full[h], 1);
      mbar_init(&p_full[h], 128);
      mbar_init(&o_full[h], 1);
      mbar_init(&o_drained[h], 128);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  if (warp == CW0 && lane == 0) tma_prefetch_desc(&tmOut);
  // zero the A tile once (padding rows 49..63 / 113..127 stay zero: finite q/k/v rows) and stage the small tables
  for (int idx = threadIdx.x; idx < 3 * 8192 / 16; idx += THREADS) reinterpret_cast<uint4*>(smem + OFF_A)[idx] = make_uint4(0, 0, 0, 0);
  for (int idx = threadIdx.x; idx < ATTN96_TAB_FLOATS; idx += THREADS) s_tab[idx] = p.tab[idx];
  for (int idx = threadIdx.x; idx < 288; idx += THREADS) s_vec[idx] = p.qkv_b[idx];
  for (int idx = threadIdx.x; idx < 96; idx += THREADS) {
    s_vec[288 + idx] = p.proj_b[idx];
    s_vec[384 + idx] = p.gamma[idx];
    s_vec[480 + idx] = p.beta[idx];
  }
  if (threadIdx.x < 128) s_rid[threadIdx.x] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ weight fetch + every MMA (one thread)
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, ATTN96_IMG_BYTES);
      const uint8_t* img = reinterpret_cast<const uint8_t*>(p.img);
      for (int c = 0; c < 3; ++c) bulk_g2s(smem + OFF_W + c * 24576, img + c * 24576, 24576, &w_bar);
      const uint32_t id_qk = make_idesc_bf16(128, 192), id_96 = make_idesc_bf16(128, 96);
      const uint32_t id_s = make_idesc_bf16(128, 128), id_pv = make_idesc_bf16(128, 32, 1);
      auto issue_s = [&](int h) {
        const uint64_t a = make_smem_desc_sw64(smem_base + OFF_QKV + h * 8192);
        const uint64_t b = make_smem_desc_sw64(smem_base + OFF_QKV + (3 + h) * 8192);
        const uint32_t d = tmem_base + TM_S + static_cast<uint32_t>(h * 128);
        umma_bf16(d, a, b, id_s, 0u);
        umma_bf16(d, a + 2, b + 2, id_s, 1u);
        umma_commit(&s_full[h]);
      };
      // O_h = P_h V_h for the three heads and both windows: six independent accumulators, issued interleaved so that
      // consecutive instructions never wait for each other's accumulator (a dependent chain costs ~100 cycles per link)
      auto issue_pv_all = [&]() {
        uint64_t a[3], b[3][2];
#pragma unroll
        for (int h = 0; h < 3; ++h) {
          a[h] = make_smem_desc_sw128(smem_base + OFF_P + h * 16384);
#pragma unroll
          for (int w = 0; w < 2; ++w)   // V_h rows 64 w .. 64 w + 63 are window w's keys; MN-major: 16 keys = 1024 B
            b[h][w] = make_smem_desc_sw64(smem_base + OFF_QKV + (6 + h) * 8192 + w * 4096);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int h = 0; h < 3; ++h)
#pragma unroll
            for (int w = 0; w < 2; ++w)
              umma_bf16(tmem_base + TM_S + static_cast<uint32_t>(h * 128 + w * 32), a[h] + 2 * k, b[h][w] + 64 * k, id_pv,
                        k != 0 ? 1u : 0u);   // O_h over the consumed S_h
#pragma unroll
        for (int h = 0; h < 3; ++h) umma_commit(&o_full[h]);
      };
      mbar_wait(&w_bar, 0, 70);
      for (int i = 0; i < n_local; ++i) {
        const uint32_t par = i & 1u, ppar = par ^ 1u;
        TRACE(0);
        mbar_wait_relaxed(&a_full, par, 71, 1000);
        TRACE(1);
        if (i > 0) {   // S_h / O_h of the previous tile alias the qkv accumulator
#pragma unroll
          for (int h = 0; h < 3; ++h) mbar_wait(&o_drained[h], ppar, 72);
        }
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t a = make_smem_desc_sw64(smem_base + OFF_A + kb * 8192);
          const uint64_t bqk = make_smem_desc_sw64(smem_base + OFF_W + kb * WQKV_KB);
          const uint64_t bv = make_smem_desc_sw64(smem_base + OFF_W + kb * WQKV_KB + 192 * 64);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_bf16(tmem_base, a + 2 * k, bqk + 2 * k, id_qk, (kb | k) != 0 ? 1u : 0u);
            umma_bf16(tmem_base + TM_V, a + 2 * k, bv + 2 * k, id_96, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&qkv_full);
        umma_commit(&a_empty);
        TRACE(2);
        mbar_wait_relaxed(&qkv_ready, par, 73, 1000);     // every group has read its accumulator columns and written its Q/K/V tiles
        TRACE(3);
        tc_fence_after();
        issue_s(0);
        issue_s(1);
        issue_s(2);
#pragma unroll
        for (int h = 0; h < 3; ++h) {
          mbar_wait_relaxed(&p_full[h], par, 74, 1000);   // P_h written; S_h read out (its columns may take O_h)
          TRACE(4 + h);
        }
        tc_fence_after();
        issue_pv_all();
        mbar_wait_relaxed(&o_smem_full, par, 77, 1000);
        TRACE(7);
        if (i > 0) mbar_wait(&proj_drained, ppar, 78);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t a = make_smem_desc_sw64(smem_base + OFF_QKV + kb * 8192);     // O_kb (held where Q_kb was)
          const uint64_t b = make_smem_desc_sw64(smem_base + OFF_WPROJ + kb * WPROJ_KB);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16(tmem_base + TM_PROJ, a + 2 * k, b + 2 * k, id_96, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&proj_full);
        TRACE(8);
      }
    }
  } else if (warp < CW0) {
    // ------------------------------------------------------------------ gather + residual copy + LayerNorm -> A tile
    const int t = threadIdx.x - LN_WARP0 * 32;   // 0..255
    const int l8 = t & 7;
    const int tt = t >> 3;
    const int rg = 8 * (tt >> 3) + ((tt & 1) << 2) + ((tt >> 1) & 3);   // rows rg + 32 q: swizzle phases spread over the banks
    const bool copy_raw = p.x_out != p.x;
    auto src_row = [&](int row) -> int {   // composed roll + window_partition gather (within a frame of T tokens)
      if (p.gather == nullptr) return row;
      const int fr = row / p.T;
      return fr * p.T + __ldg(p.gather + (row - fr * p.T));
    };
    for (int i = 0; i < n_local; ++i) {
      const int tile = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      float4 xv[4][3];
      bool valid[4];
      int grow[4], src[4];
      // every load of the four rows is issued before the first store (x and x_out may alias as far as the compiler knows;
      // a store between two rows' loads serialises their HBM round trips)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = rg + 32 * q;
        const int tok = r & 63;
        valid[q] = tok < NTOK;
        grow[q] = tile * TILE_TOK + (r >> 6) * NTOK + tok;   // row in window order
        src[q] = valid[q] ? src_row(grow[q]) : 0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
          xv[q][j] = valid[q] ? __ldcg(reinterpret_cast<const float4*>(p.x + static_cast<size_t>(src[q]) * C + 4 * l8 + 32 * j))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (i + 1 < n_local && t < TILE_TOK)   // pull the next tile's rows towards L2 (one 384-byte row per thread)
        prefetch_l2_bulk(p.x + static_cast<size_t>(src_row((tile + static_cast<int>(gridDim.x)) * TILE_TOK + t)) * C, C * 4);
      if (copy_raw) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (valid[q]) {
#pragma unroll
            for (int j = 0; j < 3; ++j) *reinterpret_cast<float4*>(p.x_out + static_cast<size_t>(grow[q]) * C + 4 * l8 + 32 * j) = xv[q][j];
          }
        }
      }
      float rstd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) s += (xv[q][j].x + xv[q][j].y) + (xv[q][j].z + xv[q][j].w);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float mean = s * (1.0f / C);
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          xv[q][j].x -= mean; xv[q][j].y -= mean; xv[q][j].z -= mean; xv[q][j].w -= mean;
          v = fmaf(xv[q][j].x, xv[q][j].x, v); v = fmaf(xv[q][j].y, xv[q][j].y, v);
          v = fmaf(xv[q][j].z, xv[q][j].z, v); v = fmaf(xv[q][j].w, xv[q][j].w, v);
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        rstd[q] = rsqrtf(v * (1.0f / C) + p.eps);
      }
      mbar_wait_relaxed(&a_empty, (i & 1u) ^ 1u, 79);    // the qkv MMAs of the previous tile have consumed the A tile
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = 4 * l8 + 32 * j;         // channel; k-block j, 16-byte chunk l8 >> 1 of the 64-byte row
        const float4 g4 = *reinterpret_cast<const float4*>(s_vec + 384 + col);
        const float4 be4 = *reinterpret_cast<const float4*>(s_vec + 480 + col);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!valid[q]) continue;
          const int r = rg + 32 * q;
          const float o0 = fmaf(xv[q][j].x * rstd[q], g4.x, be4.x);
          const float o1 = fmaf(xv[q][j].y * rstd[q], g4.y, be4.y);
          const float o2 = fmaf(xv[q][j].z * rstd[q], g4.z, be4.z);
          const float o3 = fmaf(xv[q][j].w * rstd[q], g4.w, be4.w);
          uint8_t* dst = smem + OFF_A + j * 8192 + r * 64 + ((((l8 >> 1) ^ ((r >> 1) & 3))) << 4) + (l8 & 1) * 8;
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
        }
      }
      fence_proxy_async_smem();
      if (copy_raw) fence_proxy_async_global();   // the residual rows must be visible to the TMA reduce-add of this tile
      mbar_arrive(&a_full);
    }
  } else {
    // ------------------------------------------------------------------ compute warps: group h owns head h
    const int cw = warp - CW0;
    const int h = cw >> 2;
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int win = row >> 6, tok = row & 63;
    const bool valid = tok < NTOK;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int sw64 = (row >> 1) & 3, sw128 = row & 7;
    const bool elected = lane == 0 && (cw & 3) == 0;   // first warp of the group
    uint8_t* ptile = smem + OFF_P + h * 16384;         // P_h, later this group's output slab
    const int bias_base = (tok / 7 + 6) * 13 + (tok % 7) + 6;
    const float* tb = s_tab + h * 169 + bias_base;
    // one accumulator unit (32 columns of this row) -> + bias (-> * mul) -> bf16 -> 64-byte row of a SWIZZLE_64B tile
    auto store_unit = [&](const uint32_t (&v)[32], int u, float mul) {
      const float* bias = s_vec + 32 * u;
      uint8_t* dst = smem + OFF_QKV + u * 8192 + row * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * c);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + 8 * c + 4);
        *reinterpret_cast<uint4*>(dst + ((c ^ sw64) << 4)) = make_uint4(
            pack_bf16((__uint_as_float(v[8 * c + 0]) + b0.x) * mul, (__uint_as_float(v[8 * c + 1]) + b0.y) * mul),
            pack_bf16((__uint_as_float(v[8 * c + 2]) + b0.z) * mul, (__uint_as_float(v[8 * c + 3]) + b0.w) * mul),
            pack_bf16((__uint_as_float(v[8 * c + 4]) + b1.x) * mul, (__uint_as_float(v[8 * c + 5]) + b1.y) * mul),
            pack_bf16((__uint_as_float(v[8 * c + 6]) + b1.z) * mul, (__uint_as_float(v[8 * c + 7]) + b1.w) * mul));
      }
    };
    for (int i = 0; i < n_local; ++i) {
      const int tile = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const uint32_t par = i & 1u;
      // per-tile shift-mask inputs: region id of this row, "window has more than one region" flag
      int8_t myrid = 0;
      bool masked = false;
      if (p.rid != nullptr) {
        const int wi = (tile * 2 + win) % p.nW;
        masked = __ldg(p.wflag + wi) != 0;
        if (valid) myrid = __ldg(p.rid + wi * NTOK + tok);
      }
      // P_h doubled as the output slab of the previous tile: its reduce-add must have read it before P_h is rewritten
      if (elected) tma_store_wait_read<0>();
      named_bar_sync(1 + h, 128);
      // ---- q_h, k_h, v_h accumulator columns -> (+bias, q * scale * log2 e) -> bf16 tiles (loads run one unit ahead)
      if (elected) TRACE(10 + 7 * h);
      mbar_wait_relaxed(&qkv_full, par, 80, 1000);
      if (elected) TRACE(11 + 7 * h);
      tc_fence_after();
      if (h == 0) s_rid[row] = myrid;
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(32 * h), va);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(32 * (3 + h)), vb);
        store_unit(va, h, p.qscale);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(t_lane + static_cast<uint32_t>(32 * (6 + h)), va);
        store_unit(vb, 3 + h, 1.0f);
        tmem_ld_wait();
        store_unit(va, 6 + h, 1.0f);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&qkv_ready);
      if (elected) TRACE(12 + 7 * h);

      // ---- softmax of head h from TMEM -> P_h
      mbar_wait_relaxed(&s_full[h], par, 81, 1000);
      if (elected) TRACE(13 + 7 * h);
      tc_fence_after();
      float inv = 0.f;
      {
        uint32_t sv[56];
        const uint32_t ts = t_lane + TM_S + static_cast<uint32_t>(h * 128 + win * 64);
        tmem_ld_32x32b_x32(ts, reinterpret_cast<uint32_t(&)[32]>(sv[0]));
        tmem_ld_32x32b_x16(ts + 32u, sv + 32);
        tmem_ld_32x32b_x8(ts + 48u, sv + 48);
        tmem_ld_wait();
        tc_fence_before();      // S_h is read out: its columns may receive O_h once p_full[h] completes
        if (valid) {
          // scores are already in the log2 domain (q carries scale * log2 e; the table is pre-multiplied by log2 e)
#pragma unroll
          for (int j = 0; j < NTOK; ++j) sv[j] = __float_as_uint(__uint_as_float(sv[j]) + tb[-((j / 7) * 13 + (j % 7))]);
          if (masked) {
            const int8_t* rw = s_rid + win * 64;
#pragma unroll
            for (int j = 0; j < NTOK; ++j)
              if (rw[j] != myrid) sv[j] = __float_as_uint(__uint_as_float(sv[j]) - 100.0f * LOG2E);
          }
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < NTOK; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]));
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < NTOK; ++j) {
            const float e = ex2f(__uint_as_float(sv[j]) - mx);
            sum += e;
            sv[j] = __float_as_uint(e);
          }
          inv = __fdividef(1.0f, sum);
          uint8_t* prow = ptile + row * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t w4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j0 = 8 * c + 2 * e;
              const float lo = j0 < NTOK ? __uint_as_float(sv[j0 < NTOK ? j0 : 0]) : 0.f;
              const float hi = j0 + 1 < NTOK ? __uint_as_float(sv[j0 + 1 < NTOK ? j0 + 1 : 0]) : 0.f;
              w4[e] = pack_bf16(lo, hi);
            }
            *reinterpret_cast<uint4*>(prow + ((c ^ sw128) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&p_full[h]);
      if (elected) TRACE(14 + 7 * h);

      // ---- O_h accumulator -> * 1/sum -> bf16 -> the tile that held Q_h (S_h has consumed it): A operand of proj
      mbar_wait_relaxed(&o_full[h], par, 82, 1000);
      if (elected) TRACE(15 + 7 * h);
      tc_fence_after();
      {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_lane + TM_S + static_cast<uint32_t>(h * 128 + win * 32), ov);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_drained[h]);
        uint8_t* orow = smem + OFF_QKV + h * 8192 + row * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(orow + ((c ^ sw64) << 4)) = make_uint4(
              pack_bf16(__uint_as_float(ov[8 * c + 0]) * inv, __uint_as_float(ov[8 * c + 1]) * inv),
              pack_bf16(__uint_as_float(ov[8 * c + 2]) * inv, __uint_as_float(ov[8 * c + 3]) * inv),
              pack_bf16(__uint_as_float(ov[8 * c + 4]) * inv, __uint_as_float(ov[8 * c + 5]) * inv),
              pack_bf16(__uint_as_float(ov[8 * c + 6]) * inv, __uint_as_float(ov[8 * c + 7]) * inv));
      }
      fence_proxy_async_smem();
      mbar_arrive(&o_smem_full);
      if (elected) TRACE(16 + 7 * h);

      // ---- proj accumulator columns 32 h .. 32 h + 31 -> + bias -> slab (in P_h) -> TMA reduce-add into x_out, which holds
      //      the residual rows
      mbar_wait_relaxed(&proj_full, par, 83, 1000);
      if (elected) TRACE(31);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + TM_PROJ + static_cast<uint32_t>(32 * h), v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&proj_drained);
        const float* bias = s_vec + 288 + 32 * h;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * c);
          *reinterpret_cast<float4*>(ptile + row * 128 + ((c ^ sw128) << 4)) =
              make_float4(__uint_as_float(v[4 * c]) + b4.x, __uint_as_float(v[4 * c + 1]) + b4.y,
                          __uint_as_float(v[4 * c + 2]) + b4.z, __uint_as_float(v[4 * c + 3]) + b4.w);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + h, 128);
      if (elected) {
        const int row0 = tile * TILE_TOK;
        tma_reduce_add_f32_2d(&tmOut, ptile, 32 * h, row0);                       // window 0: slab rows 0..48
        tma_reduce_add_f32_2d(&tmOut, ptile + 64 * 128, 32 * h, row0 + NTOK);     // window 1: slab rows 64..112
        tma_store_commit();
      }
    }
    if (elected) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline size_t sw64_off(int r, int k) {   // byte offset of bf16 element (row r, column k < 32) in a SWIZZLE_64B tile
  return static_cast<size_t>(r) * 64 + ((((k >> 3) ^ ((r >> 1) & 3))) << 4) + (k & 7) * 2;
}

}  // namespace

void attn96_pack(const float* qkv_w, const float* proj_w, const float* rel_table, __nv_bfloat16* img_host, float* tab_host) {
  uint8_t* img = reinterpret_cast<uint8_t*>(img_host);
  auto put = [&](size_t off, float v) { *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16(v); };
  for (int n = 0; n < 288; ++n)
    for (int k = 0; k < C; ++k) put(OFF_W + static_cast<size_t>(k / 32) * WQKV_KB + sw64_off(n, k % 32), qkv_w[static_cast<size_t>(n) * C + k]);
  for (int n = 0; n < C; ++n)
    for (int k = 0; k < C; ++k) put(OFF_WPROJ + static_cast<size_t>(k / 32) * WPROJ_KB + sw64_off(n, k % 32), proj_w[static_cast<size_t>(n) * C + k]);
  // relative_position_bias_table (169, heads) -> [head][169], in the log2 domain
  for (int h = 0; h < ATTN96_HEADS; ++h)
    for (int i = 0; i < 169; ++i) tab_host[h * 169 + i] = rel_table[i * ATTN96_HEADS + h] * LOG2E;
}

FMMT_DEFINE_WATCHDOG_ADDR(watchdog_addr_attn96)

cudaError_t launch_attn96(const Attn96Args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.T <= 0 || (a.M % a.T) != 0 || (a.T % TILE_TOK) != 0) return cudaErrorInvalidValue;
  if (!a.x || !a.x_out || !a.gamma || !a.beta || !a.img || !a.tab || !a.qkv_b || !a.proj_b) return cudaErrorInvalidValue;
  if (a.gather != nullptr && a.x_out == a.x) return cudaErrorInvalidValue;      // a gather cannot run in place
  if (a.rid != nullptr && (a.nW <= 0 || (a.T % (a.nW * NTOK)) != 0)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.x_out) & 15) ||
      (reinterpret_cast<uintptr_t>(a.img) & 15))
    return cudaErrorInvalidValue;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 148;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(swin_attn96_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  });
  if (attr_err != cudaSuccess) return attr_err;
  CUtensorMap tmOut;
  if (!make_tmap_2d(&tmOut, a.x_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, C, C, 32, NTOK)) return cudaErrorInvalidValue;
  Attn96Params p{};
  p.x = a.x; p.x_out = a.x_out; p.M = a.M; p.T = a.T; p.num_tiles = a.M / TILE_TOK;
  p.gather = a.gather; p.eps = a.eps; p.img = a.img; p.tab = a.tab; p.qkv_b = a.qkv_b; p.proj_b = a.proj_b;
  p.gamma = a.gamma; p.beta = a.beta; p.rid = a.rid; p.wflag = a.wflag; p.nW = a.nW;
  p.qscale = a.scale * LOG2E;
  p.trace = a.trace;
  if (a.rid != nullptr && a.wflag == nullptr) return cudaErrorInvalidValue;
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  swin_attn96_fused_kernel<<<grid, THREADS, SMEM_BYTES + 1024, stream>>>(tmOut, p);
  return cudaGetLastError();
}

}  // namespace fmmt
