// Fused Swin MLP half-block, C = 96 / hidden = 384 (Swin-tiny stage 1; Swin_Transformer.py:24-30 Mlp.forward and
// :264-268 "x = x + drop_path(mlp(norm2(x)))"):
//     x <- x + fc2( GELU_erf( fc1( LayerNorm(x) ) ) )
// as ONE persistent kernel. Un-fused, this half-block moves 32*C bytes per token through HBM (LN out, fc1 in/out,
// fc2 in, residual in/out); fused it moves 8*C (x read once, x updated once) and the 128x384 hidden tile never leaves
// the SM:
//   * both weight matrices (147 KB as bf16) stay resident in shared memory for the life of the CTA, pre-swizzled on
//     the host into the tcgen05 K-major SWIZZLE_128B operand layout and fetched with three bulk copies;
//   * 8 LayerNorm warps read a 128-row tile of x (coalesced, 8 lanes per row), normalise and write the bf16 A operand
//     straight into its swizzled shared-memory tile;
//   * one thread issues tcgen05.mma: fc1 as two N=192 halves into 2 x 192 TMEM columns, fc2 as six K=64 partial
//     products into 96 TMEM columns;
//   * 12 GELU warps (3 groups) drain the fc1 accumulators 64 columns at a time (tcgen05.ld -> bias -> erf-GELU ->
//     bf16) into a double-buffered shared-memory chunk that is the A operand of the next fc2 partial product;
//   * four of the LayerNorm warps also drain the fc2 accumulator: + bias, 32-column slabs handed to TMA, which applies
//     the residual as an fp32 reduce-add into x (cp.reduce.async.bulk.tensor .add): the SM never loads the residual.
// Shared memory (bytes): W1 main 49152 | W1 tail 24576 | W2 73728 | A 32768 | hidden 2x16384 | out slab 16384.
#include "mlp_fused.cuh"

#include <mutex>
#include <vector>

#include "gemm.cuh"
#include "ptx.cuh"

namespace fmmt {

namespace {

constexpr int TILE_M = 128;
constexpr int GELU_GROUPS = 3;
constexpr int GELU_WARPS = 4 * GELU_GROUPS;   // warps 0..11
constexpr int LN_WARP0 = GELU_WARPS;          // warps 12..19: LayerNorm; 12..15 (warp % 4 == TMEM lane quarter) also
constexpr int LN_WARPS = 8;                   //   drain the fc2 accumulator
constexpr int MMA_WARP = LN_WARP0 + LN_WARPS; // warp 20
constexpr int THREADS = (MMA_WARP + 1) * 32;  // 672

constexpr int OFF_W1M = 0;
constexpr int OFF_W1T = 49152;
constexpr int OFF_W2 = OFF_W1T + 24576;
constexpr int OFF_A = OFF_W2 + 73728;          // two [128 x 128 B] sub-tiles: K 0..63, K 64..95 (+ unused half)
constexpr int OFF_HID = OFF_A + 32768;         // two [128 x 128 B] chunks
constexpr int OFF_IO = OFF_HID + 2 * 16384;
constexpr int SMEM_BYTES = OFF_IO + 16384;     // 229376
static_assert(OFF_W2 + 73728 == MLP96_IMG_BYTES, "image layout");
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "shared memory budget");

constexpr int TM_D0 = 0;      // fc1 accumulator, hidden columns 0..191
constexpr int TM_D1 = 192;    // hidden columns 192..383
constexpr int TM_OUT = 384;   // fc2 accumulator, 96 columns
constexpr int TM_COLS = 512;

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global tile store that ADDS into the destination (fp32 add performed by the memory system)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

struct Mlp96Params {
  const float* x;
  int M, num_tiles;
  const float* gamma;
  const float* beta;
  float eps;
  const __nv_bfloat16* img;
  const float* b1;
  const float* b2;
};

__global__ void __launch_bounds__(THREADS, 1)
swin_mlp96_fused_kernel(const __grid_constant__ CUtensorMap tmX, const Mlp96Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_bar;
  __shared__ uint64_t a_full, a_empty;
  __shared__ uint64_t d_full[2], d_empty[2];
  __shared__ uint64_t hid_full[2], hid_empty[2];
  __shared__ uint64_t out_full, out_empty;
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint32_t hid_written[2];   // completed writes of each hidden buffer (monotonic; see the GELU groups)
  __shared__ __align__(16) float sb1[MLP96_H];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    mbar_init(&a_full, LN_WARPS * 32);
    mbar_init(&a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d_full[s], 1);
      mbar_init(&d_empty[s], GELU_WARPS * 32);
      mbar_init(&hid_full[s], 128);
      mbar_init(&hid_empty[s], 1);
    }
    mbar_init(&out_full, 1);
    mbar_init(&out_empty, 128);
    hid_written[0] = hid_written[1] = 0;
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(&tmem_base_slot, TM_COLS);
    tmem_relinquish();
  }
  if (warp == LN_WARP0 && lane == 0) tma_prefetch_desc(&tmX);
  for (int idx = threadIdx.x; idx < MLP96_H; idx += THREADS) sb1[idx] = p.b1[idx];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);

  if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ weight fetch + MMA issue (one thread)
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, MLP96_IMG_BYTES);
      const uint8_t* img = reinterpret_cast<const uint8_t*>(p.img);
      bulk_copy_g2s(smem + OFF_W1M, img + OFF_W1M, 49152, &w_bar);
      bulk_copy_g2s(smem + OFF_W1T, img + OFF_W1T, 24576, &w_bar);
      bulk_copy_g2s(smem + OFF_W2, img + OFF_W2, 73728, &w_bar);

      const uint32_t idesc1 = make_idesc_bf16(TILE_M, 192);
      const uint32_t idesc2 = make_idesc_bf16(TILE_M, MLP96_C);
      const uint64_t a0 = make_smem_desc_sw128(smem_base + OFF_A);
      const uint64_t a1 = make_smem_desc_sw128(smem_base + OFF_A + 16384);
      auto issue_fc1 = [&](int h) {
        const uint32_t d = tmem_base + static_cast<uint32_t>(h ? TM_D1 : TM_D0);
        const uint64_t b0 = make_smem_desc_sw128(smem_base + OFF_W1M + h * 24576);
        // tail tile: row r holds W1[r, 64:96] in bytes 0..63 and W1[r + 192, 64:96] in bytes 64..127
        const uint64_t b1 = make_smem_desc_sw128(smem_base + OFF_W1T) + static_cast<uint64_t>(4 * h);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, a0 + 2 * k, b0 + 2 * k, idesc1, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_bf16(d, a1 + 2 * k, b1 + 2 * k, idesc1, 1u);
      };
      auto issue_fc2 = [&](uint32_t n, int c) {
        mbar_wait(&hid_full[n & 1], (n >> 1) & 1u, 16);
        tc_fence_after();
        const uint64_t a = make_smem_desc_sw128(smem_base + OFF_HID + (n & 1) * 16384);
        const uint64_t b = make_smem_desc_sw128(smem_base + OFF_W2 + c * 12288);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + TM_OUT, a + 2 * k, b + 2 * k, idesc2, (c | k) != 0 ? 1u : 0u);
        umma_commit(&hid_empty[n & 1]);
      };

      mbar_wait(&w_bar, 0, 17);
      mbar_wait(&a_full, 0, 18);
      tc_fence_after();
      issue_fc1(0);
      umma_commit(&d_full[0]);
      issue_fc1(1);
      umma_commit(&d_full[1]);
      umma_commit(&a_empty);
      uint32_t n = 0;
      for (int i = 0; i < n_local; ++i) {
        const bool has_next = i + 1 < n_local;
        if (i > 0) {
          mbar_wait(&out_empty, (i - 1) & 1u, 19);   // the output warps have drained the previous fc2 accumulator
          tc_fence_after();
        }
        for (int c = 0; c < 3; ++c, ++n) issue_fc2(n, c);
        if (has_next) {
          mbar_wait(&a_full, (i + 1) & 1u, 18);
          mbar_wait(&d_empty[0], i & 1u, 20);
          tc_fence_after();
          issue_fc1(0);
          umma_commit(&d_full[0]);
        }
        for (int c = 3; c < 6; ++c, ++n) issue_fc2(n, c);
        umma_commit(&out_full);
        if (has_next) {
          mbar_wait(&d_empty[1], i & 1u, 21);
          tc_fence_after();
          issue_fc1(1);
          umma_commit(&d_full[1]);
          umma_commit(&a_empty);
        }
      }
    }
  } else if (warp >= LN_WARP0) {
    // ------------------------------------------------------------------ LayerNorm -> bf16 A tile  (+ output drain)
    const int t = threadIdx.x - LN_WARP0 * 32;   // 0..255
    const int l8 = t & 7;                        // lane within the 8-lane row team
    const int tt = t >> 3;                       // team 0..31 -> rows rg + 32q; the four teams of a warp sit on rows
    const int rg = 8 * (tt >> 3) + ((tt & 1) << 2) + ((tt >> 1) & 3);   // whose swizzle phases spread over all banks
    auto layer_norm_tile = [&](int i) {
      const int m0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TILE_M;
      float4 xv[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = m0 + rg + 32 * q;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (row < p.M)
            xv[q][j] = __ldcg(reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * MLP96_C + 4 * l8 + 32 * j));
          else
            xv[q][j] = make_float4(0.f, 0.f, 0.f, 0.f);
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
        const float mean = s * (1.0f / MLP96_C);
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
        rstd[q] = rsqrtf(v * (1.0f / MLP96_C) + p.eps);
      }
      mbar_wait_relaxed(&a_empty, (i & 1u) ^ 1u, 22);    // fc1 of the previous tile has consumed the A tile
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = 4 * l8 + 32 * j;         // 0..95
        const int kc = col & 63;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
        const float4 be4 = __ldg(reinterpret_cast<const float4*>(p.beta + col));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = rg + 32 * q;
          const float o0 = fmaf(xv[q][j].x * rstd[q], g4.x, be4.x);
          const float o1 = fmaf(xv[q][j].y * rstd[q], g4.y, be4.y);
          const float o2 = fmaf(xv[q][j].z * rstd[q], g4.z, be4.z);
          const float o3 = fmaf(xv[q][j].w * rstd[q], g4.w, be4.w);
          uint8_t* dst = smem + OFF_A + (col >> 6) * 16384 + r * 128 + ((((kc >> 3) ^ (r & 7))) << 4) + (kc & 7) * 2;
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full);
    };
    // fc2 accumulator -> + bias -> 32-column slabs -> TMA reduce-add into x (warps LN_WARP0 .. LN_WARP0 + 3)
    const bool out_warp = warp < LN_WARP0 + 4;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int sw = row & 7;
    const bool elected = threadIdx.x == LN_WARP0 * 32;
    uint8_t* my_out = smem + OFF_IO + row * 128;
    auto drain_tile = [&](int i) {
      const int m0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TILE_M;
      mbar_wait_relaxed(&out_full, i & 1u, 23, 500);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + TM_OUT;
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(32 * s), v);
        tmem_ld_wait();
        if (s == 2) {
          tc_fence_before();
          mbar_arrive(&out_empty);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + 32 * s) + j);
          v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b4.x);
          v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b4.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b4.z);
          v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b4.w);
        }
        if (elected) tma_store_wait_read<0>();   // the previous reduce has read the slab
        named_bar_sync(1, 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(my_out + ((j ^ sw) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (elected) {
          tma_reduce_add_2d(&tmX, smem + OFF_IO, 32 * s, m0);   // rows >= M are clipped by the tensor map
          tma_store_commit();
        }
      }
    };
    layer_norm_tile(0);
    for (int i = 0; i < n_local; ++i) {
      if (i + 1 < n_local) layer_norm_tile(i + 1);
      if (out_warp) drain_tile(i);
    }
    if (elected) tma_store_wait_all();
  } else {
    // ------------------------------------------------------------------ GELU groups: fc1 accumulator -> hidden chunk
    const int group = warp >> 2;        // 0..2
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int sw = row & 7;
    for (int i = 0; i < n_local; ++i) {
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int c = group + 3 * hh;                       // chunk of this tile: hidden columns [64c, 64c + 64)
        const uint32_t n = 6u * static_cast<uint32_t>(i) + static_cast<uint32_t>(c);
        const int buf = n & 1;
        uint8_t* my_hid = smem + OFF_HID + buf * 16384 + row * 128;
        const float* bias = sb1 + 64 * c;
        mbar_wait(&d_full[hh], i & 1u, 24);
        tc_fence_after();
        const uint32_t t_col = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>((hh ? TM_D1 : TM_D0) + group * 64);
        uint32_t v[32];
        uint32_t pk[16];
        tmem_ld_32x32b_x32(t_col, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * j);
          pk[2 * j] = pack_bf16(gelu_erf(__uint_as_float(v[4 * j]) + b4.x), gelu_erf(__uint_as_float(v[4 * j + 1]) + b4.y));
          pk[2 * j + 1] = pack_bf16(gelu_erf(__uint_as_float(v[4 * j + 2]) + b4.z), gelu_erf(__uint_as_float(v[4 * j + 3]) + b4.w));
        }
        tmem_ld_32x32b_x32(t_col + 32u, v);
        // Use k of this buffer may be written once fc2 has consumed use k-1. The three groups take turns on the two
        // buffers, so a parity wait alone could be satisfied by a phase two uses back: first make sure use k-1 has
        // been WRITTEN (monotonic counter), which pins the barrier to phase k-1 or k, then wait for phase k-1.
        const uint32_t use = n >> 1;
        if (use > 0) {
          uint32_t spins = 0;
          while (*reinterpret_cast<volatile uint32_t*>(&hid_written[buf]) < use) {
            if (((++spins) & 0x3FFF) == 0 &&
                (*reinterpret_cast<volatile unsigned int*>(&g_mbar_timeout) != 0 || spins > (1u << 24))) {
              atomicCAS(&g_mbar_timeout, 0u, 0x80000000u | (26u << 24) | ((blockIdx.x & 0xFFF) << 12) | (threadIdx.x & 0xFFF));
              break;
            }
          }
        }
        mbar_wait(&hid_empty[buf], (use & 1u) ^ 1u, 25);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(my_hid + ((j ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&d_empty[hh]);                          // this thread no longer needs the accumulator half
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + 32 + 4 * j);
          pk[2 * j] = pack_bf16(gelu_erf(__uint_as_float(v[4 * j]) + b4.x), gelu_erf(__uint_as_float(v[4 * j + 1]) + b4.y));
          pk[2 * j + 1] = pack_bf16(gelu_erf(__uint_as_float(v[4 * j + 2]) + b4.z), gelu_erf(__uint_as_float(v[4 * j + 3]) + b4.w));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(my_hid + (((4 + j) ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        mbar_arrive(&hid_full[buf]);
        named_bar_sync(2 + group, 128);                     // every thread of the group has written this use
        if ((threadIdx.x & 127) == 0) *reinterpret_cast<volatile uint32_t*>(&hid_written[buf]) = use + 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

inline size_t sw128_off(int r, int k) {   // byte offset of bf16 element (row r, column k < 64) in a SWIZZLE_128B tile
  return static_cast<size_t>(r) * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2;
}

}  // namespace

void mlp96_pack_weights(const float* fc1_w, const float* fc2_w, __nv_bfloat16* img_host) {
  uint8_t* img = reinterpret_cast<uint8_t*>(img_host);
  auto put = [&](size_t byte_off, float v) { *reinterpret_cast<__nv_bfloat16*>(img + byte_off) = __float2bfloat16(v); };
  for (size_t i = 0; i < MLP96_IMG_BYTES / 2; ++i) img_host[i] = __float2bfloat16(0.f);
  for (int n = 0; n < MLP96_H; ++n) {
    for (int k = 0; k < 64; ++k) put(OFF_W1M + sw128_off(n, k), fc1_w[static_cast<size_t>(n) * MLP96_C + k]);
    const int r = n % 192, half = n / 192;
    for (int kk = 0; kk < 32; ++kk)
      put(OFF_W1T + sw128_off(r, half * 32 + kk), fc1_w[static_cast<size_t>(n) * MLP96_C + 64 + kk]);
  }
  for (int c = 0; c < 6; ++c)
    for (int n = 0; n < MLP96_C; ++n)
      for (int k = 0; k < 64; ++k)
        put(OFF_W2 + static_cast<size_t>(c) * 12288 + sw128_off(n, k), fc2_w[static_cast<size_t>(n) * MLP96_H + 64 * c + k]);
}

FMMT_DEFINE_WATCHDOG_ADDR(watchdog_addr_mlp96)

unsigned int read_mlp_timeout(bool reset) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_mbar_timeout, sizeof(v));
  if (reset && v != 0) {
    unsigned int z = 0;
    cudaMemcpyToSymbol(g_mbar_timeout, &z, sizeof(z));
  }
  return v;
}

cudaError_t launch_mlp96(const Mlp96Args& a, cudaStream_t stream) {
  if (a.M <= 0 || !a.x || !a.gamma || !a.beta || !a.img || !a.b1 || !a.b2) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.img) & 15) ||
      (reinterpret_cast<uintptr_t>(a.gamma) & 15) || (reinterpret_cast<uintptr_t>(a.beta) & 15) ||
      (reinterpret_cast<uintptr_t>(a.b1) & 15) || (reinterpret_cast<uintptr_t>(a.b2) & 15))
    return cudaErrorInvalidValue;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 148;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(swin_mlp96_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  });
  if (attr_err != cudaSuccess) return attr_err;
  CUtensorMap tmX;
  if (!make_tmap_2d(&tmX, a.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, MLP96_C, MLP96_C, 32, TILE_M))
    return cudaErrorInvalidValue;
  Mlp96Params p{a.x, a.M, (a.M + TILE_M - 1) / TILE_M, a.gamma, a.beta, a.eps, a.img, a.b1, a.b2};
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  swin_mlp96_fused_kernel<<<grid, THREADS, SMEM_BYTES + 1024, stream>>>(tmX, p);
  return cudaGetLastError();
}

}  // namespace fmmt
