// Fused Swin MLP half-block for C = 192 / 384 (Swin-tiny stages 2 and 3; Swin_Transformer.py:24-30 Mlp.forward inside
// :264-268 "x = x + drop_path(mlp(norm2(x)))"):
//     x <- x + fc2( GELU_erf( fc1( LayerNorm(x) ) ) )
// as ONE persistent kernel. At these widths the weights (0.6 / 2.4 MB) do not fit in shared memory, so they are STREAMED
// from L2 once per 128-row tile, one 64-wide hidden chunk at a time, while the tile's normalised activations stay in
// shared memory and its output accumulates in TMEM:
//   * 8 LayerNorm warps read the 128 x C tile of x (coalesced, 16 lanes per row) and write the bf16 A operand as C/64
//     K-major SWIZZLE_128B sub-tiles;
//   * two producer threads stream, per hidden chunk j, W1[64j:64j+64, :] as 3-D TMA boxes of three k-blocks (24 KB pieces,
//     two piece slots) and W2[:, 64j:64j+64] as ONE 3-D box (64 x 192 x C/192) - big boxes because the TMA unit serves
//     one box at a time at max(~600 cycles, bytes / 54 B/clk) (feed probe); separate threads so that neither
//     stream waits behind the other's free-slot wait;
//   * one thread issues tcgen05.mma: fc1_j (N = 64) into one of two 64-column TMEM buffers, fc2_j (K = 64, N = C as
//     192-column MMAs) accumulating into C TMEM columns over all chunks;
//   * 16 GELU warps in two groups (group = chunk parity, each owning one TMEM buffer and one hidden buffer) turn the
//     fc1 accumulator into the bf16 hidden chunk that is fc2_j's A operand: tcgen05.ld -> bias -> erf-GELU -> smem;
//   * after the last chunk three 128-thread streams drain the fc2 accumulator: + bias, 32-column slabs staged in the (now
//     idle) W2 slot and handed to TMA reduce-add (cp.reduce.async.bulk.tensor .add), which applies the residual in memory.
// HBM traffic per token: 8C bytes (x read, x updated) instead of 32C for LN + fc1 + fc2 as separate kernels; the 4C-wide
// hidden activation never leaves the SM. Shared memory (C = 384): A 96 KB | hidden 2 x 16 KB | W1 pieces 2 x 24 KB | W2 48 KB.
#include "mlp_stream.cuh"

#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace fmmt {

namespace {

constexpr int TILE_M = 128;
constexpr int GELU_WARPS = 16;                 // warps 0..15: group = warp / 8, quarter = warp % 4, column half = (warp / 4) % 2
constexpr int LN_WARP0 = GELU_WARPS;           // warps 16..23
constexpr int LN_WARPS = 8;
constexpr int MMA_WARP = LN_WARP0 + LN_WARPS;  // warp 24
constexpr int PROD1_WARP = MMA_WARP + 1;       // warp 25: fc1 weight stream
constexpr int PROD2_WARP = MMA_WARP + 2;       // warp 26: fc2 weight stream
constexpr int PROD3_WARP = MMA_WARP + 3;       // warp 27: fc1 weight stream, odd pieces (one TMA stream per piece slot:
                                               //   a single thread's boxes are served one at a time, fmmt_debug_feed2)
constexpr int THREADS = (PROD3_WARP + 1) * 32; // 896
constexpr int DRAIN_STREAMS = 3;               // GELU warps 0..11, four warps (128 rows) per stream

template <int C>
struct Cfg {
  static_assert(C == 192 || C == 384, "streamed fused MLP: C = 192 or 384");
  static constexpr int H = 4 * C;
  static constexpr int NKB = C / 64;            // k-blocks of fc1 (A sub-tiles)
  static constexpr int NH = C / 192;            // 192-column halves of the fc2 output
  static constexpr int CHUNKS = H / 64;
  static constexpr int A_BYTES = NKB * 16384;
  static constexpr int PIECES = NKB / 3;        // fc1 weights of a chunk arrive as pieces of 3 k-blocks (64 rows x 192 cols)
  static constexpr int W1_PIECE = 3 * 8192;     // 24 KB; two piece slots: the next piece loads while this one is multiplied
  static constexpr int W2_BYTES = NH * 24576;   // C rows x 64 bf16 (one box)
  static constexpr int R2 = C == 192 ? 2 : 1;   // fc2 weight slots (C = 384: shared memory is full with one)
  static constexpr int OFF_A = 0;
  static constexpr int OFF_HID = A_BYTES;
  static constexpr int OFF_W1 = OFF_HID + 2 * 16384;
  static constexpr int OFF_W2 = OFF_W1 + 2 * W1_PIECE;
  static constexpr int SMEM = OFF_W2 + (R2 * W2_BYTES > DRAIN_STREAMS * 16384 ? R2 * W2_BYTES : DRAIN_STREAMS * 16384);
  static constexpr int TM_OUT = 0;              // fc2 accumulator: C columns
  static constexpr int TM_D1 = 384;             // fc1 accumulators: 2 x 64 columns
  static constexpr int SLABS = C / 32;          // 32-column fp32 output slabs
  static_assert(SMEM + 1024 <= 226 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// TMEM -> registers: this warp's 32 lanes x 32 consecutive columns, no wait (pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32b_x32(taddr, v); }

struct StreamParams {
  const float* x;
  int M, num_tiles;
  const float* gamma;
  const float* beta;
  float eps;
  const float* b1;
  const float* b2;
  int copies;
  long long* trace;   // optional [CHUNKS][8] clock64 stamps of CTA 0, tile 0 (debug / DESIGN.md timeline)
};

template <int C>
__global__ void __launch_bounds__(THREADS, 1)
swin_mlp_stream_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                       const __grid_constant__ CUtensorMap tmX, const StreamParams p) {
  using K = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full, a_empty;
  __shared__ uint64_t w1_full[2], w1_empty[2], w2_full[2], w2_empty[2];
  __shared__ uint64_t d1_full[2], d1_empty[2];
  __shared__ uint64_t hid_full[2], hid_empty[2];
  __shared__ uint64_t out_full, out_empty, drain_done;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    mbar_init(&a_full, LN_WARPS * 32);
    mbar_init(&a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&w1_full[s], 1);
      mbar_init(&w1_empty[s], 1);
      mbar_init(&w2_full[s], 1);
      mbar_init(&w2_empty[s], 1);
      mbar_init(&d1_full[s], 1);
      mbar_init(&d1_empty[s], 256);
      mbar_init(&hid_full[s], 256);
      mbar_init(&hid_empty[s], 1);
    }
    mbar_init(&out_full, 1);
    mbar_init(&out_empty, DRAIN_STREAMS * 128);
    mbar_init(&drain_done, DRAIN_STREAMS);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  if (warp == PROD1_WARP && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  const int copy = static_cast<int>(blockIdx.x) % p.copies;

  if (warp == PROD1_WARP || warp == PROD3_WARP) {
    // ------------------------------------------------------------------ fc1 weight streams (one thread per piece slot)
    if (lane == 0) {
      const uint32_t mine = warp == PROD1_WARP ? 0u : 1u;
      uint32_t u = 0;   // running piece number of this CTA: slot u & 1, use u >> 1
      for (int i = 0; i < n_local; ++i)
        for (int j = 0; j < K::CHUNKS; ++j)
          for (int pc = 0; pc < K::PIECES; ++pc, ++u) {
            const uint32_t sl = u & 1u;
            if (sl != mine) continue;
            mbar_wait(&w1_empty[sl], ((u >> 1) & 1u) ^ 1u, 40);      // the MMAs on this slot's previous piece are done
            if (p.trace != nullptr && blockIdx.x == 0 && i == 0 && pc == 0) p.trace[j * 8 + 5] = clock64();   // piece 0 requested
            mbar_arrive_expect_tx(&w1_full[sl], K::W1_PIECE);
            tma_load_3d(smem + K::OFF_W1 + sl * K::W1_PIECE, &tmW1, &w1_full[sl], 0, copy * K::H + 64 * j, 3 * pc);
          }
    }
  } else if (warp == PROD2_WARP) {
    // ------------------------------------------------------------------ fc2 weight stream (one thread)
    if (lane == 0) {
      uint32_t n = 0;   // running chunk number of this CTA: slot n % R2, use n / R2
      for (int i = 0; i < n_local; ++i)
        for (int j = 0; j < K::CHUNKS; ++j, ++n) {
          const uint32_t sl = n % K::R2, use = n / K::R2;
          mbar_wait(&w2_empty[sl], (use & 1u) ^ 1u, 41);            // fc2 of this slot's previous chunk is done
          if (i > 0 && j < K::R2) mbar_wait(&drain_done, (i - 1) & 1u, 42);   // ... and the drain staged here is out
          if (p.trace != nullptr && blockIdx.x == 0 && i == 0) p.trace[j * 8 + 6] = clock64();   // fc2 weights requested
          mbar_arrive_expect_tx(&w2_full[sl], K::W2_BYTES);
          tma_load_3d(smem + K::OFF_W2 + sl * K::W2_BYTES, &tmW2, &w2_full[sl], 64 * j, 0, copy * K::NH);
        }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issue (one thread)
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(TILE_M, 64);
      const uint32_t idesc2 = make_idesc_bf16(TILE_M, 192);
      uint32_t n = 0;
      auto issue_fc1 = [&](uint32_t nn) {       // chunk nn of this CTA -> TMEM buffer nn & 1
        const uint32_t b = nn & 1u, use = nn >> 1;
        mbar_wait(&d1_empty[b], (use & 1u) ^ 1u, 44);              // the GELU group has drained this buffer's last use
        const uint32_t d = tmem_base + K::TM_D1 + 64u * b;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && nn < static_cast<uint32_t>(K::CHUNKS);
        if (tr) p.trace[nn * 8 + 0] = clock64();                   // d1 buffer free
#pragma unroll
        for (int pc = 0; pc < K::PIECES; ++pc) {
          const uint32_t u = nn * K::PIECES + pc, sl = u & 1u;
          mbar_wait(&w1_full[sl], (u >> 1) & 1u, 43);
          if (tr) p.trace[nn * 8 + 1 + pc] = clock64();            // fc1 weight piece landed
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const uint64_t a = make_smem_desc_sw128(smem_base + K::OFF_A + (3 * pc + kk) * 16384);
            const uint64_t w = make_smem_desc_sw128(smem_base + K::OFF_W1 + sl * K::W1_PIECE + kk * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a + 2 * k, w + 2 * k, idesc1, (pc | kk | k) != 0 ? 1u : 0u);
          }
          umma_commit(&w1_empty[sl]);
        }
        umma_commit(&d1_full[b]);
      };
      for (int i = 0; i < n_local; ++i) {
        mbar_wait(&a_full, i & 1u, 45);
        tc_fence_after();
        issue_fc1(n);
        for (int j = 0; j < K::CHUNKS; ++j, ++n) {
          if (j + 1 < K::CHUNKS) issue_fc1(n + 1);
          else umma_commit(&a_empty);            // every fc1 of this tile has been issued: A may be rewritten after them
          const uint32_t b = n & 1u, use = n >> 1;
          const uint32_t sl2 = n % K::R2;
          const bool tr = p.trace != nullptr && blockIdx.x == 0 && i == 0;
          mbar_wait(&w2_full[sl2], (n / K::R2) & 1u, 46);
          if (tr) p.trace[j * 8 + 3] = clock64();                  // fc2 weights landed
          mbar_wait(&hid_full[b], use & 1u, 47);
          if (tr) p.trace[j * 8 + 4] = clock64();                  // hidden chunk written
          if (j == 0 && i > 0) mbar_wait(&out_empty, (i - 1) & 1u, 48);   // previous tile's accumulator has been drained
          tc_fence_after();
          const uint64_t a = make_smem_desc_sw128(smem_base + K::OFF_HID + b * 16384);
#pragma unroll
          for (int h = 0; h < K::NH; ++h) {
            const uint64_t w = make_smem_desc_sw128(smem_base + K::OFF_W2 + sl2 * K::W2_BYTES + h * 24576);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + K::TM_OUT + 192u * h, a + 2 * k, w + 2 * k, idesc2, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&w2_empty[sl2]);
          umma_commit(&hid_empty[b]);
        }
        umma_commit(&out_full);
      }
    }
  } else if (warp >= LN_WARP0) {
    // ------------------------------------------------------------------ LayerNorm -> bf16 A tile
    const int t = threadIdx.x - LN_WARP0 * 32;   // 0..255
    const int l16 = t & 15;                      // lane within the 16-lane row team
    const int team = t >> 4;                     // 0..15: rows team + 16 * pass
    constexpr int Q = K::NKB;                    // float4 per lane per row (one per 64-column k-block)
    constexpr int BATCH = 12 / Q;                // rows in flight per thread: 48 data registers
    for (int i = 0; i < n_local; ++i) {
      const int tile = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int m0 = tile * TILE_M;
      bool waited = false;
#pragma unroll 1
      for (int pass0 = 0; pass0 < 8; pass0 += BATCH) {
        float4 xv[BATCH][Q];
#pragma unroll
        for (int bq = 0; bq < BATCH; ++bq) {
          const int row = m0 + team + 16 * (pass0 + bq);
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            if (row < p.M)
              xv[bq][q] = __ldcg(reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * C + 4 * l16 + 64 * q));
            else
              xv[bq][q] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float rstd[BATCH];
#pragma unroll
        for (int bq = 0; bq < BATCH; ++bq) {
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < Q; ++q) s += (xv[bq][q].x + xv[bq][q].y) + (xv[bq][q].z + xv[bq][q].w);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          const float mean = s * (1.0f / C);
          float v = 0.f;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            xv[bq][q].x -= mean; xv[bq][q].y -= mean; xv[bq][q].z -= mean; xv[bq][q].w -= mean;
            v = fmaf(xv[bq][q].x, xv[bq][q].x, v); v = fmaf(xv[bq][q].y, xv[bq][q].y, v);
            v = fmaf(xv[bq][q].z, xv[bq][q].z, v); v = fmaf(xv[bq][q].w, xv[bq][q].w, v);
          }
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          rstd[bq] = rsqrtf(v * (1.0f / C) + p.eps);
        }
        if (!waited) {
          mbar_wait_relaxed(&a_empty, (i & 1u) ^ 1u, 49, 500);   // every fc1 of the previous tile has consumed the A tile
          waited = true;
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int col = 4 * l16 + 64 * q;
          const int kc = 4 * l16;
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
          const float4 be4 = __ldg(reinterpret_cast<const float4*>(p.beta + col));
#pragma unroll
          for (int bq = 0; bq < BATCH; ++bq) {
            const int r = team + 16 * (pass0 + bq);
            const float o0 = fmaf(xv[bq][q].x * rstd[bq], g4.x, be4.x);
            const float o1 = fmaf(xv[bq][q].y * rstd[bq], g4.y, be4.y);
            const float o2 = fmaf(xv[bq][q].z * rstd[bq], g4.z, be4.z);
            const float o3 = fmaf(xv[bq][q].w * rstd[bq], g4.w, be4.w);
            uint8_t* dst = smem + K::OFF_A + q * 16384 + r * 128 + ((((kc >> 3) ^ (r & 7))) << 4) + (kc & 7) * 2;
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full);
      // pull the next tile's rows towards L2 while this tile computes (its LayerNorm is on the critical path)
      if (i + 1 < n_local) {
        const int nm0 = (tile + static_cast<int>(gridDim.x)) * TILE_M;
        for (int idx = t; idx < TILE_M * (C * 4 / 128); idx += LN_WARPS * 32) {
          const int row = nm0 + idx / (C * 4 / 128);
          if (row < p.M) prefetch_l2(p.x + static_cast<size_t>(row) * C + (idx % (C * 4 / 128)) * 32);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ GELU groups (+ output drain on warps 0..11)
    const int group = warp >> 3;          // chunk parity handled == TMEM buffer == hidden buffer
    const int quarter = warp & 3;
    const int half = (warp >> 2) & 1;     // which 32 of the chunk's 64 columns
    const int row = quarter * 32 + lane;
    const int sw = row & 7;
    uint8_t* my_hid = smem + K::OFF_HID + group * 16384 + row * 128;
    // drain: stream = warp / 4 (0..2), slabs stream, stream + 3, ...
    const int stream = warp >> 2;
    const bool drains = stream < DRAIN_STREAMS;
    const bool elected = drains && (threadIdx.x & 127) == 0;
    uint8_t* stage = smem + K::OFF_W2 + stream * 16384;
    uint8_t* my_out = stage + row * 128;
    for (int i = 0; i < n_local; ++i) {
      const int m0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TILE_M;
#pragma unroll 1
      for (int jj = 0; jj < K::CHUNKS / 2; ++jj) {
        const int j = 2 * jj + group;
        const uint32_t use = static_cast<uint32_t>(i) * (K::CHUNKS / 2) + static_cast<uint32_t>(jj);
        mbar_wait(&d1_full[group], use & 1u, 50);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + K::TM_D1 + 64u * group + 32u * half, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&d1_empty[group]);
        const float* bias = p.b1 + 64 * j + 32 * half;
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + q);
          pk[2 * q] = pack_bf16(gelu_erf(__uint_as_float(v[4 * q]) + b4.x), gelu_erf(__uint_as_float(v[4 * q + 1]) + b4.y));
          pk[2 * q + 1] = pack_bf16(gelu_erf(__uint_as_float(v[4 * q + 2]) + b4.z), gelu_erf(__uint_as_float(v[4 * q + 3]) + b4.w));
        }
        mbar_wait(&hid_empty[group], (use & 1u) ^ 1u, 51);   // fc2 has consumed this buffer's previous chunk
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(my_hid + (((4 * half + q) ^ sw) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        fence_proxy_async_smem();
        mbar_arrive(&hid_full[group]);
        if (p.trace != nullptr && blockIdx.x == 0 && i == 0 && (threadIdx.x & 255) == 0) p.trace[j * 8 + 7] = clock64();
      }
      if (drains) {
        // fc2 accumulator -> + bias -> 32-column slabs staged in the W2 slot -> TMA reduce-add into x
        mbar_wait(&out_full, i & 1u, 52);
        tc_fence_after();
#pragma unroll 1
        for (int s = stream; s < K::SLABS; s += DRAIN_STREAMS) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + K::TM_OUT + 32u * s, v);
          tmem_ld_wait();
          if (s + DRAIN_STREAMS >= K::SLABS) {
            tc_fence_before();
            mbar_arrive(&out_empty);             // my last read of this tile's accumulator
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + 32 * s) + q);
            v[4 * q + 0] = __float_as_uint(__uint_as_float(v[4 * q + 0]) + b4.x);
            v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + b4.y);
            v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + b4.z);
            v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + b4.w);
          }
          if (elected) tma_store_wait_read<0>();   // the previous reduce of this stream has read the staging slab
          named_bar_sync(1 + stream, 128);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(my_out + ((q ^ sw) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          fence_proxy_async_smem();
          named_bar_sync(1 + stream, 128);
          if (elected) {
            tma_reduce_add_2d(&tmX, stage, 32 * s, m0);   // rows >= M are clipped by the tensor map
            tma_store_commit();
          }
        }
        if (elected) {
          tma_store_wait_read<0>();              // the W2 slot may be refilled
          mbar_arrive(&drain_done);
        }
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

// 3-D view of fc2.weight [C rows, H cols] (row-major): (column, row within a 192-row half, half)
bool make_tmap_w2(CUtensorMap* tm, const void* base, int C, int H, int ld, int copies) {
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(H), 192, static_cast<cuuint64_t>(C / 192) * copies};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * 192};
  cuuint32_t box[3] = {64, 192, static_cast<cuuint32_t>(C / 192)};
  cuuint32_t estr[3] = {1, 1, 1};
  return reinterpret_cast<PFN>(ptr)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C>
cudaError_t launch_c(const MlpStreamArgs& a, cudaStream_t stream) {
  using K = Cfg<C>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 148;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(swin_mlp_stream_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM + 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  });
  if (attr_err != cudaSuccess) return attr_err;
  CUtensorMap tmW1, tmW2, tmX;
  if (!make_tmap_kblocks_2d(&tmW1, a.w1, static_cast<long long>(K::H) * a.copies, C, a.ldw1, 64, 3)) return cudaErrorInvalidValue;
  if (!make_tmap_w2(&tmW2, a.w2, C, K::H, a.ldw2, a.copies)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&tmX, a.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, C, C, 32, TILE_M)) return cudaErrorInvalidValue;
  StreamParams p{a.x, a.M, (a.M + TILE_M - 1) / TILE_M, a.gamma, a.beta, a.eps, a.b1, a.b2, a.copies, a.trace};
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  swin_mlp_stream_kernel<C><<<grid, THREADS, K::SMEM + 1024, stream>>>(tmW1, tmW2, tmX, p);
  return cudaGetLastError();
}

}  // namespace

FMMT_DEFINE_WATCHDOG_ADDR(watchdog_addr_mlp_stream)

unsigned int read_mlp_stream_timeout(bool reset) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_mbar_timeout, sizeof(v));
  if (reset && v != 0) {
    unsigned int z = 0;
    cudaMemcpyToSymbol(g_mbar_timeout, &z, sizeof(z));
  }
  return v;
}

cudaError_t launch_mlp_stream(const MlpStreamArgs& a, cudaStream_t stream) {
  if (a.copies < 1 || a.copies > 64) return cudaErrorInvalidValue;
  if (a.M <= 0 || !a.x || !a.gamma || !a.beta || !a.w1 || !a.w2 || !a.b1 || !a.b2) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.w1) & 15) ||
      (reinterpret_cast<uintptr_t>(a.w2) & 15) || (reinterpret_cast<uintptr_t>(a.gamma) & 15) ||
      (reinterpret_cast<uintptr_t>(a.beta) & 15) || (reinterpret_cast<uintptr_t>(a.b1) & 15) ||
      (reinterpret_cast<uintptr_t>(a.b2) & 15) || (a.ldw1 % 8) != 0 || (a.ldw2 % 8) != 0)
    return cudaErrorInvalidValue;
  if (a.C == 192) return launch_c<192>(a, stream);
  if (a.C == 384) return launch_c<384>(a, stream);
  return cudaErrorInvalidValue;
}

}  // namespace fmmt
