// LayerNorm + window gather fused into the qkv Linear of a Swin block, C = 192 / 384 (Swin-tiny stages 2 and 3):
//     out[r, :] = LayerNorm1(x[g(r), :]) @ W^T + b            (bf16 out),      x_raw[r, :] = x[g(r), :]   (fp32, optional)
// (Swin_Transformer.py:238-247: norm1 -> roll -> window_partition, then WindowAttention's qkv Linear :119; g = the composed
// roll + window_partition gather of this block). Un-fused this was a LayerNorm kernel writing a bf16 copy of every row plus a
// GEMM reading it back; here the normalised 128-row tile never leaves the SM:
//   * 8 LayerNorm warps gather the tile's rows of x (coalesced, 16 lanes per row), write them out as the block's new
//     residual stream (window order), normalise and write the bf16 A operand as C/64 K-major SWIZZLE_128B sub-tiles;
//   * 3 / 4 producer threads (one per ring slot) stream the weights of a 256-column output chunk one k-block at a time
//     (256 rows x 64 columns = 32 KB boxes);
//   * one thread issues tcgen05.mma (M = 128, N = 256, K = 16 per instruction) into one of two 256-column TMEM buffers. N = 256
//     on purpose: the instructions of one accumulator form a dependent chain, and a chain link costs ~150-200 cycles whatever
//     N is (measured here: 64-column chunks ran at 4400 cycles per 24 instructions); at N = 256 the 128 execution cycles of an
//     instruction cover it;
//   * two drain groups read the accumulator 64 columns at a time, add the bias, pack bf16 into a 128B-swizzled slab and hand
//     it to a TMA store.
// Shared memory: A 96 KB (C = 384: one tile; C = 192: two tiles of 48 KB) | weight ring 3 x 32 KB | output slabs 2 x 16 KB.
#include "ln_qkv.cuh"

#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace fmmt {

namespace {

constexpr int TILE_M = 128;
constexpr int DRAIN_WARPS = 8;                  // warps 0..7: group = warp / 4 = chunk parity, quarter = warp % 4
constexpr int LN_WARP0 = DRAIN_WARPS;           // warps 8..15
constexpr int LN_WARPS = 8;
constexpr int MMA_WARP = LN_WARP0 + LN_WARPS;   // warp 16
constexpr int PROD0_WARP = MMA_WARP + 1;        // warps 17..: one producer thread per weight-piece slot (a thread's TMA boxes are
constexpr int MAX_SLOTS = 4;                    //   served one at a time)
constexpr int THREADS = (PROD0_WARP + MAX_SLOTS) * 32;  // 672
constexpr int BN = 256;                         // output columns per chunk (last chunk: the remainder, a multiple of 64)

template <int C>
struct Cfg {
  static_assert(C == 192 || C == 384, "LN + qkv: C = 192 or 384");
  static constexpr int NKB = C / 64;
  static constexpr int PIECE = BN * 128;        // one k-block of a chunk's weights: 256 rows x 64 bf16 = 32 KB
  static constexpr int NSLOT = 3;                   // k-blocks in flight
  static constexpr int A_BYTES = NKB * 16384;       // one normalised 128-row tile
  static constexpr int NA = C == 384 ? 1 : 2;       // A tiles: at C = 192 two fit, so the LayerNorm of tile i + 1 overlaps the
                                                    // MMAs of tile i; at C = 384 (96 KB per tile) they alternate
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = NA * A_BYTES;
  static constexpr int OFF_OUT = OFF_W + NSLOT * PIECE;
  static constexpr int SMEM = OFF_OUT + 2 * 16384;
  static_assert(SMEM + 1024 <= 227 * 1024, "shared memory budget");
};

struct LnQkvParams {
  const float* x; float* x_raw;
  int M, T, num_tiles, chunks, N;
  const int* gather;
  const float* gamma; const float* beta; float eps;
  const float* bias;
  long long* trace;   // optional [8 tiles][32] clock64 stamps of CTA 0 (debug)
  int dbg;            // debug probes: 1 = skip the x_raw stores, 2 = skip the output stores (wrong results)
};

#define TRACE(slot)                                                                          \
  do {                                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0 && i < 8) p.trace[i * 32 + (slot)] = clock64(); \
  } while (0)

template <int C>
__global__ void __launch_bounds__(THREADS, 1)
ln_qkv_stream_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWrem,
                     const __grid_constant__ CUtensorMap tmOut, const LnQkvParams p) {
  using K = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[2], a_empty[2], w_full[MAX_SLOTS], w_empty[MAX_SLOTS], d_full[2], d_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], LN_WARPS * 32);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < MAX_SLOTS; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d_full[s], 1);
      mbar_init(&d_empty[s], 256);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  if (warp == PROD0_WARP && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmWrem);
    tma_prefetch_desc(&tmOut);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp >= PROD0_WARP) {
    // ------------------------------------------------------------------ weight stream (one thread per ring slot)
    const uint32_t mine = static_cast<uint32_t>(warp - PROD0_WARP);
    if (lane == 0 && mine < static_cast<uint32_t>(K::NSLOT)) {
      uint32_t u = 0;   // running k-block number of this CTA: slot u % NSLOT, use u / NSLOT
      for (int i = 0; i < n_local; ++i)
        for (int j = 0; j < p.chunks; ++j) {
          const int bn = p.N - j * BN < BN ? p.N - j * BN : BN;
          for (int kb = 0; kb < K::NKB; ++kb, ++u) {
            const uint32_t sl = u % K::NSLOT;
            if (sl != mine) continue;
            mbar_wait(&w_empty[sl], ((u / K::NSLOT) & 1u) ^ 1u, 90);
            mbar_arrive_expect_tx(&w_full[sl], static_cast<uint32_t>(bn * 128));
            tma_load_2d(smem + K::OFF_W + sl * K::PIECE, bn == BN ? &tmW : &tmWrem, &w_full[sl], kb * 64, j * BN);
          }
        }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issue (one thread)
    if (lane == 0) {
      uint32_t u = 0;
      uint32_t cnt[2] = {0u, 0u};    // uses of each TMEM buffer so far (chunk j of a tile goes to buffer j & 1)
      for (int i = 0; i < n_local; ++i) {
        TRACE(7);
        const uint32_t ab = K::NA == 2 ? (i & 1u) : 0u;                       // A tile buffer of this tile and the parity
        const uint32_t apar = K::NA == 2 ? ((i >> 1) & 1u) : (i & 1u);        // of its current use
        mbar_wait(&a_full[ab], apar, 91);
        TRACE(8);
        tc_fence_after();
        for (int j = 0; j < p.chunks; ++j) {
          const int bn = p.N - j * BN < BN ? p.N - j * BN : BN;
          const uint32_t idesc = make_idesc_bf16(TILE_M, bn);
          const uint32_t b = j & 1u;
          mbar_wait(&d_empty[b], (cnt[b] & 1u) ^ 1u, 92);            // both drain groups have read this buffer's last use
          ++cnt[b];
          tc_fence_after();
          const uint32_t d = tmem_base + 256u * b;
#pragma unroll
          for (int kb = 0; kb < K::NKB; ++kb, ++u) {
            const uint32_t sl = u % K::NSLOT;
            mbar_wait(&w_full[sl], (u / K::NSLOT) & 1u, 93);
            tc_fence_after();
            const uint64_t a = make_smem_desc_sw128(smem_base + K::OFF_A + ab * K::A_BYTES + kb * 16384);
            const uint64_t w = make_smem_desc_sw128(smem_base + K::OFF_W + sl * K::PIECE);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a + 2 * k, w + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&w_empty[sl]);
          }
          umma_commit(&d_full[b]);
          TRACE(9 + j);
        }
        umma_commit(&a_empty[ab]);   // every MMA of this tile has been issued: its A tile may be rewritten once they complete
      }
    }
  } else if (warp >= LN_WARP0) {
    // ------------------------------------------------------------------ gather + residual copy + LayerNorm -> bf16 A tile
    const int t = threadIdx.x - LN_WARP0 * 32;   // 0..255
    const int l16 = t & 15;
    const int team = t >> 4;                     // rows team + 16 * pass
    constexpr int Q = K::NKB;                    // float4 per lane per row (one per 64-column k-block)
    constexpr int BATCH = 12 / Q;                // rows in flight per thread
    // row -> source row of x (the composed roll + window_partition gather works within a frame of T tokens)
    auto src_row = [&](int row) -> int {
      if (p.gather == nullptr) return row;
      const int fr = row / p.T;
      return fr * p.T + __ldg(p.gather + (row - fr * p.T));
    };
    for (int i = 0; i < n_local; ++i) {
      const int tile = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int m0 = tile * TILE_M;
      bool waited = false;
      const uint32_t ab = K::NA == 2 ? (i & 1u) : 0u;
      const uint32_t apar = K::NA == 2 ? ((i >> 1) & 1u) : (i & 1u);
      if (t == 0) TRACE(0);
#pragma unroll 1
      for (int pass0 = 0; pass0 < 8; pass0 += BATCH) {
        float4 xv[BATCH][Q];
        // every load of the batch is issued before the first store: x and x_raw may alias as far as the compiler knows,
        // and a store placed between two rows' loads serialises their HBM round trips
        int src[BATCH];
#pragma unroll
        for (int bq = 0; bq < BATCH; ++bq) {
          const int row = m0 + team + 16 * (pass0 + bq);
          src[bq] = row < p.M ? src_row(row) : -1;
        }
#pragma unroll
        for (int bq = 0; bq < BATCH; ++bq) {
#pragma unroll
          for (int q = 0; q < Q; ++q)
            xv[bq][q] = src[bq] >= 0 ? __ldcg(reinterpret_cast<const float4*>(p.x + static_cast<size_t>(src[bq]) * C + 4 * l16 + 64 * q))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (pass0 + BATCH >= 8 && i + 1 < n_local) {
          // pull the next tile's rows towards L2 while this tile is multiplied (the LayerNorm is on the critical path: the A
          // tile is single-buffered)
          const int nrow = m0 + static_cast<int>(gridDim.x) * TILE_M + (t >> 1);
          if (nrow < p.M) prefetch_l2_bulk(p.x + static_cast<size_t>(src_row(nrow)) * C + (t & 1) * (C / 2), C * 2);
        }
        if (t == 0 && pass0 == BATCH) TRACE(27);
        if (p.x_raw != nullptr && !(p.dbg & 1)) {
#pragma unroll
          for (int bq = 0; bq < BATCH; ++bq) {
            const int row = m0 + team + 16 * (pass0 + bq);
            if (src[bq] >= 0) {
#pragma unroll
              for (int q = 0; q < Q; ++q)
                *reinterpret_cast<float4*>(p.x_raw + static_cast<size_t>(row) * C + 4 * l16 + 64 * q) = xv[bq][q];
            }
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
        if (t == 0) TRACE(1 + pass0 / BATCH);
        if (t == 0 && pass0 == BATCH) TRACE(28);
        if (!waited) {
          mbar_wait_relaxed(&a_empty[ab], apar ^ 1u, 94, 500);   // the MMAs of the previous use of this A tile are complete
          waited = true;
          if (t == 0) TRACE(5);
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
            uint8_t* dst = smem + K::OFF_A + ab * K::A_BYTES + q * 16384 + r * 128 + ((((kc >> 3) ^ (r & 7))) << 4) + (kc & 7) * 2;
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
          }
        }
        if (t == 0 && pass0 == BATCH) TRACE(29);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[ab]);
      if (t == 0) TRACE(6);
    }
  } else {
    // ------------------------------------------------------------------ drain groups: accumulator -> + bias -> bf16 -> TMA store
    const int group = warp >> 2;          // 64-column slabs group, group + 2 of every chunk
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int sw = row & 7;
    const bool elected = (threadIdx.x & 127) == 0;
    uint8_t* slab = smem + K::OFF_OUT + group * 16384;
    uint8_t* my_out = slab + row * 128;
    uint32_t cnt[2] = {0u, 0u};
    for (int i = 0; i < n_local; ++i) {
      const int m0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TILE_M;
#pragma unroll 1
      for (int j = 0; j < p.chunks; ++j) {
        const int bn = p.N - j * BN < BN ? p.N - j * BN : BN;
        const uint32_t b = j & 1u;
        mbar_wait(&d_full[b], cnt[b] & 1u, 95);
        if (threadIdx.x == 0) TRACE(16 + j);
        ++cnt[b];
        tc_fence_after();
        const int nslab = bn >> 6;
#pragma unroll 1
        for (int sl = group; sl < nslab; sl += 2) {
          uint32_t va[32], vb[32];
          const uint32_t ta = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256u * b + 64u * sl;
          tmem_ld_32x32b_x32(ta, va);
          tmem_ld_32x32b_x32(ta + 32u, vb);
          tmem_ld_wait();
          const float* bias = p.bias + j * BN + 64 * sl;
          uint32_t pk[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + q);
            pk[2 * q] = pack_bf16(__uint_as_float(va[4 * q]) + b4.x, __uint_as_float(va[4 * q + 1]) + b4.y);
            pk[2 * q + 1] = pack_bf16(__uint_as_float(va[4 * q + 2]) + b4.z, __uint_as_float(va[4 * q + 3]) + b4.w);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + 32) + q);
            pk[16 + 2 * q] = pack_bf16(__uint_as_float(vb[4 * q]) + b4.x, __uint_as_float(vb[4 * q + 1]) + b4.y);
            pk[16 + 2 * q + 1] = pack_bf16(__uint_as_float(vb[4 * q + 2]) + b4.z, __uint_as_float(vb[4 * q + 3]) + b4.w);
          }
          if (elected) tma_store_wait_read<0>();    // the previous store from this slab has read its data
          named_bar_sync(1 + group, 128);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(my_out + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          fence_proxy_async_smem();
          named_bar_sync(1 + group, 128);
          if (elected && !(p.dbg & 2)) {
            tma_store_2d(&tmOut, slab, j * BN + 64 * sl, m0);   // rows >= M are clipped by the tensor map
            tma_store_commit();
          }
        }
        tc_fence_before();
        mbar_arrive(&d_empty[b]);     // this thread's reads of the chunk's accumulator are done
        if (threadIdx.x == 0) TRACE(22 + j);
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

template <int C>
cudaError_t launch_c(const LnQkvArgs& a, cudaStream_t stream) {
  using K = Cfg<C>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 148;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(ln_qkv_stream_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM + 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  });
  if (attr_err != cudaSuccess) return attr_err;
  CUtensorMap tmW, tmWrem, tmOut;
  const int rem = a.N % BN;
  if (!make_tmap_2d(&tmW, a.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.N, C, a.ldw, 64, a.N >= BN ? BN : rem)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&tmWrem, a.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.N, C, a.ldw, 64, rem > 0 ? rem : BN)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&tmOut, a.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, a.N, a.ldo, 64, TILE_M)) return cudaErrorInvalidValue;
  LnQkvParams p{a.x, a.x_raw, a.M, a.T, (a.M + TILE_M - 1) / TILE_M, (a.N + BN - 1) / BN, a.N, a.gather, a.gamma, a.beta, a.eps,
                a.bias, a.trace, a.dbg};
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  ln_qkv_stream_kernel<C><<<grid, THREADS, K::SMEM + 1024, stream>>>(tmW, tmWrem, tmOut, p);
  return cudaGetLastError();
}

}  // namespace

FMMT_DEFINE_WATCHDOG_ADDR(watchdog_addr_ln_qkv)

cudaError_t launch_ln_qkv(const LnQkvArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || !a.x || !a.gamma || !a.beta || !a.w || !a.bias || !a.out) return cudaErrorInvalidValue;
  if (!ln_qkv_supported(a.C, a.N) || (a.ldw % 8) != 0 || (a.ldo % 8) != 0) return cudaErrorInvalidValue;
  if (a.gather != nullptr && (a.T <= 0 || (a.M % a.T) != 0)) return cudaErrorInvalidValue;
  if (a.gather != nullptr && a.x_raw == nullptr) return cudaErrorInvalidValue;   // the gathered rows are the new residual stream
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.w) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15) ||
      (reinterpret_cast<uintptr_t>(a.gamma) & 15) || (reinterpret_cast<uintptr_t>(a.beta) & 15) ||
      (reinterpret_cast<uintptr_t>(a.bias) & 15) || (a.x_raw && (reinterpret_cast<uintptr_t>(a.x_raw) & 15)))
    return cudaErrorInvalidValue;
  return a.C == 192 ? launch_c<192>(a, stream) : launch_c<384>(a, stream);
}

}  // namespace fmmt
