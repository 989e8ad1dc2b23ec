// CTA-pair (tcgen05 cta_group::2) variant of the streamed fused Swin MLP half-block, C = 192 / 384 (Swin_Transformer.py:24-30
// Mlp.forward inside :264-268):
//     x <- x + fc2( GELU_erf( fc1( LayerNorm(x) ) ) )
// Same fusion as mlp_stream.cu, but a CLUSTER of two CTAs works on 256 rows: every tcgen05.mma is one M = 256 instruction over
// both SMs, each CTA holding its 128 rows of the A operands (the LayerNorm'd tile, the GELU'd hidden chunk) and HALF of the
// rows of every weight chunk. Why: the single-CTA kernel re-streams all of fc1/fc2 (2.36 MB at C = 384) from L2 per 128-row
// tile and has shared memory for exactly one chunk of weights, so the next chunk's boxes can only be requested when the
// current one is consumed and their L2 latency is exposed (about 1400 of 3470 cycles per chunk). Here the bytes
// per SM and per chunk halve, and the freed shared memory holds TWO chunks of fc1 and of fc2 weights in flight.
//   per CTA: A 96 KB (C = 384) | hidden 2 x 16 KB | fc1 piece slots 4 x 12 KB | fc2 slots 2 x 24 KB (= the drain staging)
// Protocol (barriers live at the same offsets in both CTAs; "leader" = cluster rank 0, whose MMA thread issues everything):
//   leader-side waits fed from BOTH CTAs:  a_full, hid_full (one release.cluster arrive per warp, after fence.proxy.async of
//     every lane), d1_empty, out_empty (one arrive per warp after its tcgen05.ld completed), w1_full / w2_full (the leader
//     expects the bytes of both CTAs' boxes; the peer's TMA credits the leader's barrier, cta_group::2 form);
//   multicast commits (tcgen05.commit ... multicast::cluster, both CTAs):  a_empty, w1_empty, w2_empty, d1_full, hid_empty,
//     out_full.
#include "mlp_pair.cuh"

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
constexpr int PROD1_WARP = MMA_WARP + 1;       // warp 25: fc1 weight stream, even pieces
constexpr int PROD2_WARP = MMA_WARP + 2;       // warp 26: fc2 weight stream
constexpr int PROD3_WARP = MMA_WARP + 3;       // warp 27: fc1 weight stream, odd pieces
constexpr int THREADS = (PROD3_WARP + 1) * 32; // 896
constexpr int DRAIN_STREAMS = 3;               // GELU warps 0..11, four warps (128 rows) per stream

template <int C>
struct Cfg {
  static_assert(C == 192 || C == 384, "streamed fused MLP: C = 192 or 384");
  static constexpr int H = 4 * C;
  static constexpr int NKB = C / 64;            // k-blocks of fc1 (A sub-tiles)
  static constexpr int NH = C / 192;            // 192-column halves of the fc2 output (one MMA each)
  static constexpr int CHUNKS = H / 64;
  static constexpr int A_BYTES = NKB * 16384;
  static constexpr int PIECES = NKB / 3;        // fc1 weights of a chunk arrive as pieces of 3 k-blocks
  static constexpr int W1_PIECE = 3 * 4096;     // this CTA's 32 hidden rows x 192 columns = 12 KB
  static constexpr int NS1 = 4;                 // fc1 piece slots (two chunks at C = 384, four at C = 192)
  static constexpr int W2_BYTES = NH * 12288;   // this CTA's 96 rows of each 192-row half x 64 bf16
  static constexpr int NS2 = 48 * 1024 / W2_BYTES;   // fc2 slots: 2 (C = 384) / 4 (C = 192) chunks
  static constexpr int OFF_A = 0;
  static constexpr int OFF_HID = A_BYTES;
  static constexpr int OFF_W1 = OFF_HID + 2 * 16384;
  static constexpr int OFF_W2 = OFF_W1 + NS1 * W1_PIECE;
  static constexpr int SMEM = OFF_W2 + 48 * 1024;      // fc2 slots == DRAIN_STREAMS x 16 KB of drain staging
  static constexpr int TM_OUT = 0;              // fc2 accumulator: C columns
  static constexpr int TM_D1 = 384;             // fc1 accumulators: 2 x 64 columns
  static constexpr int SLABS = C / 32;          // 32-column fp32 output slabs
  static_assert(NS2 * W2_BYTES == DRAIN_STREAMS * 16384, "fc2 slots double as the drain staging");
  static_assert(SMEM + 1024 <= 226 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// arrive on an mbarrier of a CTA of this cluster; everything this thread did / observed before is released to the cluster
__device__ __forceinline__ void arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait with cluster-scope acquire (the barrier receives arrivals from the peer CTA that publish its shared memory)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t tag) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    ++spins;
    if ((spins & 0x3FFF) == 0) {
      if (*reinterpret_cast<volatile unsigned int*>(&g_mbar_timeout) != 0) return;
      if (spins > (1u << 24)) {
        atomicCAS(&g_mbar_timeout, 0u, 0x80000000u | (tag << 24) | ((blockIdx.x & 0xFFF) << 12) | (threadIdx.x & 0xFFF));
        return;
      }
    }
  }
}

struct PairParams {
  const float* x;
  int M, num_ptiles;     // 256-row pair tiles
  const float* gamma;
  const float* beta;
  float eps;
  const float* b1;
  const float* b2;
};

template <int C>
__global__ void __launch_bounds__(THREADS, 1)
swin_mlp_pair_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                     const __grid_constant__ CUtensorMap tmX, const PairParams p) {
  using K = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full, a_empty;
  __shared__ uint64_t w1_full[K::NS1], w1_empty[K::NS1], w2_full[4], w2_empty[4];
  __shared__ uint64_t d1_full[2], d1_empty[2];
  __shared__ uint64_t hid_full[2], hid_empty[2];
  __shared__ uint64_t out_full, out_empty, drain_done;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    mbar_init(&a_full, 2 * LN_WARPS);           // one arrive per LayerNorm warp of either CTA
    mbar_init(&a_empty, 1);
    for (int s = 0; s < K::NS1; ++s) {
      mbar_init(&w1_full[s], 1);                // the leader's producer (expects the bytes of both CTAs' boxes)
      mbar_init(&w1_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&w2_full[s], 1);
      mbar_init(&w2_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d1_full[s], 1);
      mbar_init(&d1_empty[s], 2 * 8);           // one arrive per GELU warp of the group, either CTA
      mbar_init(&hid_full[s], 2 * 8);
      mbar_init(&hid_empty[s], 1);
    }
    mbar_init(&out_full, 1);
    mbar_init(&out_empty, 2 * DRAIN_STREAMS * 4);   // one arrive per draining warp, either CTA
    mbar_init(&drain_done, DRAIN_STREAMS);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc_pair(&tmem_base_slot, 512);
    tmem_relinquish_pair();
  }
  if (warp == PROD1_WARP && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive_release();      // the peer's barriers are initialised and its TMEM is allocated before anyone signals
  cluster_wait_acquire();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int pid = static_cast<int>(blockIdx.x >> 1);
  const int num_pairs = static_cast<int>(gridDim.x >> 1);
  const int n_local = (p.num_ptiles - pid + num_pairs - 1) / num_pairs;

  if (warp == PROD1_WARP || warp == PROD3_WARP) {
    // ------------------------------------------------------------------ fc1 weight streams (two threads: even / odd pieces)
    if (lane == 0) {
      const uint32_t mine = warp == PROD1_WARP ? 0u : 1u;
      uint32_t u = 0;   // running piece number of this pair: slot u % NS1, use u / NS1
      for (int i = 0; i < n_local; ++i)
        for (int j = 0; j < K::CHUNKS; ++j)
          for (int pc = 0; pc < K::PIECES; ++pc, ++u) {
            if ((u & 1u) != mine) continue;
            const uint32_t sl = u % K::NS1, use = u / K::NS1;
            mbar_wait(&w1_empty[sl], (use & 1u) ^ 1u, 40);           // the MMAs on this slot's previous piece are done
            // arrivals come from the leader only; the peer's bytes always land in the right phase because the peer refills a
            // slot only after the commit that followed the leader's wait on the slot's previous phase
            if (leader) mbar_arrive_expect_tx(&w1_full[sl], 2u * K::W1_PIECE);
            tma_load_3d_pair(smem + K::OFF_W1 + sl * K::W1_PIECE, &tmW1, mapa_u32(smem_u32(&w1_full[sl]), 0), 0,
                             64 * j + 32 * static_cast<int>(rank), 3 * pc);
          }
    }
  } else if (warp == PROD2_WARP) {
    // ------------------------------------------------------------------ fc2 weight stream (one thread)
    if (lane == 0) {
      uint32_t n = 0;   // running chunk number of this pair: slot n % NS2, use n / NS2
      for (int i = 0; i < n_local; ++i)
        for (int j = 0; j < K::CHUNKS; ++j, ++n) {
          const uint32_t sl = n % K::NS2, use = n / K::NS2;
          mbar_wait(&w2_empty[sl], (use & 1u) ^ 1u, 41);             // fc2 of this slot's previous chunk is done
          if (i > 0 && j < K::NS2) mbar_wait(&drain_done, (i - 1) & 1u, 42);   // ... and the drain staged here is out
          if (leader) mbar_arrive_expect_tx(&w2_full[sl], 2u * K::W2_BYTES);
          tma_load_3d_pair(smem + K::OFF_W2 + sl * K::W2_BYTES, &tmW2, mapa_u32(smem_u32(&w2_full[sl]), 0), 64 * j,
                           96 * static_cast<int>(rank), 0);
        }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issue (one thread of the leader CTA)
    if (lane == 0 && leader) {
      const uint32_t idesc1 = make_idesc_bf16(2 * TILE_M, 64);
      const uint32_t idesc2 = make_idesc_bf16(2 * TILE_M, 192);
      uint32_t n = 0;
      auto issue_fc1 = [&](uint32_t nn) {       // chunk nn of this pair -> TMEM buffer nn & 1
        const uint32_t b = nn & 1u, use = nn >> 1;
        mbar_wait(&d1_empty[b], (use & 1u) ^ 1u, 44);              // both CTAs' GELU groups have drained the buffer's last use
        tc_fence_after();
        const uint32_t d = tmem_base + K::TM_D1 + 64u * b;
#pragma unroll
        for (int pc = 0; pc < K::PIECES; ++pc) {
          const uint32_t u = nn * K::PIECES + pc, sl = u % K::NS1;
          mbar_wait(&w1_full[sl], (u / K::NS1) & 1u, 43);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const uint64_t a = make_smem_desc_sw128(smem_base + K::OFF_A + (3 * pc + kk) * 16384);
            const uint64_t w = make_smem_desc_sw128(smem_base + K::OFF_W1 + sl * K::W1_PIECE + kk * 4096);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(d, a + 2 * k, w + 2 * k, idesc1, (pc | kk | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&w1_empty[sl]);
        }
        umma_commit_pair(&d1_full[b]);
      };
      for (int i = 0; i < n_local; ++i) {
        mbar_wait_cluster(&a_full, i & 1u, 45);                    // both CTAs' normalised tiles are in their shared memory
        tc_fence_after();
        // fc1 runs TWO chunks ahead of fc2: fc1(j + 2) only needs the GELU group of chunk j to have READ its accumulator
        // (early in its pass), not to have finished, so each group finds its next accumulator ready when it comes back and
        // the two groups' GELU passes run back to back (with a look-ahead of one, fc1(j + 2) was issued after fc2(j), which
        // waits for the whole GELU pass of chunk j: GELU -> fc2 -> fc1 formed one serial chain per group)
        issue_fc1(n);
        issue_fc1(n + 1);
        for (int j = 0; j < K::CHUNKS; ++j, ++n) {
          if (j + 2 < K::CHUNKS) issue_fc1(n + 2);
          else if (j + 2 == K::CHUNKS) umma_commit_pair(&a_empty);   // every fc1 of this tile has been issued: A may be rewritten
          const uint32_t b = n & 1u, use = n >> 1;
          const uint32_t sl2 = n % K::NS2;
          mbar_wait(&w2_full[sl2], (n / K::NS2) & 1u, 46);
          mbar_wait_cluster(&hid_full[b], use & 1u, 47);           // both CTAs' hidden chunks are written
          if (j == 0 && i > 0) mbar_wait(&out_empty, (i - 1) & 1u, 48);   // previous tile's accumulator has been drained
          tc_fence_after();
          const uint64_t a = make_smem_desc_sw128(smem_base + K::OFF_HID + b * 16384);
#pragma unroll
          for (int h = 0; h < K::NH; ++h) {
            const uint64_t w = make_smem_desc_sw128(smem_base + K::OFF_W2 + sl2 * K::W2_BYTES + h * 12288);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(tmem_base + K::TM_OUT + 192u * h, a + 2 * k, w + 2 * k, idesc2, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&w2_empty[sl2]);
          umma_commit_pair(&hid_empty[b]);
        }
        umma_commit_pair(&out_full);
      }
    }
  } else if (warp >= LN_WARP0) {
    // ------------------------------------------------------------------ LayerNorm -> bf16 A tile (this CTA's 128 rows)
    const int t = threadIdx.x - LN_WARP0 * 32;   // 0..255
    const int l16 = t & 15;                      // lane within the 16-lane row team
    const int team = t >> 4;                     // 0..15: rows team + 16 * pass
    constexpr int Q = K::NKB;                    // float4 per lane per row (one per 64-column k-block)
    constexpr int BATCH = 12 / Q;                // rows in flight per thread: 48 data registers
    const uint32_t lead_a_full = mapa_u32(smem_u32(&a_full), 0);
    for (int i = 0; i < n_local; ++i) {
      const int ptile = pid + i * num_pairs;
      const int m0 = (2 * ptile + static_cast<int>(rank)) * TILE_M;
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
      __syncwarp();
      if (lane == 0) arrive_release_cluster(lead_a_full);
      // pull the next tile's rows towards L2 while this tile computes (its LayerNorm is on the critical path)
      if (i + 1 < n_local) {
        const int nm0 = (2 * (ptile + num_pairs) + static_cast<int>(rank)) * TILE_M;
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
    const uint32_t lead_d1_empty = mapa_u32(smem_u32(&d1_empty[group]), 0);
    const uint32_t lead_hid_full = mapa_u32(smem_u32(&hid_full[group]), 0);
    const uint32_t lead_out_empty = mapa_u32(smem_u32(&out_empty), 0);
    // drain: stream = warp / 4 (0..2), slabs stream, stream + 3, ...
    const int stream = warp >> 2;
    const bool drains = stream < DRAIN_STREAMS;
    const bool elected = drains && (threadIdx.x & 127) == 0;
    uint8_t* stage = smem + K::OFF_W2 + stream * 16384;
    uint8_t* my_out = stage + row * 128;
    for (int i = 0; i < n_local; ++i) {
      const int m0 = (2 * (pid + i * num_pairs) + static_cast<int>(rank)) * TILE_M;
#pragma unroll 1
      for (int jj = 0; jj < K::CHUNKS / 2; ++jj) {
        const int j = 2 * jj + group;
        const uint32_t use = static_cast<uint32_t>(i) * (K::CHUNKS / 2) + static_cast<uint32_t>(jj);
        mbar_wait(&d1_full[group], use & 1u, 50);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + K::TM_D1 + 64u * group + 32u * half, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_d1_empty);     // this warp no longer needs the fc1 accumulator
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
        __syncwarp();
        if (lane == 0) arrive_release_cluster(lead_hid_full);
      }
      if (drains) {
        // fc2 accumulator -> + bias -> 32-column slabs staged in the fc2 slots -> TMA reduce-add into x
        mbar_wait(&out_full, i & 1u, 52);
        tc_fence_after();
#pragma unroll 1
        for (int s = stream; s < K::SLABS; s += DRAIN_STREAMS) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + K::TM_OUT + 32u * s, v);
          tmem_ld_wait();
          if (s + DRAIN_STREAMS >= K::SLABS) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_out_empty);   // this warp's last read of the tile's accumulator
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
          tma_store_wait_read<0>();              // the fc2 slots may be refilled
          mbar_arrive(&drain_done);
        }
      }
    }
    if (elected) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_arrive_release();      // neither CTA may exit (or free TMEM) while the pair's MMAs / remote arrives are in flight
  cluster_wait_acquire();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encode encode_fn() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<PFN_encode>(ptr);
}

// 3-D view of fc2.weight [C rows, H cols]: (column, row within a 192-row half, half); box = 64 columns x 96 rows x all halves
bool make_tmap_w2(CUtensorMap* tm, const void* base, int C, int H, int ld) {
  PFN_encode f = encode_fn();
  if (!f) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(H), 192, static_cast<cuuint64_t>(C / 192)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * 192};
  cuuint32_t box[3] = {64, 96, static_cast<cuuint32_t>(C / 192)};
  cuuint32_t estr[3] = {1, 1, 1};
  return f(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C>
cudaError_t launch_c(const MlpStreamArgs& a, cudaStream_t stream) {
  using K = Cfg<C>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 148;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(swin_mlp_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM + 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  });
  if (attr_err != cudaSuccess) return attr_err;
  CUtensorMap tmW1, tmW2, tmX;
  if (!make_tmap_kblocks_2d(&tmW1, a.w1, K::H, C, a.ldw1, 32, 3)) return cudaErrorInvalidValue;   // 32 rows x 3 k-blocks
  if (!make_tmap_w2(&tmW2, a.w2, C, K::H, a.ldw2)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&tmX, a.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, C, C, 32, TILE_M)) return cudaErrorInvalidValue;
  PairParams p{a.x, a.M, (a.M + 2 * TILE_M - 1) / (2 * TILE_M), a.gamma, a.beta, a.eps, a.b1, a.b2};
  const int max_pairs = num_sms / 2;
  const int pairs = p.num_ptiles < max_pairs ? p.num_ptiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = K::SMEM + 1024;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, swin_mlp_pair_kernel<C>, tmW1, tmW2, tmX, p);
}

}  // namespace

FMMT_DEFINE_WATCHDOG_ADDR(watchdog_addr_mlp_pair)

cudaError_t launch_mlp_pair(const MlpStreamArgs& a, cudaStream_t stream) {
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
