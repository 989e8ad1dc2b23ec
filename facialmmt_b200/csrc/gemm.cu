// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma (fp32 accumulators in TMEM,
//   double-buffered) -> tcgen05.ld epilogue with fused bias / GELU(erf) / tanh / residual / row scatter.
// One CTA per SM; warps 0-7 = epilogue (two per TMEM lane quarter, smem-transposed coalesced stores), warp 8 = TMA
// producer, warp 9 = MMA issuer + TMEM owner.
//
// Replaces every nn.Linear on the reference path (e.g. Swin_Transformer.py:24-30,119,142,325; Transformer.py:87-89,
// 132,145,159; multihead_attention.py:151-158; CrossmodalTransformer.py:157-160; src/models.py:107,158,165).
#include "gemm.cuh"
#include "ptx.cuh"

#include <cmath>
#include <mutex>

namespace fmmt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;          // 2 per TMEM lane quarter
constexpr int PRODUCER_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int GEMM_THREADS = 320;
constexpr int EPI_PITCH = 36;         // floats per staged row (32 + 4 pad): conflict-free v4 writes and row reads
constexpr int EPI_STAGING_BYTES = ((EPI_WARPS * 32 * EPI_PITCH * 4 + 1023) / 1024) * 1024;
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 512;       // two accumulator stages of up to 256 fp32 columns
constexpr int ACC_STRIDE = 256;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr int SMEM_BUDGET = 220 * 1024;

struct GemmKernelParams {
  int M, N, K;
  int block_n, num_stages, stage_bytes;
  int m_tiles, n_tiles, num_kb;
  const float* bias;
  int act;
  const float* residual;
  int ldr;
  int res_mod;
  float* out_f32;
  int ldo32;
  __nv_bfloat16* out_bf16;
  int ldo16;
  const int* row_map;
  int map_period;
  int rows_in, rows_out, row_off;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // SWIZZLE_128B tiles need 1024-byte alignment. Layout: [epilogue staging | pipeline stages]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* stage_f32 = reinterpret_cast<float*>(smem_gen);            // EPI_WARPS x [32][EPI_PITCH] floats
  uint8_t* pipe_gen = smem_gen + EPI_STAGING_BYTES;
  const uint32_t pipe_base = smem_base + EPI_STAGING_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == MMA_WARP) {
    tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == PRODUCER_WARP) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles) * BM;
        const int n0 = (tile % p.n_tiles) * p.block_n;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = pipe_gen + stage * p.stage_bytes;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.stage_bytes));
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * ACC_STRIDE);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = pipe_base + stage * p.stage_bytes;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + A_TILE_BYTES);
          int ksteps = (p.K - kb * BK + UMMA_K - 1) / UMMA_K;
          if (ksteps > BK / UMMA_K) ksteps = BK / UMMA_K;
          for (int k = 0; k < ksteps; ++k) {
            // advancing 16 bf16 (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 0..7
    // Warp w owns TMEM lanes 32*(w%4).. (hardware restriction) and the 32-column chunks with parity w/4.
    // Each chunk: tcgen05.ld (thread = row) -> per-warp smem transpose -> lanes (8 per row, float4 each) apply
    // bias / activation / residual and store 128-byte row segments: fully coalesced global traffic.
    const int quarter = warp & 3;
    const int parity = warp >> 2;
    float* st = stage_f32 + warp * (32 * EPI_PITCH);
    const int sub = lane >> 3;   // row within a group of 4
    const int c4 = lane & 7;     // float4 column within the 32-column chunk
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.n_tiles) * BM;
      const int n0 = (tile % p.n_tiles) * p.block_n;
      // destination rows of the 8 rows this lane stores: local row 4*i + sub
      long long drow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = m0 + quarter * 32 + 4 * i + sub;
        long long dest = -1;
        if (row < p.M) {
          dest = row;
          if (p.row_map != nullptr) {
            const int q = row / p.map_period;
            dest = static_cast<long long>(q) * p.map_period + __ldg(p.row_map + (row - q * p.map_period));
          }
          if (p.rows_in > 0) {
            const long long q = dest / p.rows_in;
            dest = q * p.rows_out + p.row_off + (dest - q * p.rows_in);
          }
        }
        drow[i] = dest;
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * ACC_STRIDE);
      for (int c = parity * 32; c < p.block_n; c += 64) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(st + lane * EPI_PITCH + 4 * j) =
              make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int gc = n0 + c + 4 * c4;  // first of this lane's 4 columns
        if (gc < p.N) {
        const bool full4 = gc + 4 <= p.N;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) {
          if (full4) {
            bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + gc));
          } else {
            bias4.x = p.bias[gc];
            if (gc + 1 < p.N) bias4.y = p.bias[gc + 1];
            if (gc + 2 < p.N) bias4.z = p.bias[gc + 2];
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long dest = drow[i];
          if (dest < 0) continue;
          float4 x = *reinterpret_cast<const float4*>(st + (4 * i + sub) * EPI_PITCH + 4 * c4);
          x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
          if (p.act != ACT_NONE) {
            x.x = apply_act(x.x, p.act); x.y = apply_act(x.y, p.act);
            x.z = apply_act(x.z, p.act); x.w = apply_act(x.w, p.act);
          }
          if (full4) {
            if (p.residual != nullptr) {
              const long long rrow = (p.res_mod > 0) ? (dest % p.res_mod) : dest;
              const float4 r = *reinterpret_cast<const float4*>(p.residual + rrow * p.ldr + gc);
              x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
            }
            if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + dest * p.ldo32 + gc) = x;
            if (p.out_bf16 != nullptr)
              *reinterpret_cast<uint2*>(p.out_bf16 + dest * p.ldo16 + gc) =
                  make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
          } else {
            // ragged last columns (N not a multiple of 4): scalar, predicated
            const float xs[4] = {x.x, x.y, x.z, x.w};
            const long long rrow = (p.res_mod > 0) ? (dest % p.res_mod) : dest;
            for (int k = 0; k < 4 && gc + k < p.N; ++k) {
              float y = xs[k];
              if (p.residual != nullptr) y += p.residual[rrow * p.ldr + gc + k];
              if (p.out_f32 != nullptr) p.out_f32[dest * p.ldo32 + gc + k] = y;
              if (p.out_bf16 != nullptr) p.out_bf16[dest * p.ldo16 + gc + k] = __float2bfloat16(y);
            }
          }
        }
        }
        __syncwarp();  // staged chunk fully consumed (and the warp reconverged) before the next tcgen05.ld / overwrite
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ====================================================================================================================
// Fast path: TMA-fed epilogue. The register-path epilogue above keeps only a few KB of loads in flight per SM
// (Little's law caps it near 1 TB/s); here the residual tile is TMA-loaded into a shared-memory ring by its own
// producer warp, and results are staged in 128B-swizzled shared-memory slabs and written with TMA stores, so the SM
// always has >100 KB of bulk traffic in flight and no thread ever waits on a global load.
//   warps 0-3 / 4-7 : two epilogue groups (each covers the 128 accumulator rows; they alternate over column slabs)
//   warp 8 : A/B TMA producer   warp 9 : MMA issuer + TMEM owner   warp 10 : residual-slab TMA producer
// A slab is 128 rows x 128 bytes (32 fp32 or 64 bf16 columns), SWIZZLE_128B, so thread-per-row 16-byte accesses are
// bank-conflict free (chunk j of row r lives at chunk j ^ (r & 7)).
constexpr int EPI_GROUPS = 4;                       // 4 groups x 4 warps: every SMSP hosts 4 epilogue warps
constexpr int FAST_EPI_WARPS = 4 * EPI_GROUPS;
constexpr int FAST_PRODUCER_WARP = FAST_EPI_WARPS;
constexpr int FAST_MMA_WARP = FAST_EPI_WARPS + 1;
constexpr int FAST_RES_WARP = FAST_EPI_WARPS + 2;
constexpr int FAST_THREADS = (FAST_EPI_WARPS + 3) * 32;   // 608
constexpr int SLAB_BYTES = 128 * 128;
constexpr int RES_SLOTS = EPI_GROUPS;   // slot g is produced for / consumed by group g only

struct FastParams {
  int M, N, K;
  int block_n, num_stages, stage_bytes;
  int m_tiles, n_tiles, num_kb;
  const float* bias;
  int act;
  int has_res;
  int out_bf16;    // 1: bf16 output (64-column slabs), 0: fp32 output (32-column slabs)
  int slab_cols;
  int groups;      // active epilogue groups (2 or 4) == staging slabs == residual ring slots: compute-heavy GEMMs
                   // (large K) give the shared memory to the A/B pipeline instead of the epilogue rings
};

template <int ACT>
__device__ __forceinline__ float act_fn(float x, int act_rt) {
  if (ACT == ACT_NONE) return x;
  if (ACT == ACT_GELU) return gelu_erf(x);
  return apply_act(x, act_rt);
}

// bias + activation on 32 accumulator columns of this thread's row, in place (v holds fp32 bit patterns)
template <int ACT>
__device__ __forceinline__ void epi_math32(uint32_t (&v)[32], const float* bias, int gc, int N, int act_rt) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr && gc + 4 * j + 4 <= N) b4 = __ldg(reinterpret_cast<const float4*>(bias + gc) + j);
    v[4 * j + 0] = __float_as_uint(act_fn<ACT>(__uint_as_float(v[4 * j + 0]) + b4.x, act_rt));
    v[4 * j + 1] = __float_as_uint(act_fn<ACT>(__uint_as_float(v[4 * j + 1]) + b4.y, act_rt));
    v[4 * j + 2] = __float_as_uint(act_fn<ACT>(__uint_as_float(v[4 * j + 2]) + b4.z, act_rt));
    v[4 * j + 3] = __float_as_uint(act_fn<ACT>(__uint_as_float(v[4 * j + 3]) + b4.w, act_rt));
  }
}
__device__ __forceinline__ uint32_t pack_bf16_bits(uint32_t a, uint32_t b) {
  return pack_bf16(__uint_as_float(a), __uint_as_float(b));
}

template <int ACT>
__global__ void __launch_bounds__(FAST_THREADS, 1)
gemm_bf16_tcgen05_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmOut,
                             const FastParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint64_t res_full_bar[RES_SLOTS];
  __shared__ uint64_t res_empty_bar[RES_SLOTS];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* out_ring = smem_gen;                                   // p.groups slabs
  uint8_t* res_ring = out_ring + p.groups * SLAB_BYTES;           // p.groups slabs (only if has_res)
  const int ring_bytes = (p.has_res ? 2 : 1) * p.groups * SLAB_BYTES;
  uint8_t* pipe_gen = smem_gen + ring_bytes;
  const uint32_t pipe_base = smem_base + ring_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], FAST_EPI_WARPS * 32);
    }
    for (int s = 0; s < RES_SLOTS; ++s) {
      mbar_init(&res_full_bar[s], 1);
      mbar_init(&res_empty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == FAST_PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (p.has_res) tma_prefetch_desc(&tmRes);
  }
  if (warp == FAST_MMA_WARP) {
    tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == FAST_PRODUCER_WARP) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles) * BM;
        const int n0 = (tile % p.n_tiles) * p.block_n;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          uint8_t* sa = pipe_gen + stage * p.stage_bytes;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.stage_bytes));
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == FAST_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * ACC_STRIDE);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = pipe_base + stage * p.stage_bytes;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + A_TILE_BYTES);
          int ksteps = (p.K - kb * BK + UMMA_K - 1) / UMMA_K;
          if (ksteps > BK / UMMA_K) ksteps = BK / UMMA_K;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp == FAST_RES_WARP) {
    // ------------------------------------------------------------ residual slab producer
    if (lane == 0 && p.has_res) {
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles) * BM;
        const int n0 = (tile % p.n_tiles) * p.block_n;
        int ncols = p.N - n0;
        if (ncols > p.block_n) ncols = p.block_n;
        const int nsl = (ncols + p.slab_cols - 1) / p.slab_cols;
        for (int s = 0; s < nsl; ++s, ++cnt) {
          const int slot = cnt % p.groups;
          const uint32_t ph = (cnt / p.groups) & 1u;
          mbar_wait(&res_empty_bar[slot], ph ^ 1u, 4);
          mbar_arrive_expect_tx(&res_full_bar[slot], SLAB_BYTES);
          tma_load_2d(res_ring + slot * SLAB_BYTES, &tmRes, &res_full_bar[slot], n0 + s * p.slab_cols, m0);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue groups
    const int group = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // accumulator row == TMEM lane
    const int sw = row & 7;                        // 128B-swizzle phase of this row
    const bool elected = (threadIdx.x & 127) == 0;
    uint8_t* out_slot = out_ring + group * SLAB_BYTES;
    uint8_t* my_out = out_slot + row * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_base = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.n_tiles) * BM;
      const int n0 = (tile % p.n_tiles) * p.block_n;
      int ncols = p.N - n0;
      if (ncols > p.block_n) ncols = p.block_n;
      const int nsl = (ncols + p.slab_cols - 1) / p.slab_cols;
      mbar_wait(&tmem_full_bar[acc], acc_phase, 5);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * ACC_STRIDE);
      // Slab with running number c = res_base + s goes to group c % EPI_GROUPS (== its residual-ring slot): work
      // rotates over the groups from tile to tile, and every group consumes EVERY use of "its" ring slot in order,
      // which is what makes the parity waits on res_full/res_empty alias-free.
      const uint32_t ng = static_cast<uint32_t>(p.groups);
      int s_first = nsl;   // inactive groups only take part in the accumulator hand-shake
      if (group < p.groups) s_first = static_cast<int>((static_cast<uint32_t>(group) + ng - res_base % ng) % ng);
      for (int s = s_first; s < nsl; s += p.groups) {
        const int gc0 = n0 + s * p.slab_cols;
        if (p.out_bf16) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(s * 64), v);
          tmem_ld_wait();
          epi_math32<ACT>(v, p.bias, gc0, p.N, p.act);
          if (elected) tma_store_wait_read<0>();             // previous store from this slot has read its data
          named_bar_sync(1 + group, 128);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(my_out + ((j ^ sw) << 4)) =
                make_uint4(pack_bf16_bits(v[8 * j], v[8 * j + 1]), pack_bf16_bits(v[8 * j + 2], v[8 * j + 3]),
                           pack_bf16_bits(v[8 * j + 4], v[8 * j + 5]), pack_bf16_bits(v[8 * j + 6], v[8 * j + 7]));
          tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(s * 64 + 32), v);
          tmem_ld_wait();
          epi_math32<ACT>(v, p.bias, gc0 + 32, p.N, p.act);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(my_out + (((4 + j) ^ sw) << 4)) =
                make_uint4(pack_bf16_bits(v[8 * j], v[8 * j + 1]), pack_bf16_bits(v[8 * j + 2], v[8 * j + 3]),
                           pack_bf16_bits(v[8 * j + 4], v[8 * j + 5]), pack_bf16_bits(v[8 * j + 6], v[8 * j + 7]));
        } else {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(s * 32), v);
          tmem_ld_wait();
          epi_math32<ACT>(v, p.bias, gc0, p.N, p.act);
          if (p.has_res) {
            const uint32_t cnt = res_base + static_cast<uint32_t>(s);
            const int slot = cnt % p.groups;   // == group
            mbar_wait(&res_full_bar[slot], (cnt / p.groups) & 1u, 6);
            const uint8_t* my_res = res_ring + slot * SLAB_BYTES + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r = *reinterpret_cast<const float4*>(my_res + ((j ^ sw) << 4));
              v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + r.x);
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + r.y);
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + r.z);
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + r.w);
            }
            mbar_arrive(&res_empty_bar[slot]);               // this thread is done reading the residual slab
          }
          if (elected) tma_store_wait_read<0>();
          named_bar_sync(1 + group, 128);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(my_out + ((j ^ sw) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + group, 128);
        if (elected) {
          tma_store_2d(&tmOut, out_slot, gc0, m0);
          tma_store_commit();
        }
      }
      res_base += static_cast<uint32_t>(nsl);
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (elected) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FAST_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

// 2-D tensor map: dim0 = columns (contiguous), dim1 = rows; 128-byte swizzle (box_cols * esize must be 128); OOB
// reads give zeros, OOB writes are clipped.
bool make_tmap(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int esize, long long rows, long long cols,
               long long ld_elems, int box_cols, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld_elems) * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int pick_block_n(int M, int N, int K, int num_sms, int gran = 32, int max_bn = 256) {
  // Candidates are multiples of the epilogue granularity (32-column chunk, or the TMA slab width) up to the
  // 256-column UMMA limit. Cost model fitted to B200 sweeps (tests/gpu_gemm_shapes.py, SWEEP=1): a tile costs a fixed
  // ~0.9 us (barrier round trips, TMEM drain, store hand-off) plus the larger of its operand feed ((128 + bn) rows of
  // K bf16 at ~45 B/ns per SM) and its epilogue (~6 ns per output column); the launch costs waves x tile.
  const int cands[] = {256, 224, 192, 160, 128, 96, 64, 32};
  const int m_tiles = (M + BM - 1) / BM;
  int best = gran;
  double best_cost = 1e30;
  for (int bn : cands) {
    if (bn % gran != 0 || bn > max_bn) continue;
    const int n_tiles = (N + bn - 1) / bn;
    const long long tiles = static_cast<long long>(m_tiles) * n_tiles;
    const double waves = static_cast<double>(tiles) / num_sms;
    const double waves_q = waves < 1.0 ? 1.0 : (waves > 4.0 ? waves : std::ceil(waves));
    const double feed = (128.0 + bn) * K * 2.0 / 45e3;
    const double epi = 0.006 * bn;
    const double cost = waves_q * (0.9 + (feed > epi ? feed : epi));
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int g_num_sms = 0;

}  // namespace

bool make_tmap_2d(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int esize, long long rows, long long cols,
                  long long ld_elems, int box_cols, int box_rows) {
  return make_tmap(tm, base, dt, esize, rows, cols, ld_elems, box_cols, box_rows);
}

unsigned int read_mbar_timeout(bool reset) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_mbar_timeout, sizeof(v));
  if (reset && v != 0) {
    unsigned int z = 0;
    cudaMemcpyToSymbol(g_mbar_timeout, &z, sizeof(z));
  }
  return v;
}

cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaErrorInvalidValue;
  if ((a.lda % 8) != 0 || (a.ldw % 8) != 0) return cudaErrorInvalidValue;  // TMA: 16-byte global strides
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return cudaErrorInvalidValue;
  if (a.out_f32 && ((a.ldo32 % 4) != 0 || (reinterpret_cast<uintptr_t>(a.out_f32) & 15))) return cudaErrorInvalidValue;
  if (a.out_bf16 && ((a.ldo16 % 8) != 0 || (reinterpret_cast<uintptr_t>(a.out_bf16) & 15))) return cudaErrorInvalidValue;
  if (a.residual && ((a.ldr % 4) != 0 || (reinterpret_cast<uintptr_t>(a.residual) & 15))) return cudaErrorInvalidValue;
  if (a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)) return cudaErrorInvalidValue;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    SMEM_BUDGET + 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  });
  if (attr_err != cudaSuccess) return attr_err;

  const bool one_out = (a.out_f32 != nullptr) != (a.out_bf16 != nullptr);
  const bool fast = !a.force_generic && one_out && a.row_map == nullptr && a.rows_in == 0 && a.res_mod == 0 &&
                    (a.residual == nullptr || a.out_f32 != nullptr) && a.N >= 32;
  CUtensorMap tmA, tmB;
  if (fast) {
    static std::once_flag once2;
    static cudaError_t attr_err2 = cudaSuccess;
    std::call_once(once2, [] {
      attr_err2 = cudaFuncSetAttribute(gemm_bf16_tcgen05_tma_kernel<ACT_NONE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 1024);
      if (attr_err2 == cudaSuccess)
        attr_err2 = cudaFuncSetAttribute(gemm_bf16_tcgen05_tma_kernel<ACT_GELU>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 1024);
      if (attr_err2 == cudaSuccess)
        attr_err2 = cudaFuncSetAttribute(gemm_bf16_tcgen05_tma_kernel<ACT_TANH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 1024);
    });
    if (attr_err2 != cudaSuccess) return attr_err2;
    FastParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.out_bf16 = a.out_bf16 != nullptr;
    p.slab_cols = p.out_bf16 ? 64 : 32;
    p.has_res = a.residual != nullptr;
    p.block_n = a.block_n > 0 ? a.block_n : pick_block_n(a.M, a.N, a.K, g_num_sms, p.slab_cols, 256);
    if (p.block_n % p.slab_cols != 0 || p.block_n < 32 || p.block_n > 256) return cudaErrorInvalidValue;
    p.stage_bytes = A_TILE_BYTES + p.block_n * BK * 2;
    // epilogue-bound shapes (small K) want all 4 groups; compute-bound ones want pipeline depth
    p.groups = (a.K >= 512) ? 2 : EPI_GROUPS;
    int ring = (p.has_res ? 2 : 1) * p.groups * SLAB_BYTES;
    if ((SMEM_BUDGET - ring) / p.stage_bytes < 3 && p.groups > 2) {
      p.groups = 2;
      ring = (p.has_res ? 2 : 1) * p.groups * SLAB_BYTES;
    }
    p.num_stages = (SMEM_BUDGET - ring) / p.stage_bytes;
    if (p.num_stages > MAX_STAGES) p.num_stages = MAX_STAGES;
    if (p.num_stages < 2) return cudaErrorInvalidValue;
    p.m_tiles = (a.M + BM - 1) / BM;
    p.n_tiles = (a.N + p.block_n - 1) / p.block_n;
    p.num_kb = (a.K + BK - 1) / BK;
    p.bias = a.bias; p.act = a.act;
    CUtensorMap tmRes, tmOut;
    if (!make_tmap(&tmA, a.A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, a.K, a.lda, BK, BM)) return cudaErrorInvalidValue;
    if (!make_tmap(&tmB, a.W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.N, a.K, a.ldw, BK, p.block_n)) return cudaErrorInvalidValue;
    if (p.out_bf16) {
      if (!make_tmap(&tmOut, a.out_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, a.N, a.ldo16, 64, BM)) return cudaErrorInvalidValue;
    } else {
      if (!make_tmap(&tmOut, a.out_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, a.N, a.ldo32, 32, BM)) return cudaErrorInvalidValue;
    }
    if (p.has_res) {
      if (!make_tmap(&tmRes, a.residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, a.N, a.ldr, 32, BM)) return cudaErrorInvalidValue;
    } else {
      tmRes = tmOut;
    }
    const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
    const int grid = static_cast<int>(tiles < g_num_sms ? tiles : g_num_sms);
    const size_t smem = static_cast<size_t>(p.num_stages) * p.stage_bytes + ring + 1024;
    if (a.act == ACT_NONE)
      gemm_bf16_tcgen05_tma_kernel<ACT_NONE><<<grid, FAST_THREADS, smem, stream>>>(tmA, tmB, tmRes, tmOut, p);
    else if (a.act == ACT_GELU)
      gemm_bf16_tcgen05_tma_kernel<ACT_GELU><<<grid, FAST_THREADS, smem, stream>>>(tmA, tmB, tmRes, tmOut, p);
    else  // runtime-switched activation (ReLU / tanh)
      gemm_bf16_tcgen05_tma_kernel<ACT_TANH><<<grid, FAST_THREADS, smem, stream>>>(tmA, tmB, tmRes, tmOut, p);
    return cudaGetLastError();
  }

  GemmKernelParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.block_n = a.block_n > 0 ? a.block_n : pick_block_n(a.M, a.N, a.K, g_num_sms);
  if (p.block_n % 32 != 0 || p.block_n < 32 || p.block_n > 256) return cudaErrorInvalidValue;
  p.stage_bytes = A_TILE_BYTES + p.block_n * BK * 2;
  p.num_stages = (SMEM_BUDGET - EPI_STAGING_BYTES) / p.stage_bytes;
  if (p.num_stages > MAX_STAGES) p.num_stages = MAX_STAGES;
  p.m_tiles = (a.M + BM - 1) / BM;
  p.n_tiles = (a.N + p.block_n - 1) / p.block_n;
  p.num_kb = (a.K + BK - 1) / BK;
  p.bias = a.bias; p.act = a.act;
  p.residual = a.residual; p.ldr = a.ldr; p.res_mod = a.res_mod;
  p.out_f32 = a.out_f32; p.ldo32 = a.ldo32;
  p.out_bf16 = a.out_bf16; p.ldo16 = a.ldo16;
  p.row_map = a.row_map; p.map_period = a.map_period;
  p.rows_in = a.rows_in; p.rows_out = a.rows_out; p.row_off = a.row_off;

  if (!make_tmap(&tmA, a.A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, a.K, a.lda, BK, BM)) return cudaErrorInvalidValue;
  if (!make_tmap(&tmB, a.W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.N, a.K, a.ldw, BK, p.block_n)) return cudaErrorInvalidValue;

  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
  const int grid = static_cast<int>(tiles < g_num_sms ? tiles : g_num_sms);
  const size_t smem = static_cast<size_t>(p.num_stages) * p.stage_bytes + EPI_STAGING_BYTES + 1024;
  gemm_bf16_tcgen05_kernel<<<grid, GEMM_THREADS, smem, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

}  // namespace fmmt
