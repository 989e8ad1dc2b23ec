// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma (fp32 accumulators in TMEM,
//   double-buffered) -> tcgen05.ld epilogue with fused bias / GELU(erf) / tanh / residual / row scatter.
// One CTA per SM; warps 0-3 = epilogue (TMEM lane quarters), warp 4 = TMA producer, warp 5 = MMA issuer + TMEM owner.
//
// Replaces every nn.Linear on the reference path (e.g. Swin_Transformer.py:24-30,119,142,325; Transformer.py:87-89,
// 132,145,159; multihead_attention.py:151-158; CrossmodalTransformer.py:157-160; src/models.py:107,158,165).
#include "gemm.cuh"
#include "ptx.cuh"

#include <mutex>

namespace fmmt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;
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
  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 5) {
    tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles) * BM;
        const int n0 = (tile % p.n_tiles) * p.block_n;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem_gen + stage * p.stage_bytes;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.stage_bytes));
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
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
          const uint32_t sa = smem_base + stage * p.stage_bytes;
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
    // ------------------------------------------------------------ epilogue warps 0..3 (TMEM lanes 32*warp..)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.n_tiles) * BM;
      const int n0 = (tile % p.n_tiles) * p.block_n;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + warp * 32 + lane;
      const bool row_ok = row < p.M;
      long long dest = row;
      if (row_ok) {
        if (p.row_map != nullptr) {
          const int q = row / p.map_period;
          dest = static_cast<long long>(q) * p.map_period + p.row_map[row - q * p.map_period];
        }
        if (p.rows_in > 0) {
          const long long q = dest / p.rows_in;
          dest = q * p.rows_out + p.row_off + (dest - q * p.rows_in);
        }
      }
      const long long rrow = (p.res_mod > 0) ? (dest % p.res_mod) : dest;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) +
                             static_cast<uint32_t>(acc * ACC_STRIDE);
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c), v);
        tmem_ld_wait();
        const int gc = n0 + c;
        if (!row_ok || gc >= p.N) continue;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (gc + 32 <= p.N) {
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + gc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              x[4 * j + 0] += b.x; x[4 * j + 1] += b.y; x[4 * j + 2] += b.z; x[4 * j + 3] += b.w;
            }
          }
          if (p.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = apply_act(x[j], p.act);
          }
          if (p.residual != nullptr) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + rrow * p.ldr + gc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r = r4[j];
              x[4 * j + 0] += r.x; x[4 * j + 1] += r.y; x[4 * j + 2] += r.z; x[4 * j + 3] += r.w;
            }
          }
          if (p.out_f32 != nullptr) {
            float4* o4 = reinterpret_cast<float4*>(p.out_f32 + dest * p.ldo32 + gc);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          }
          if (p.out_bf16 != nullptr) {
            uint4* o4 = reinterpret_cast<uint4*>(p.out_bf16 + dest * p.ldo16 + gc);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o4[j] = make_uint4(pack_bf16(x[8 * j], x[8 * j + 1]), pack_bf16(x[8 * j + 2], x[8 * j + 3]),
                                 pack_bf16(x[8 * j + 4], x[8 * j + 5]), pack_bf16(x[8 * j + 6], x[8 * j + 7]));
          }
        } else {
          // ragged last column chunk (N not a multiple of 32): scalar, predicated
          for (int j = 0; j < 32 && gc + j < p.N; ++j) {
            float y = x[j];
            if (p.bias != nullptr) y += p.bias[gc + j];
            y = apply_act(y, p.act);
            if (p.residual != nullptr) y += p.residual[rrow * p.ldr + gc + j];
            if (p.out_f32 != nullptr) p.out_f32[dest * p.ldo32 + gc + j] = y;
            if (p.out_bf16 != nullptr) p.out_bf16[dest * p.ldo16 + gc + j] = __float2bfloat16(y);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
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

// 2-D bf16 tensor map: dim0 = K (contiguous), dim1 = rows; box = {64, box_rows}; 128-byte swizzle; OOB -> zeros.
bool make_tmap(CUtensorMap* tm, const void* base, int rows, int cols, int ld_elems, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int pick_block_n(int M, int N, int num_sms) {
  // Candidates are multiples of 32 (epilogue chunk) up to the 256-column UMMA limit. Minimise padded columns
  // first, then prefer a tile count that fills the SMs, then the larger tile.
  const int cands[] = {256, 192, 128, 96, 64, 32};
  const int m_tiles = (M + BM - 1) / BM;
  int best = 32;
  double best_cost = 1e30;
  for (int bn : cands) {
    const int n_tiles = (N + bn - 1) / bn;
    const double padded = static_cast<double>(n_tiles) * bn;
    const long long tiles = static_cast<long long>(m_tiles) * n_tiles;
    const long long waves = (tiles + num_sms - 1) / num_sms;
    // time ~ waves * (tile cost ~ bn columns + fixed per-tile overhead); padded columns are paid for as tile cost
    (void)padded;
    const double cost = static_cast<double>(waves) * (bn + 24.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int g_num_sms = 0;

}  // namespace

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

  GemmKernelParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.block_n = a.block_n > 0 ? a.block_n : pick_block_n(a.M, a.N, g_num_sms);
  if (p.block_n % 32 != 0 || p.block_n < 32 || p.block_n > 256) return cudaErrorInvalidValue;
  p.stage_bytes = A_TILE_BYTES + p.block_n * BK * 2;
  p.num_stages = SMEM_BUDGET / p.stage_bytes;
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

  CUtensorMap tmA, tmB;
  if (!make_tmap(&tmA, a.A, a.M, a.K, a.lda, BM)) return cudaErrorInvalidValue;
  if (!make_tmap(&tmB, a.W, a.N, a.K, a.ldw, p.block_n)) return cudaErrorInvalidValue;

  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
  const int grid = static_cast<int>(tiles < g_num_sms ? tiles : g_num_sms);
  const size_t smem = static_cast<size_t>(p.num_stages) * p.stage_bytes + 1024;
  gemm_bf16_tcgen05_kernel<<<grid, GEMM_THREADS, smem, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

}  // namespace fmmt
