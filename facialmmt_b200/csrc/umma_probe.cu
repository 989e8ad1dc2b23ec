// Layout probe for tcgen05.mma shared-memory operand descriptors (test infrastructure behind fmmt_debug_umma).
// The fused attention kernel (attn_fused.cu) relies on three operand layouts: K-major SWIZZLE_128B (as the GEMM),
// K-major SWIZZLE_64B (32-channel k-blocks: 64-byte rows) and MN-major SWIZZLE_64B (V as [key][head_dim]). This kernel runs
// ONE accumulator group D[128 x N] = sum over k-steps of A_k * B_k with caller-supplied shared-memory images and descriptor
// fields, so that tests/test_umma_layouts_gpu.py can pin each layout against a plain matmul.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace fmmt {

namespace {

constexpr int PROBE_SMEM = 160 * 1024;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ a_img, int a_bytes, const uint8_t* __restrict__ b_img, int b_bytes,
                  unsigned long long adesc_tpl, unsigned long long bdesc_tpl, unsigned int a_off, unsigned int b_off,
                  unsigned int idesc, int ksteps, int a_step, int b_step, int ncols, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = (static_cast<uint32_t>(a_bytes) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < a_bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(a_img)[i];
  for (int i = threadIdx.x; i < b_bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem + b_base)[i] = reinterpret_cast<const uint32_t*>(b_img)[i];
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) {
    const uint64_t a0 = adesc_tpl | static_cast<uint64_t>(((smem_base + a_off) & 0x3FFFF) >> 4);
    const uint64_t b0 = bdesc_tpl | static_cast<uint64_t>(((smem_base + b_base + b_off) & 0x3FFFF) >> 4);
    for (int k = 0; k < ksteps; ++k)
      umma_bf16(tmem_base, a0 + static_cast<uint64_t>(k) * a_step, b0 + static_cast<uint64_t>(k) * b_step, idesc, k != 0 ? 1u : 0u);
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0, 60);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  for (int c = 0; c < ncols; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c), v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c + j < ncols; ++j) out[static_cast<size_t>(row) * ncols + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

// Host entry (synchronous). Images are device pointers.
int umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, unsigned long long adesc_tpl,
               unsigned long long bdesc_tpl, unsigned int a_off, unsigned int b_off, unsigned int idesc, int ksteps,
               int a_step, int b_step, int ncols, float* out) {
  if (a_bytes <= 0 || b_bytes <= 0 || (a_bytes & 3) || (b_bytes & 3) || a_bytes + b_bytes + 3072 > PROBE_SMEM || ncols <= 0 ||
      ncols > 512)
    return -1;
  if (cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM) != cudaSuccess) return -2;
  umma_probe_kernel<<<1, 128, PROBE_SMEM>>>(static_cast<const uint8_t*>(a_img), a_bytes, static_cast<const uint8_t*>(b_img),
                                            b_bytes, adesc_tpl, bdesc_tpl, a_off, b_off, idesc, ksteps, a_step, b_step, ncols,
                                            out);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -3;
}

}  // namespace fmmt
