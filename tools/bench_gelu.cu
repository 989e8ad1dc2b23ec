// Micro-benchmark (B200): issue cost of the erf-GELU epilogue math, scalar fp32 vs packed f32x2 (FFMA2).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_gelu tools/bench_gelu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float gelu1(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(z, -2.98332428e-03f, 2.97336457e-02f);
  p = fmaf(z, p, -1.48837507e-01f);
  p = fmaf(z, p, -9.18433869e-01f);
  p = fmaf(z, p, -1.62789775e+00f);
  p = fmaf(z, p, -1.0f - 2.71726947e-07f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  return fmaf(-fabsf(x), e, fmaxf(x, 0.0f));
}
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void gelu2(float& a, float& b) {
  const float na = __uint_as_float(__float_as_uint(a) | 0x80000000u), nb = __uint_as_float(__float_as_uint(b) | 0x80000000u);
  const uint64_t nx = pk(na, nb);
  const uint64_t z = mul2(nx, pk(-0.70710678118654752f, -0.70710678118654752f));
  uint64_t p = fma2(z, pk(-2.98332428e-03f, -2.98332428e-03f), pk(2.97336457e-02f, 2.97336457e-02f));
  p = fma2(z, p, pk(-1.48837507e-01f, -1.48837507e-01f));
  p = fma2(z, p, pk(-9.18433869e-01f, -9.18433869e-01f));
  p = fma2(z, p, pk(-1.62789775e+00f, -1.62789775e+00f));
  p = fma2(z, p, pk(-1.0f - 2.71726947e-07f, -1.0f - 2.71726947e-07f));
  float pa, pb, ea, eb;
  upk(p, pa, pb);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(pa));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(pb));
  const uint64_t r = fma2(nx, pk(ea, eb), pk(fmaxf(a, 0.f), fmaxf(b, 0.f)));
  upk(r, a, b);
}

template <int MODE>
__global__ void __launch_bounds__(512) k(const float* in, uint32_t* out, int iters) {
  float v[32];
  for (int j = 0; j < 32; ++j) v[j] = in[(threadIdx.x + j * 37) & 1023];
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a = v[j] + 0.001f * it, b = v[j + 1] - 0.001f * it;
      if (MODE == 0) { a = gelu1(a); b = gelu1(b); } else { gelu2(a, b); }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      acc ^= *reinterpret_cast<uint32_t*>(&h);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  float* in; uint32_t* out;
  cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 512 * 4);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (i % 97) * 0.1f - 4.8f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, 512>>>(in, out, iters); else k<1><<<148, 512>>>(in, out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double elems = 148.0 * 512 * 32 * iters;
    printf("%s: %.3f ms, %.1f Gelem/s, %.2f SM-cycles per warp-element-row (1.965 GHz): issue slots/element = %.2f\n",
           mode == 0 ? "scalar" : "packed f32x2", ms, elems / ms / 1e6, 0.0, ms * 1e-3 * 1.965e9 * 4 / (512 / 32 * 32.0 * iters));
    uint32_t o[4]; cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost); printf("  check %08x\n", o[0]);
  }
  return 0;
}
