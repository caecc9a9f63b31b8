// Micro-benchmark: issue rate of FFMA (one fp32 FMA per lane) against FFMA2 (fma.rn.f32x2, two per lane) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b),
                     uc = *reinterpret_cast<unsigned long long*>(&c), ud;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  return *reinterpret_cast<float2*>(&ud);
}

template <bool PACKED>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) acc[i] = ffma2(acc[i], av, bv);
      else { acc[i].x = fmaf(acc[i].x, av.x, bv.x); acc[i].y = fmaf(acc[i].y, av.y, bv.y); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  const int grid = 148 * 8, iters = 20000;
  cudaMalloc(&out, grid * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int packed = 0; packed < 2; ++packed) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (packed) k<true><<<grid, 256>>>(out, 0.999f, 0.001f, iters); else k<false><<<grid, 256>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)grid * 256 * iters * 16;
      if (rep) printf("%s: %.3f ms, %.1f TFLOP/s fp32 (%.1f FMA/clk/SM at 1.9 GHz)\n", packed ? "FFMA2" : "FFMA ", ms,
                      2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
