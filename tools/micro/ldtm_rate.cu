// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps (1 CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../pytorch_glow_b200/csrc -o ldtm_rate ldtm_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace glowk::tc;

namespace glowk {
static char g_err[512];
char* last_error_buf() { return g_err; }
int fail(int code, const char*, ...) { return code; }
}

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// MODE 0: ld x32, wait after every load; 1: ld x32, 4 loads in flight per wait; 2: st x32 + wait::st
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
  uint32_t acc = 0;
  uint32_t r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = threadIdx.x + j;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      tmem_ld32_async(tb + (uint32_t)((i & 3) * 32), r);
      tmem_ld_wait();
  #pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= r[j];
    } else if (MODE == 1) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld32_async(tb, a); tmem_ld32_async(tb + 32, b); tmem_ld32_async(tb + 64, c); tmem_ld32_async(tb + 96, d);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j] ^ b[j] ^ c[j] ^ d[j];
    } else {
      st32(tb + (uint32_t)((i & 3) * 32), r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
  if (acc == 0x12345678u) out[1] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
  }
}

template <int MODE>
static void run(long long* d, int warps, const char* name) {
  const int iters = 1024;
  long long c = 0;
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE><<<148, warps * 32>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  }
  const double bytes = (double)warps * iters * 4096.0 * (MODE == 1 ? 4 : 1);
  printf("%-34s warps %2d : %7.1f B/clk/SM  (%6.1f cycles per 4 KB warp-load)\n", name, warps, bytes / (double)c,
         (double)c / (iters * (MODE == 1 ? 4 : 1)));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int w : {1, 4, 8, 16}) run<0>(d, w, "tcgen05.ld 32x32b.x32, wait each");
  for (int w : {1, 4, 8, 16}) run<1>(d, w, "tcgen05.ld 32x32b.x32, 4 in flight");
  for (int w : {1, 4, 8, 16}) run<2>(d, w, "tcgen05.st 32x32b.x32, wait each");
  return 0;
}
