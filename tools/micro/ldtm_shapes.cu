// Micro-benchmark: tcgen05.ld throughput per SM for the different data-path shapes (4 KB per warp instruction each).
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace glowk::tc;
namespace glowk { static char g_err[512]; char* last_error_buf() { return g_err; } int fail(int code, const char*, ...) { return code; } }

#define R32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
#define O32(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])

template <int SHAPE>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]) {
  if (SHAPE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " R32 ", [%32];" : O32(r) : "r"(taddr) : "memory");
  if (SHAPE == 1) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " R32 ", [%32];" : O32(r) : "r"(taddr) : "memory");
  if (SHAPE == 2) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " R32 ", [%32];" : O32(r) : "r"(taddr) : "memory");
  if (SHAPE == 3) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " R32 ", [%32];" : O32(r) : "r"(taddr) : "memory");
}

template <int SHAPE>
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
  const uint32_t tb = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  uint32_t acc = 0;
  uint32_t r[32];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    ld<SHAPE>(tb, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= r[j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) out[1] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
  }
}

template <int SHAPE>
static void run(long long* d, int warps, const char* name) {
  const int iters = 1024;
  long long c = 0;
  for (int rep = 0; rep < 2; ++rep) {
    k<SHAPE><<<148, warps * 32>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: error: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  }
  printf("%-28s warps %2d : %7.1f B/clk/SM\n", name, warps, (double)warps * iters * 4096.0 / (double)c);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int w : {4, 8}) {
    run<0>(d, w, "tcgen05.ld 32x32b.x32");
    run<1>(d, w, "tcgen05.ld 16x256b.x8");
    run<2>(d, w, "tcgen05.ld 16x128b.x16");
    run<3>(d, w, "tcgen05.ld 16x64b.x32");
  }
  return 0;
}
