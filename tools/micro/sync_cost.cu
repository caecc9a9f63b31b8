// Micro-benchmark: issue cost (cycles seen by the issuing warp) of the synchronisation primitives a tcgen05 pipeline
// executes per stage, measured in a warp-uniform loop on one warp per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../pytorch_glow_b200/csrc -o sync_cost sync_cost.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace glowk::tc;

namespace glowk {
static char g_err[512];
char* last_error_buf() { return g_err; }
int fail(int code, const char*, ...) { return code; }
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int WHAT>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  __shared__ uint64_t done, dummy;
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    mbar_init(&dummy, 1 << 20); mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive(&done);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (WHAT == 0) { if (elect_one_sync()) sink += 1; __syncwarp(); }
      if (WHAT == 1) { if (elect_one_sync()) tcgen05_commit(&dummy); __syncwarp(); }
      if (WHAT == 2) { tcgen05_fence_after(); }
      if (WHAT == 3) { mbar_wait(&done, 0); }
      if (WHAT == 4) { sink += mbar_test(&done, 0); }
      if (WHAT == 5) { if (threadIdx.x == 0) mbar_arrive(&dummy); __syncwarp(); }
      if (WHAT == 6) { fence_proxy_async(); }
      if (WHAT == 7) { __syncwarp(); }
      if (WHAT == 8) { mbar_wait(&done, 0); tcgen05_fence_after(); if (elect_one_sync()) tcgen05_commit(&dummy); __syncwarp(); }
      if (WHAT == 9) { tcgen05_fence_before(); }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = sink; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32) : "memory");
  }
}

template <int WHAT>
static void run(long long* d, const char* name) {
  const int iters = 2000;
  long long c[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    k<WHAT><<<148, 128>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
  }
  printf("%-58s %7.1f cycles\n", name, (double)c[0] / iters);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<7>(d, "__syncwarp");
  run<0>(d, "elect.sync + branch + __syncwarp");
  run<1>(d, "elect + tcgen05.commit + __syncwarp");
  run<2>(d, "tcgen05.fence::after_thread_sync");
  run<9>(d, "tcgen05.fence::before_thread_sync");
  run<3>(d, "mbarrier.try_wait loop on a completed phase (32 lanes)");
  run<4>(d, "mbarrier.test_wait on a completed phase, result consumed");
  run<5>(d, "lane 0 mbarrier.arrive + __syncwarp");
  run<6>(d, "fence.proxy.async");
  run<8>(d, "try_wait + fence::after + elect + commit + syncwarp");
  return 0;
}
