// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, of the operand source of A
// (shared memory / tensor memory) and of how many accumulators a dependent chain of MMAs is spread over.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../pytorch_glow_b200/csrc -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace glowk::tc;

namespace glowk {   // symbols tc_common.cuh expects from api.cu
static char g_err[512];
char* last_error_buf() { return g_err; }
int fail(int code, const char*, ...) { return code; }
}

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS, 1: TS.  nacc accumulators of n columns each, used round-robin; `group` consecutive MMAs go to the same
// accumulator before moving on.
__global__ void __launch_bounds__(128, 1) k(long long* out, int n, int nacc, int group, int mode, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n, 0, 0);
    const uint64_t adesc = make_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t bdesc = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
    const int acc0 = mode ? 256 : 0;          // TS: A lives in columns [0,256)
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int a = (i / group) % nacc;
      const uint32_t d = tb + (uint32_t)(acc0 + a * n);
      const int k = i & 3;
      if (mode == 0) tcgen05_mma_bf16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, i >= nacc * group);
      else mma_ts(d, tb + (uint32_t)((i & 31) * 8), bdesc + (uint64_t)(k * 2), idesc, i >= nacc * group);
    }
    tcgen05_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
  const int iters = 4096;
  printf("%-4s %-5s %-5s %-6s %-8s %s\n", "mode", "N", "nacc", "group", "cyc/MMA", "floor(N/2)");
  for (int mode = 0; mode < 2; ++mode)
    for (int n : {32, 64, 128, 256})
      for (int nacc : {1, 2, 4})
        for (int group : {1, 4, 32}) {
          if (nacc * n > 256) continue;
          if (nacc == 1 && group != 1) continue;
          long long c = 0;
          for (int rep = 0; rep < 2; ++rep) {
            k<<<148, 128, 49152, 0>>>(d, n, nacc, group, mode, iters);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          }
          printf("%-4s %-5d %-5d %-6d %-8.1f %d\n", mode ? "TS" : "SS", n, nacc, group, (double)c / iters, n / 2);
        }
  return 0;
}
