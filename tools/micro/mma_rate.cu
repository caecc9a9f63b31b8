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

// MODE 0: SS, 1: TS.  The issuing warp runs warp-uniform (elect.sync); one loop iteration = one "stage" of G MMAs
// (compile-time), optionally followed by a tcgen05.commit (COMMIT) and preceded by a try_wait on an mbarrier whose
// phase completed long ago (WAIT) -- the per-stage protocol of a TMA-fed pipeline without the TMA.
template <int MODE, int G, int COMMIT, int WAIT>
__global__ void __launch_bounds__(128, 1) k(long long* out, int n, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy, done;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&dummy, 1 << 20); mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive(&done);                      // phase 0 of `done` is complete: waiting for parity 0 succeeds at once
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(n, 0, 0);
    const uint64_t adesc = make_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t bdesc = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
    const uint32_t acc0 = tb + (MODE ? 256u : 0u);          // TS: A lives in columns [0,256)
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += G) {
      if (WAIT) { mbar_wait(&done, 0); tcgen05_fence_after(); }
      if (elect_one_sync()) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
          if (MODE == 0) tcgen05_mma_bf16(acc0, adesc + (uint64_t)((j & 3) * 2), bdesc + (uint64_t)((j & 3) * 2), idesc, 1u);
          else mma_ts(acc0, tb + (uint32_t)((j & 31) * 8), bdesc + (uint64_t)((j & 3) * 2), idesc, 1u);
        }
        if (COMMIT) tcgen05_commit(&dummy);
      }
      __syncwarp();
    }
    if (elect_one_sync()) tcgen05_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

template <int MODE, int G, int COMMIT, int WAIT>
static double run(long long* d, int n, int iters) {
  auto kern = k<MODE, G, COMMIT, WAIT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
  long long c = 0;
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<148, 128, 49152, 0>>>(d, n, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  }
  return (double)c / iters;
}

template <int MODE, int G>
static void row(long long* d, int n, int iters) {
  printf("%-4s %-5d %-3d  plain %6.1f   +commit %6.1f   +wait %6.1f   +both %6.1f   (floor %d)\n", MODE ? "TS" : "SS", n, G,
         run<MODE, G, 0, 0>(d, n, iters), run<MODE, G, 1, 0>(d, n, iters), run<MODE, G, 0, 1>(d, n, iters),
         run<MODE, G, 1, 1>(d, n, iters), n / 2);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 8192;
  printf("cycles per MMA (M=128, K=16); G = MMAs per stage\n");
  for (int n : {16, 32, 64, 96, 112, 128, 192, 256}) {
    printf("%-4s %-5d rate %6.1f   ", "SS", n, run<0, 8, 0, 0>(d, n, iters));
    printf("%-4s rate %6.1f  (floor %d)\n", "TS", run<1, 8, 0, 0>(d, n, iters), n / 2);
  }
  printf("\nper-stage protocol cost\n");
  for (int n : {64, 128, 256}) {
    row<1, 4>(d, n, iters); row<1, 8>(d, n, iters); row<1, 16>(d, n, iters); row<1, 32>(d, n, iters);
    row<0, 4>(d, n, iters); row<0, 8>(d, n, iters); row<0, 16>(d, n, iters);
  }
  return 0;
}
