// Shared helpers for the glowk kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/glowk.h"

namespace glowk {

// Thread-local error message (no global mutable state, re-entrant).
char* last_error_buf();
int fail(int code, const char* fmt, ...);

#define GLOWK_CHECK_ARG(cond, ...)                            \
  do {                                                        \
    if (!(cond)) return ::glowk::fail(GLOWK_EINVAL, __VA_ARGS__); \
  } while (0)

#define GLOWK_CHECK_LAUNCH(name)                                                        \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess)                                                             \
      return ::glowk::fail(GLOWK_ECUDA, "%s: %s", name, cudaGetErrorString(e__));       \
  } while (0)

#define GLOWK_CUDA(call)                                                                \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return ::glowk::fail(GLOWK_ECUDA, "%s: %s", #call, cudaGetErrorString(e__));      \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached per device id, read-only after first query).
int sm_count();
// CTAs of a kernel resident on the whole device (occupancy x SMs), cached; see api.cu.
int resident_ctas(const void* kern, int threads, size_t smem);

// Programmatic dependent launch (PDL).  The per-step kernels are short (a K=32, L=3 Glow launches ~1500 of them
// per training step, many of a few microseconds), so each one is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: it may be scheduled while its predecessor drains, does its
// memory-free set-up (shared-memory carve-up, mbarrier init, TMEM allocation, tensor-map prefetch) and then blocks
// in pdl_wait() until every earlier grid has completed and flushed.  Rules kept by every kernel that uses it:
// NO global-memory access before pdl_wait(); kernels that write parameters / packed weights never trigger early.
// The attribute is OFF unless GLOWK_PDL=1 is set in the environment (read once): see pdl_enabled() in api.cu.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// The persistent tcgen05 GEMMs (one CTA per SM for the whole kernel) can release their dependents either at entry
// (default) or -- build with -DGLOWK_PDL_LATE -- once their TMA producer has issued the loads of its last tile.
#ifdef GLOWK_PDL_LATE
__device__ __forceinline__ void pdl_trigger_entry() {}
__device__ __forceinline__ void pdl_trigger_drain() { pdl_trigger(); }
#else
__device__ __forceinline__ void pdl_trigger_entry() { pdl_trigger(); }
__device__ __forceinline__ void pdl_trigger_drain() {}
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Division by a launch-constant divisor without the ~20-instruction integer-divide sequence (the flow kernels
// move a few bytes per thread, so their index arithmetic is a first-order cost): q = umulhi(n, mul) >> shr,
// exact for 0 <= n < 2^31 (Granlund-Montgomery round-up method).
struct FastDiv {
  uint32_t d, mul, shr;
};
static inline FastDiv make_fastdiv(int64_t d) {
  FastDiv f;
  f.d = (uint32_t)d; f.mul = 0; f.shr = 0;
  if (d > 1) {
    int lg = 0;
    while ((1ll << lg) < d) ++lg;                      // ceil(log2 d)
    const int p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + (uint64_t)d - 1) / (uint64_t)d);
    f.shr = (uint32_t)(p - 32);
  }
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast to all if `bcast`).  `red` >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) r = warp_sum(r);
  return r;  // valid in warp 0
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Streaming (read-once) 128-bit load / store: keep L1 for reused data.
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace glowk
