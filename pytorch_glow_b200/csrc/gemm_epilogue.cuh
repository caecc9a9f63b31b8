// Epilogue functors shared by the fp32 CUDA-core GEMM and the tcgen05 GEMM.
// A call handles NV consecutive output columns n0..n0+NV-1 of one output row m.
#pragma once
#include "common.cuh"

namespace glowk {

struct EpiParams {
  const float* bias;   // [N]
  const float* logs;   // [N]
  float f;             // logscale factor
  const void* y;       // RELU_BWD: saved forward output [M][ldy]
  int64_t ldy;
  int y_bf16;
  float* dlogs;        // RELU_BWD: [N] accumulators
  float* dbias;
};

template <int EPI, int NV>
__device__ __forceinline__ void epilogue_apply(const EpiParams& ep, int64_t m, int n0, int N, float (&v)[NV],
                                               float (&csum_a)[NV], float (&csum_b)[NV]) {
  if (EPI == GLOWK_EPI_STORE) return;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int n = n0 + j;
    if (n >= N) continue;
    if (EPI == GLOWK_EPI_ACTNORM_RELU || EPI == GLOWK_EPI_ACTNORM || EPI == GLOWK_EPI_ZEROS) {
      // Conv2d -> ActNorm: (x + bias) * exp(logs*f)  (module.py:48,73);  Conv2dZeros: (conv + bias) * exp(logs*f)
      float r = (v[j] + ep.bias[n]) * expf(ep.logs[n] * ep.f);
      if (EPI == GLOWK_EPI_ACTNORM_RELU) r = fmaxf(r, 0.f);
      v[j] = r;
    } else if (EPI == GLOWK_EPI_RELU_BWD) {
      const float yv = ep.y_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.y)[m * ep.ldy + n])
                                 : reinterpret_cast<const float*>(ep.y)[m * ep.ldy + n];
      const float g = yv > 0.f ? v[j] : 0.f;
      csum_a[j] += g * yv;
      csum_b[j] += g;
      v[j] = g * expf(ep.logs[n] * ep.f);
    }
  }
}

__device__ __forceinline__ void epilogue_commit_colsums(const EpiParams& ep, int n, float sum_gy, float sum_g) {
  atomicAdd(ep.dlogs + n, ep.f * sum_gy);
  atomicAdd(ep.dbias + n, expf(ep.logs[n] * ep.f) * sum_g);
}

// implemented in gemm_simt.cu
int gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
             int epilogue, const EpiParams& ep, void* out, int out_dtype, int64_t ldo, cudaStream_t st);
int wgrad_simt(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype, int64_t P, int64_t Mo,
               int64_t No, float* dW, int64_t lddw, cudaStream_t st);
// implemented in gemm_sm100.cu
int gemm_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                 int epilogue, const EpiParams& ep, void* out, int out_dtype, int64_t ldo, int cluster_m, int cluster_n,
                 cudaStream_t st);
int wgrad_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t P, int64_t Mo, int64_t No,
                  float* dW, int64_t lddw, cudaStream_t st);
bool tc_available();
int gemm_debug_trace(unsigned long long* out16);

}  // namespace glowk
