// fp32 CUDA-core GEMMs for the strict-parity (GLOWK_F32) mode of the coupling network.
// Same operand layout and epilogues as the tcgen05 path (gemm_sm100.cu); 64x64x16 tiles,
// 256 threads, 4x4 register micro-tiles.
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace glowk {

constexpr int BM = 64, BN = 64, BK = 16;

// out[M][N] = epi(A[M][K] . B[N][K]^T)
template <int EPI, typename OutT>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                int64_t M, int N, int K, EpiParams ep, OutT* __restrict__ out, int64_t ldo) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ float red_a[16][BN];
  __shared__ float red_b[16][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (m0 + lrow < M && k0 + lk < K) a = *reinterpret_cast<const float4*>(A + (m0 + lrow) * lda + k0 + lk);
    if (n0 + lrow < N && k0 + lk < K) b = *reinterpret_cast<const float4*>(B + (int64_t)(n0 + lrow) * ldb + k0 + lk);
    As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
    Bs[lk + 0][lrow] = b.x; Bs[lk + 1][lrow] = b.y; Bs[lk + 2][lrow] = b.z; Bs[lk + 3][lrow] = b.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue: 4 rows x 4 consecutive columns per thread
  float csum_a[4] = {0.f, 0.f, 0.f, 0.f}, csum_b[4] = {0.f, 0.f, 0.f, 0.f};
  const int nc = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    epilogue_apply<EPI, 4>(ep, m, nc, N, v, csum_a, csum_b);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (nc + j < N) out[m * ldo + nc + j] = from_f32<OutT>(v[j]);
  }
  if (EPI == GLOWK_EPI_RELU_BWD) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { red_a[ty][tx * 4 + j] = csum_a[j]; red_b[ty][tx * 4 + j] = csum_b[j]; }
    __syncthreads();
    if (tid < BN && n0 + tid < N) {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) { sa += red_a[r][tid]; sb += red_b[r][tid]; }
      epilogue_commit_colsums(ep, n0 + tid, sa, sb);
    }
  }
}

// dW[Mo][No] += A[P][Mo]^T . B[P][No]   (reduction over pixels, split over grid.z)
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, int64_t P,
                  int Mo, int No, float* __restrict__ dW, int64_t lddw, int64_t chunk) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int64_t p_begin = (int64_t)blockIdx.z * chunk;
  const int64_t p_end = p_begin + chunk < P ? p_begin + chunk : P;
  const int lr = tid >> 4, lc = (tid & 15) * 4;  // 16 rows x 64 cols per pass
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    const int64_t p = p0 + lr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ma = m0 + lc + j, nb = n0 + lc + j;
      As[lr][lc + j] = (p < p_end && ma < Mo) ? to_f32<T>(A[p * lda + ma]) : 0.f;
      Bs[lr][lc + j] = (p < p_end && nb < No) ? to_f32<T>(B[p * ldb + nb]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < Mo && n < No) atomicAdd(dW + (int64_t)m * lddw + n, acc[i][j]);
    }
}

template <int EPI>
static int launch_gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N,
                           int64_t K, const EpiParams& ep, void* out, int out_dtype, int64_t ldo,
                           cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
  if (out_dtype == GLOWK_F32)
    gemm_f32_kernel<EPI, float><<<grid, 256, 0, st>>>(A, lda, B, ldb, M, (int)N, (int)K, ep, (float*)out, ldo);
  else
    gemm_f32_kernel<EPI, __nv_bfloat16><<<grid, 256, 0, st>>>(A, lda, B, ldb, M, (int)N, (int)K, ep, (__nv_bfloat16*)out, ldo);
  GLOWK_CHECK_LAUNCH("glowk_gemm(f32)");
  return GLOWK_OK;
}

int gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
             int epilogue, const EpiParams& ep, void* out, int out_dtype, int64_t ldo, cudaStream_t st) {
  GLOWK_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && K % 4 == 0, "glowk_gemm(f32): lda, ldb, K must be multiples of 4");
  GLOWK_CHECK_ARG(((uintptr_t)A | (uintptr_t)B) % 16 == 0, "glowk_gemm(f32): operands must be 16-byte aligned");
  GLOWK_CHECK_ARG(ceil_div(M, BM) <= 65535, "glowk_gemm(f32): M too large for one launch");
  switch (epilogue) {
    case GLOWK_EPI_STORE: return launch_gemm_f32<GLOWK_EPI_STORE>(A, lda, B, ldb, M, N, K, ep, out, out_dtype, ldo, st);
    case GLOWK_EPI_ACTNORM_RELU: return launch_gemm_f32<GLOWK_EPI_ACTNORM_RELU>(A, lda, B, ldb, M, N, K, ep, out, out_dtype, ldo, st);
    case GLOWK_EPI_ACTNORM: return launch_gemm_f32<GLOWK_EPI_ACTNORM>(A, lda, B, ldb, M, N, K, ep, out, out_dtype, ldo, st);
    case GLOWK_EPI_ZEROS: return launch_gemm_f32<GLOWK_EPI_ZEROS>(A, lda, B, ldb, M, N, K, ep, out, out_dtype, ldo, st);
    case GLOWK_EPI_RELU_BWD: return launch_gemm_f32<GLOWK_EPI_RELU_BWD>(A, lda, B, ldb, M, N, K, ep, out, out_dtype, ldo, st);
  }
  return fail(GLOWK_EINVAL, "glowk_gemm: unknown epilogue %d", epilogue);
}

int wgrad_simt(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype, int64_t P, int64_t Mo,
               int64_t No, float* dW, int64_t lddw, cudaStream_t st) {
  const int64_t tiles = ceil_div(Mo, BM) * ceil_div(No, BN);
  int64_t split = ceil_div((int64_t)4 * sm_count(), tiles);
  int64_t chunk = ceil_div(ceil_div(P, split), BK) * BK;
  if (chunk < 256) chunk = 256;
  split = ceil_div(P, chunk);
  GLOWK_CHECK_ARG(split <= 65535, "glowk_gemm_wgrad: too many pixel chunks");
  dim3 grid((unsigned)ceil_div(No, BN), (unsigned)ceil_div(Mo, BM), (unsigned)split);
  if (act_dtype == GLOWK_F32)
    wgrad_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)A, lda, (const float*)B, ldb, P, (int)Mo, (int)No, dW, lddw, chunk);
  else
    wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)B, ldb, P, (int)Mo, (int)No, dW, lddw, chunk);
  GLOWK_CHECK_LAUNCH("glowk_gemm_wgrad(simt)");
  return GLOWK_OK;
}

}  // namespace glowk
