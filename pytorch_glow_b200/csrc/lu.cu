// Small dense-matrix kernels for the invertible 1x1 convolution weight (one CTA each):
// log|det W| and W^-1 by LU with partial pivoting (replaces torch.det / Tensor.inverse,
// network/module.py:357,365), and the LU parameterisation W = P L (U + diag(s)).
#include "common.cuh"

namespace glowk {

// A: [C][C] fp64 workspace (shared or global), perm: [C] ints.
__device__ void lu_factor_inplace(double* A, int* perm, int C, double* logabs_out) {
  __shared__ int s_piv;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < C; i += nthr) perm[i] = i;
  __syncthreads();
  double logabs = 0.0;
  for (int k = 0; k < C; ++k) {
    if (tid < 32) {  // warp 0: arg max |A[i][k]|, i >= k
      double best = -1.0;
      int bi = k;
      for (int i = k + tid; i < C; i += 32) {
        const double v = fabs(A[i * C + k]);
        if (v > best) { best = v; bi = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) s_piv = bi;
    }
    __syncthreads();
    const int piv = s_piv;
    if (piv != k) {
      for (int j = tid; j < C; j += nthr) {
        const double t = A[k * C + j]; A[k * C + j] = A[piv * C + j]; A[piv * C + j] = t;
      }
      if (tid == 0) { const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
    }
    __syncthreads();
    const double d = A[k * C + k];
    logabs += log(fabs(d));
    for (int i = k + 1 + tid; i < C; i += nthr) A[i * C + k] /= d;
    __syncthreads();
    const int rem = C - k - 1;
    for (int e = tid; e < rem * rem; e += nthr) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      A[i * C + j] -= A[i * C + k] * A[k * C + j];
    }
    __syncthreads();
  }
  *logabs_out = logabs;
}

// grid.x = number of matrices (one CTA each): the 96 invconv weights of a K=32, L=3 model are factorised
// by three launches (one per channel count) instead of 96 host-synchronising cuSOLVER calls.
__global__ void invconv_prepare_kernel(const float* __restrict__ w, int C, float* __restrict__ logabsdet_out,
                                       float* __restrict__ winv_out, double* gwork) {
  extern __shared__ __align__(16) unsigned char lu_smem[];
  double* A;
  double* X;
  int* perm;
  const size_t mat = blockIdx.x;
  w += mat * C * C;
  logabsdet_out += mat;
  if (winv_out) winv_out += mat * C * C;
  if (gwork) {
    gwork += mat * 2 * (size_t)C * C;
    A = gwork; X = gwork + (size_t)C * C; perm = reinterpret_cast<int*>(lu_smem);
  } else {
    A = reinterpret_cast<double*>(lu_smem);
    X = A + (size_t)C * C;
    perm = reinterpret_cast<int*>(X + (winv_out ? (size_t)C * C : 0));
  }
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int e = tid; e < C * C; e += nthr) A[e] = (double)w[e];
  __syncthreads();
  double logabs;
  lu_factor_inplace(A, perm, C, &logabs);
  if (tid == 0) logabsdet_out[0] = (float)logabs;
  if (!winv_out) return;
  // column j of W^-1: solve L U x = P e_j
  for (int j = tid; j < C; j += nthr) {
    for (int i = 0; i < C; ++i) {
      double s = (perm[i] == j) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= A[i * C + k] * X[k * C + j];
      X[i * C + j] = s;
    }
    for (int i = C - 1; i >= 0; --i) {
      double s = X[i * C + j];
      for (int k = i + 1; k < C; ++k) s -= A[i * C + k] * X[k * C + j];
      s /= A[i * C + i];
      X[i * C + j] = s;
      winv_out[i * C + j] = (float)s;
    }
  }
}

__global__ void lu_assemble_kernel(const float* __restrict__ p, const float* __restrict__ l,
                                   const float* __restrict__ u, const float* __restrict__ sign_s,
                                   const float* __restrict__ log_s, int C, float* __restrict__ w_out,
                                   float* __restrict__ winv_out, float* __restrict__ logabsdet_out,
                                   double* X) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  // W[i][j] = M[pi(i)][j], M = Lf . Uf, pi(i) = column of the 1 in row i of P
  for (int e = tid; e < C * C; e += nthr) {
    const int i = e / C, j = e - i * C;
    int r = 0;
    for (int k = 0; k < C; ++k) if (p[i * C + k] != 0.f) r = k;
    double s = 0.0;
    const int mmax = r < j ? r : j;  // Lf[r][m]!=0 for m<=r ; Uf[m][j]!=0 for m<=j
    for (int m = 0; m <= mmax; ++m) {
      const double lv = (m == r) ? 1.0 : (double)l[r * C + m];
      const double uv = (m == j) ? (double)sign_s[j] * exp((double)log_s[j]) : (double)u[m * C + j];
      s += lv * uv;
    }
    w_out[e] = (float)s;
  }
  if (tid == 0) {
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += (double)log_s[c];
    logabsdet_out[0] = (float)s;
  }
  if (!winv_out) return;
  // W^-1 = Uf^-1 Lf^-1 P^T : column j solves Lf y = P^T e_j, Uf x = y  (two triangular solves)
  for (int j = tid; j < C; j += nthr) {
    for (int i = 0; i < C; ++i) {
      double s = (double)p[j * C + i];
      for (int k = 0; k < i; ++k) s -= (double)l[i * C + k] * X[k * C + j];
      X[i * C + j] = s;
    }
    for (int i = C - 1; i >= 0; --i) {
      double s = X[i * C + j];
      for (int k = i + 1; k < C; ++k) s -= (double)u[i * C + k] * X[k * C + j];
      s /= (double)sign_s[i] * exp((double)log_s[i]);
      X[i * C + j] = s;
      winv_out[i * C + j] = (float)s;
    }
  }
}

}  // namespace glowk

using namespace glowk;

// Workspace policy: fp64 scratch lives in dynamic shared memory when it fits (C <= 96 with the
// inverse, C <= 160 without); otherwise in a per-call device allocation from the stream-ordered
// pool (cudaMallocAsync), which neither synchronises nor keeps global state.
extern "C" int glowk_invconv_prepare(const float* w, int64_t C, float* logabsdet_out, float* winv_out, void* stream) {
  return glowk_invconv_prepare_batched(w, 1, C, logabsdet_out, winv_out, stream);
}

extern "C" int glowk_invconv_prepare_batched(const float* w, int64_t batch, int64_t C, float* logabsdet_out,
                                             float* winv_out, void* stream) {
  if (batch == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(w && logabsdet_out, "glowk_invconv_prepare: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C <= 1024, "glowk_invconv_prepare: C=%lld out of range", (long long)C);
  GLOWK_CHECK_ARG(batch > 0 && batch <= 65535, "glowk_invconv_prepare: batch=%lld out of range", (long long)batch);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t mats = winv_out ? 2 : 1;
  size_t smem = mats * (size_t)C * C * sizeof(double) + (size_t)C * sizeof(int);
  double* gwork = nullptr;
  if (smem > 200 * 1024) {
    GLOWK_CUDA(cudaMallocAsync((void**)&gwork, (size_t)batch * 2 * (size_t)C * C * sizeof(double), st));
    smem = (size_t)C * sizeof(int);
  }
  if (smem > 48 * 1024)
    GLOWK_CUDA(cudaFuncSetAttribute(invconv_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = C <= 16 ? 64 : (C <= 48 ? 128 : 256);
  invconv_prepare_kernel<<<(unsigned)batch, threads, smem, st>>>(w, (int)C, logabsdet_out, winv_out, gwork);
  cudaError_t e = cudaGetLastError();
  if (gwork) cudaFreeAsync(gwork, st);
  if (e != cudaSuccess) return fail(GLOWK_ECUDA, "glowk_invconv_prepare: %s", cudaGetErrorString(e));
  return GLOWK_OK;
}

extern "C" int glowk_invconv_lu_assemble(const float* p, const float* l, const float* u, const float* sign_s,
                                         const float* log_s, int64_t C, float* w_out, float* winv_out,
                                         float* logabsdet_out, void* stream) {
  GLOWK_CHECK_ARG(p && l && u && sign_s && log_s && w_out && logabsdet_out, "glowk_invconv_lu_assemble: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C <= 1024, "glowk_invconv_lu_assemble: C out of range");
  cudaStream_t st = (cudaStream_t)stream;
  double* X = nullptr;
  if (winv_out) GLOWK_CUDA(cudaMallocAsync((void**)&X, (size_t)C * C * sizeof(double), st));
  const int threads = C <= 16 ? 64 : (C <= 48 ? 128 : 256);
  lu_assemble_kernel<<<1, threads, 0, st>>>(p, l, u, sign_s, log_s, (int)C, w_out, winv_out, logabsdet_out, X);
  cudaError_t e = cudaGetLastError();
  if (X) cudaFreeAsync(X, st);
  if (e != cudaSuccess) return fail(GLOWK_ECUDA, "glowk_invconv_lu_assemble: %s", cudaGetErrorString(e));
  return GLOWK_OK;
}
