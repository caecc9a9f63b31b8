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

// ---- large matrices (fp64 workspace in global memory: C > 96 with the inverse, the CelebA-HQ levels 5 / 6) ----
// ncu launch list of the 256^2 L=6 model: the one-CTA kernel above took 7.5 ms per launch (45 ms of a 215 ms step):
// 256 threads per matrix, and one thread per inverse column walking 2 x C^2/2 dependent fp64 FMAs.  Here the
// factorisation runs with 1024 threads per matrix and the inverse is a second kernel with ONE WARP PER COLUMN
// (lanes split each row's dot product; the 8 warps of a CTA walk the same rows of LU, which they share through L1).
// Blocked right-looking LU with partial pivoting (panel width LU_NB): the panel [C-kb][NB] and the row block
// U12 [NB][C-kb-NB] live in shared memory, so every trailing element is loaded / stored once per NB columns (the
// unblocked rank-1 loop touched each one C times through L2: 4.3 ms per launch at C = 384).  Same pivot choice
// (largest |a|, lowest index on ties) as lu_factor_inplace.
constexpr int LU_NB = 16;
__global__ void __launch_bounds__(1024)
invconv_lu_big_kernel(const float* __restrict__ w, int C, float* __restrict__ logabsdet_out, double* gwork, int* gperm) {
  extern __shared__ __align__(16) unsigned char lu_smem[];
  double* Pn = reinterpret_cast<double*>(lu_smem);          // [C][NB]   panel (rows kb.. of columns kb..kb+NB)
  double* U12 = Pn + (size_t)C * LU_NB;                     // [NB][C]   row block right of the panel
  __shared__ int s_piv[LU_NB];
  __shared__ double s_best[32];
  __shared__ int s_bi[32];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const size_t mat = blockIdx.x;
  double* A = gwork + mat * (size_t)C * C;
  int* perm = gperm + mat * (size_t)C;
  w += mat * (size_t)C * C;
  for (int e = tid; e < C * C; e += nthr) A[e] = (double)w[e];
  for (int i = tid; i < C; i += nthr) perm[i] = i;
  __syncthreads();
  double logabs = 0.0;
  for (int kb = 0; kb < C; kb += LU_NB) {
    const int nb = min(LU_NB, C - kb), rows = C - kb;
    // ---- panel -> shared memory
    for (int e = tid; e < rows * nb; e += nthr) {
      const int r = e / nb, c = e - r * nb;
      Pn[r * LU_NB + c] = A[(size_t)(kb + r) * C + kb + c];
    }
    __syncthreads();
    // ---- unblocked LU of the panel (pivots recorded relative to kb)
    for (int k = 0; k < nb; ++k) {
      double best = -1.0; int bi = k;
      for (int r = k + tid; r < rows; r += nthr) {
        const double v = fabs(Pn[r * LU_NB + k]);
        if (v > best) { best = v; bi = r; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { s_best[warp] = best; s_bi[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        best = lane < nwarp ? s_best[lane] : -1.0; bi = lane < nwarp ? s_bi[lane] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) s_piv[k] = bi;
      }
      __syncthreads();
      const int piv = s_piv[k];
      if (piv != k && tid < nb) {
        const double t = Pn[k * LU_NB + tid]; Pn[k * LU_NB + tid] = Pn[piv * LU_NB + tid]; Pn[piv * LU_NB + tid] = t;
      }
      __syncthreads();
      const double d = Pn[k * LU_NB + k];
      if (tid == 0) logabs += log(fabs(d));
      for (int r = k + 1 + tid; r < rows; r += nthr) Pn[r * LU_NB + k] /= d;
      __syncthreads();
      const int rc = nb - k - 1;
      for (int e = tid; e < (rows - k - 1) * rc; e += nthr) {
        const int r = k + 1 + e / rc, c = k + 1 + e % rc;
        Pn[r * LU_NB + c] -= Pn[r * LU_NB + k] * Pn[k * LU_NB + c];
      }
      __syncthreads();
    }
    // ---- apply the panel's row swaps to the columns outside the panel (in order) and to perm; write the panel back
    for (int c = tid; c < C; c += nthr) {
      if (c >= kb && c < kb + nb) continue;
      for (int k = 0; k < nb; ++k) {
        const int piv = s_piv[k];
        if (piv != k) {
          const double t = A[(size_t)(kb + k) * C + c];
          A[(size_t)(kb + k) * C + c] = A[(size_t)(kb + piv) * C + c];
          A[(size_t)(kb + piv) * C + c] = t;
        }
      }
    }
    if (tid == 0)
      for (int k = 0; k < nb; ++k) {
        const int piv = s_piv[k];
        if (piv != k) { const int t = perm[kb + k]; perm[kb + k] = perm[kb + piv]; perm[kb + piv] = t; }
      }
    for (int e = tid; e < rows * nb; e += nthr) {
      const int r = e / nb, c = e - r * nb;
      A[(size_t)(kb + r) * C + kb + c] = Pn[r * LU_NB + c];
    }
    __syncthreads();
    const int rest = C - kb - nb;                           // columns (and rows) right of / below the panel
    if (rest <= 0) break;
    // ---- U12 = L11^-1 A12: one thread per column, forward substitution over the nb panel rows
    for (int c = tid; c < rest; c += nthr) {
      double u[LU_NB];
#pragma unroll
      for (int k = 0; k < LU_NB; ++k) u[k] = k < nb ? A[(size_t)(kb + k) * C + kb + nb + c] : 0.0;
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
#pragma unroll
        for (int m = 0; m < LU_NB; ++m)
          if (m < k) u[k] -= Pn[k * LU_NB + m] * u[m];
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
        if (k < nb) { U12[k * C + c] = u[k]; A[(size_t)(kb + k) * C + kb + nb + c] = u[k]; }
    }
    __syncthreads();
    // ---- trailing update A22 -= L21 U12
    for (int e = tid; e < rest * rest; e += nthr) {
      const int r = e / rest, c = e - r * rest;
      const double* l = Pn + (size_t)(nb + r) * LU_NB;
      double acc = A[(size_t)(kb + nb + r) * C + kb + nb + c];
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
        if (k < nb) acc -= l[k] * U12[k * C + c];
      A[(size_t)(kb + nb + r) * C + kb + nb + c] = acc;
    }
    __syncthreads();
  }
  if (tid == 0) logabsdet_out[mat] = (float)logabs;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
invconv_inverse_big_kernel(const double* __restrict__ gwork, const int* __restrict__ gperm, int C,
                           float* __restrict__ winv_out) {
  extern __shared__ __align__(16) unsigned char lu_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;                       // column of W^-1: solve L U x = P e_j
  if (j >= C) return;
  const size_t mat = blockIdx.y;
  const double* A = gwork + mat * (size_t)C * C;
  const int* perm = gperm + mat * (size_t)C;
  double* x = reinterpret_cast<double*>(lu_smem) + (size_t)warp * C;
  for (int i = 0; i < C; ++i) {                              // L y = P e_j (unit lower triangle)
    const double* row = A + (size_t)i * C;
    double s = 0.0;
    for (int k = lane; k < i; k += 32) s += row[k] * x[k];
    s = warp_sum_f64(s);
    if (lane == 0) x[i] = ((perm[i] == j) ? 1.0 : 0.0) - s;
    __syncwarp();
  }
  for (int i = C - 1; i >= 0; --i) {                         // U x = y
    const double* row = A + (size_t)i * C;
    double s = 0.0;
    for (int k = i + 1 + lane; k < C; k += 32) s += row[k] * x[k];
    s = warp_sum_f64(s);
    if (lane == 0) x[i] = (x[i] - s) / row[i];
    __syncwarp();
  }
  float* out = winv_out + mat * (size_t)C * C;
  for (int i = lane; i < C; i += 32) out[(size_t)i * C + j] = (float)x[i];
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


// Chain rule of the LU parameterisation W = P Lf Uf (Lf = tril(l,-1) + I, Uf = triu(u,1) + diag(sign_s e^{log_s})):
// with G = P^T dW,  dLf = G Uf^T,  dUf = Lf^T G:
//   dl += tril(dLf, -1);  du += triu(dUf, 1);  dlog_s += diag(dUf) * sign_s e^{log_s}.
// One CTA; P^T dW is a row gather (row r of G = row i of dW with P[i][r] = 1).  Replaces five cuBLAS matmuls +
// torch.eye / tril / triu per FlowStep and backward pass on the LU path.
__global__ void lu_grads_kernel(const float* __restrict__ dw, const float* __restrict__ p, const float* __restrict__ l,
                                const float* __restrict__ u, const float* __restrict__ sign_s,
                                const float* __restrict__ log_s, int C, float* __restrict__ dl, float* __restrict__ du,
                                float* __restrict__ dlog_s) {
  __shared__ int rowsrc[1024];                      // (C <= 1024, checked by the entry point; no allocation: graph-safe)
  const int tid = threadIdx.x, nthr = blockDim.x;
  // rowsrc[r] = i with P[i][r] = 1  (G[r][:] = dW[i][:])
  for (int r = tid; r < C; r += nthr) {
    int src = 0;
    for (int i = 0; i < C; ++i) if (p[i * C + r] != 0.f) src = i;
    rowsrc[r] = src;
  }
  __syncthreads();
  for (int e = tid; e < C * C; e += nthr) {
    const int i = e / C, j = e - i * C;
    if (i > j) {
      // dLf[i][j] = sum_k G[i][k] Uf[j][k]   (Uf[j][k] != 0 for k >= j)
      const float* g = dw + (size_t)rowsrc[i] * C;
      double s = 0.0;
      for (int k = j; k < C; ++k) {
        const double uv = (k == j) ? (double)sign_s[j] * exp((double)log_s[j]) : (double)u[j * C + k];
        s += (double)g[k] * uv;
      }
      dl[e] += (float)s;
    } else {
      // dUf[i][j] = sum_k Lf[k][i] G[k][j]   (Lf[k][i] != 0 for k >= i)
      double s = 0.0;
      for (int k = i; k < C; ++k) {
        const double lv = (k == i) ? 1.0 : (double)l[k * C + i];
        s += lv * (double)dw[(size_t)rowsrc[k] * C + j];
      }
      if (i < j) du[e] += (float)s;
      else dlog_s[i] += (float)(s * (double)sign_s[i] * exp((double)log_s[i]));
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
    // two kernels over a global fp64 workspace: [batch][C][C] LU factors followed by [batch][C] pivots
    const size_t work = (size_t)batch * C * C * sizeof(double), all = work + (size_t)batch * C * sizeof(int);
    GLOWK_CUDA(cudaMallocAsync((void**)&gwork, all, st));
    int* gperm = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(gwork) + work);
    const size_t lu_smem_bytes = 2 * (size_t)C * LU_NB * sizeof(double);            // panel + U12 (96 KB at C = 384)
    cudaError_t e1 = cudaSuccess;
    if (lu_smem_bytes > 40 * 1024)            // (static shared memory counts against the 48 KB default too)
      e1 = cudaFuncSetAttribute(invconv_lu_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lu_smem_bytes);
    if (e1 == cudaSuccess) {
      invconv_lu_big_kernel<<<(unsigned)batch, 1024, lu_smem_bytes, st>>>(w, (int)C, logabsdet_out, gwork, gperm);
      e1 = cudaGetLastError();
    }
    if (e1 == cudaSuccess && winv_out) {
      const size_t xs = 8 * (size_t)C * sizeof(double);
      if (xs > 40 * 1024)
        e1 = cudaFuncSetAttribute(invconv_inverse_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xs);
      if (e1 == cudaSuccess) {
        invconv_inverse_big_kernel<<<dim3((unsigned)ceil_div(C, 8), (unsigned)batch), 256, xs, st>>>(gwork, gperm, (int)C, winv_out);
        e1 = cudaGetLastError();
      }
    }
    cudaFreeAsync(gwork, st);
    if (e1 != cudaSuccess) return fail(GLOWK_ECUDA, "glowk_invconv_prepare: %s", cudaGetErrorString(e1));
    return GLOWK_OK;
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

extern "C" int glowk_invconv_lu_grads(const float* dw, const float* p, const float* l, const float* u,
                                      const float* sign_s, const float* log_s, int64_t C, float* dl, float* du,
                                      float* dlog_s, void* stream) {
  GLOWK_CHECK_ARG(dw && p && l && u && sign_s && log_s && dl && du && dlog_s, "glowk_invconv_lu_grads: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C <= 1024, "glowk_invconv_lu_grads: C out of range");
  const int threads = C <= 16 ? 64 : (C <= 48 ? 256 : 1024);
  lu_grads_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(dw, p, l, u, sign_s, log_s, (int)C, dl, du, dlog_s);
  GLOWK_CHECK_LAUNCH("glowk_invconv_lu_grads");
  return GLOWK_OK;
}
