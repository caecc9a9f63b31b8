// C-ABI plumbing: error reporting, device queries, GEMM dispatch.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace glowk {

static thread_local char g_err[512];

char* last_error_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  // measured on B200 (gpurun_out/ab1_bench.log): inside the captured train-step graph PDL edges cost
  // ~1.2 us MORE per kernel than plain edges (37.8 vs 36.0 ms/step), so the attribute is opt-in (GLOWK_PDL=1)
  static const bool on = []() { const char* e = getenv("GLOWK_PDL"); return e && e[0] == '1'; }();
  return on;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) cached = n;
    cached_dev = dev;
  }
  return cached;
}

// CTAs of `kern` that fit on the device at once (occupancy x SM count): the grid of the looping flow kernels.
// Cached per thread (no shared mutable state); keyed by kernel, block size and dynamic shared memory.
int resident_ctas(const void* kern, int threads, size_t smem) {
  struct Entry { const void* k; int threads; size_t smem; int dev; int ctas; };
  static thread_local Entry cache[32];
  static thread_local int used = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < used; ++i)
    if (cache[i].k == kern && cache[i].threads == threads && cache[i].smem == smem && cache[i].dev == dev) return cache[i].ctas;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 1;
  }
  const int ctas = occ * sm_count();
  if (used < 32) cache[used++] = Entry{kern, threads, smem, dev, ctas};
  return ctas;
}

}  // namespace glowk

using namespace glowk;

extern "C" const char* glowk_last_error(void) { return last_error_buf(); }
extern "C" int glowk_version(void) { return 100; }
extern "C" int glowk_has_tcgen05(void) { return tc_available() ? 1 : 0; }
extern "C" int glowk_debug_gemm_trace(unsigned long long* out16_host) { return gemm_debug_trace(out16_host); }

extern "C" int glowk_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype, int64_t M,
                          int64_t N, int64_t K, int epilogue, const float* bias, const float* logs,
                          float logscale_factor, const void* y, int64_t ldy, float* dlogs, float* dbias,
                          void* out, int out_dtype, int64_t ldo, void* stream) {
  return glowk_gemm_ex(A, lda, B, ldb, act_dtype, M, N, K, epilogue, bias, logs, logscale_factor, y, ldy, dlogs,
                       dbias, out, out_dtype, ldo, 0, 0, stream);
}

extern "C" int glowk_gemm_ex(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype, int64_t M,
                             int64_t N, int64_t K, int epilogue, const float* bias, const float* logs,
                             float logscale_factor, const void* y, int64_t ldy, float* dlogs, float* dbias,
                             void* out, int out_dtype, int64_t ldo, int cluster_m, int cluster_n, void* stream) {
  if (M == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(A && B && out, "glowk_gemm: null pointer");
  GLOWK_CHECK_ARG(M >= 0 && N > 0 && K > 0, "glowk_gemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  GLOWK_CHECK_ARG(lda >= K && ldb >= K && ldo >= N, "glowk_gemm: leading dimensions too small");
  GLOWK_CHECK_ARG(out_dtype == GLOWK_F32 || out_dtype == GLOWK_BF16, "glowk_gemm: bad out_dtype");
  if (epilogue != GLOWK_EPI_STORE) GLOWK_CHECK_ARG(logs, "glowk_gemm: epilogue %d needs logs", epilogue);
  if (epilogue == GLOWK_EPI_ACTNORM_RELU || epilogue == GLOWK_EPI_ACTNORM || epilogue == GLOWK_EPI_ZEROS)
    GLOWK_CHECK_ARG(bias, "glowk_gemm: epilogue %d needs bias", epilogue);
  // dlogs / dbias may be NULL on the bf16 path: the caller recovers them with glowk_conv_actnorm_finish_batched
  // (dbias from a ones column of its weight-gradient GEMM, dlogs from W, dW and dbias)
  if (epilogue == GLOWK_EPI_RELU_BWD)
    GLOWK_CHECK_ARG(y && ldy >= N && ((dlogs && dbias) || act_dtype == GLOWK_BF16), "glowk_gemm: RELU_BWD needs y (and dlogs, dbias on the fp32 path)");
  if (M == 0) return GLOWK_OK;
  EpiParams ep;
  ep.bias = bias; ep.logs = logs; ep.f = logscale_factor; ep.y = y; ep.ldy = ldy;
  ep.y_bf16 = (act_dtype == GLOWK_BF16); ep.dlogs = dlogs; ep.dbias = dbias;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == GLOWK_F32)
    return gemm_f32((const float*)A, lda, (const float*)B, ldb, M, N, K, epilogue, ep, out, out_dtype, ldo, st);
  if (act_dtype == GLOWK_BF16)
    return gemm_bf16_tc(A, lda, B, ldb, M, N, K, epilogue, ep, out, out_dtype, ldo, cluster_m, cluster_n, st);
  return fail(GLOWK_EINVAL, "glowk_gemm: bad act_dtype %d", act_dtype);
}

extern "C" int glowk_gemm_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype, int64_t P,
                                int64_t Mo, int64_t No, float* dW, int64_t lddw, void* stream) {
  GLOWK_CHECK_ARG(A && B && dW, "glowk_gemm_wgrad: null pointer");
  GLOWK_CHECK_ARG(P >= 0 && Mo > 0 && No > 0 && lda >= Mo && ldb >= No && lddw >= No, "glowk_gemm_wgrad: bad shape");
  if (P == 0) return GLOWK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == GLOWK_F32) return wgrad_simt(A, lda, B, ldb, act_dtype, P, Mo, No, dW, lddw, st);
  if (act_dtype == GLOWK_BF16) return wgrad_bf16_tc(A, lda, B, ldb, P, Mo, No, dW, lddw, st);
  return fail(GLOWK_EINVAL, "glowk_gemm_wgrad: bad act_dtype %d", act_dtype);
}
