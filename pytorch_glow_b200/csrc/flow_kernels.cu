// Memory-bound Glow flow kernels (fp32 NCHW flow state), sm_100a.
// Each kernel cites the reference lines (corenel/pytorch-glow) whose ATen sequence it replaces.
#include "common.cuh"

namespace glowk {

// ------------------------------------------------------------------------------------------
// ActNorm elementwise (network/module.py:34-84)
// ------------------------------------------------------------------------------------------
template <int V>
__global__ void actnorm_kernel(const float* __restrict__ x, float* __restrict__ y,
                               const float* __restrict__ bias, const float* __restrict__ logs,
                               float f, int64_t total, int64_t C, int64_t HW, int reverse) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (i >= total) return;
  const int64_t c = (i / HW) % C;
  const float b = bias[c];
  const float l = logs[c] * f;
  if (V == 4) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    if (!reverse) {
      const float s = expf(l);
      v.x = (v.x + b) * s; v.y = (v.y + b) * s; v.z = (v.z + b) * s; v.w = (v.w + b) * s;
    } else {
      const float s = expf(-l);
      v.x = v.x * s - b; v.y = v.y * s - b; v.z = v.z * s - b; v.w = v.w * s - b;
    }
    *reinterpret_cast<float4*>(y + i) = v;
  } else {
    float v = x[i];
    v = reverse ? (v * expf(-l) - b) : ((v + b) * expf(l));
    y[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// ActNorm data-dependent init (network/module.py:86-120).  One CTA per channel, two passes,
// fp64 accumulation, fixed summation order (deterministic).
// ------------------------------------------------------------------------------------------
__global__ void actnorm_init_kernel(const float* __restrict__ x, int64_t N, int64_t HW, int64_t sN,
                                    int64_t sC, int64_t sP, float scale, float f,
                                    float* __restrict__ bias_out, float* __restrict__ logs_out) {
  __shared__ double red[32];
  __shared__ double s_mean;
  const int64_t c = blockIdx.x;
  const int64_t cnt = N * HW;
  const float* base = x + c * sC;
  auto block_sum_d = [&](double v) -> double {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
      r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
  };
  double acc = 0.0;
  for (int64_t e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int64_t n = e / HW, p = e - n * HW;
    acc += (double)base[n * sN + p * sP];
  }
  double tot = block_sum_d(acc);
  if (threadIdx.x == 0) s_mean = tot / (double)cnt;
  __syncthreads();
  const float b = -(float)s_mean;  // bias = -mean (module.py:97-99)
  acc = 0.0;
  for (int64_t e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int64_t n = e / HW, p = e - n * HW;
    const float v = base[n * sN + p * sP] + b;  // centred in fp32 like the reference
    acc += (double)(v * v);
  }
  tot = block_sum_d(acc);
  if (threadIdx.x == 0) {
    const float var = (float)(tot / (double)cnt);
    bias_out[c] = b;
    logs_out[c] = logf(scale / (sqrtf(var) + 1e-6f)) / f;  // module.py:115
  }
}

// The two variants of the data-dependent init the default kernel above does not cover (network/module.py:44-45,
// 62-63, 106-120).  Pass 1, one CTA per channel: stats[c] = mean_c, stats[C + c] = second moment of channel c --
// centred (x - mean_c) in the forward direction; UNcentred in the reverse direction, where the reference initialises
// logs first, from the raw input (module.py:143-146: scale, then centre).
__global__ void actnorm_moments_kernel(const float* __restrict__ x, int64_t N, int64_t HW, int64_t sN, int64_t sC,
                                       int64_t sP, int centred, double* __restrict__ stats, int C) {
  __shared__ double red[32];
  __shared__ double s_mean;
  const int64_t c = blockIdx.x;
  const int64_t cnt = N * HW;
  const float* base = x + c * sC;
  auto block_sum_d = [&](double v) -> double {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
      r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
  };
  double acc = 0.0;
  for (int64_t e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int64_t n = e / HW, p = e - n * HW;
    acc += (double)base[n * sN + p * sP];
  }
  double tot = block_sum_d(acc);
  if (threadIdx.x == 0) s_mean = tot / (double)cnt;
  __syncthreads();
  const float b = centred ? -(float)s_mean : 0.f;
  acc = 0.0;
  for (int64_t e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int64_t n = e / HW, p = e - n * HW;
    const float v = base[n * sN + p * sP] + b;
    acc += (double)(v * v);
  }
  tot = block_sum_d(acc);
  if (threadIdx.x == 0) { stats[c] = s_mean; stats[C + c] = tot / (double)cnt; }
}

// Pass 2, one CTA: batch_variance -> one variance for all channels (every channel has the same count, so the mean of
// the per-channel moments is the reference's mean over the whole tensor, module.py:112-113);
// forward: bias = -mean;  reverse: bias = -mean(x * exp(-f*logs)) = -mean_c * exp(-f*logs_c) (module.py:143-146).
__global__ void actnorm_init_finish_kernel(const double* __restrict__ stats, int C, float scale, float f,
                                           int batch_variance, int reverse, float* __restrict__ bias_out,
                                           float* __restrict__ logs_out) {
  __shared__ double s_all;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int c = 0; c < C; ++c) t += stats[C + c];
    s_all = t / (double)C;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float var = (float)(batch_variance ? s_all : stats[C + c]);
    const float lg = logf(scale / (sqrtf(var) + 1e-6f)) / f;
    logs_out[c] = lg;
    const float m = (float)stats[c];
    bias_out[c] = reverse ? -(m * expf(-(lg * f))) : -m;
  }
}

// ------------------------------------------------------------------------------------------
// Fused ActNorm + channel mix / permutation (model.py:94-103, 142-152; module.py:356-369, 392-397)
// Tile = TP consecutive global pixels; x tile (post-actnorm in forward) staged in smem as
// xs[C][TP]; mixing weights staged transposed wt[i][o].  Thread = (pixel group of V, group of 4
// output channels); warps share the output group so W reads are smem broadcasts.
// ------------------------------------------------------------------------------------------
template <int V>
__global__ void actnorm_mix_kernel(const float* __restrict__ x, float* __restrict__ z,
                                   const float* __restrict__ w, const int64_t* __restrict__ idx,
                                   const float* __restrict__ bias, const float* __restrict__ logs,
                                   float f, int64_t NP, int C, int64_t HW, int reverse, int TP) {
  extern __shared__ __align__(16) float smem[];
  const int Cp = (C + 3) & ~3;
  float* xs = smem;                     // [C][TP]
  float* wt = xs + (size_t)C * TP;      // [C][Cp]  (mix only)
  float* sc = wt + (w ? (size_t)C * Cp : 0);  // [C] scale, [C] bias
  float* bs = sc + C;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int64_t g0 = (int64_t)blockIdx.x * TP;
  const int Q = TP / V;                 // pixel groups per tile

  for (int c = tid; c < C; c += nthr) {
    const float l = logs ? logs[c] * f : 0.f;
    sc[c] = logs ? expf(reverse ? -l : l) : 1.f;
    bs[c] = bias ? bias[c] : 0.f;
  }
  if (w) {
    for (int e = tid; e < C * C; e += nthr) {
      const int o = e / C, i = e - o * C;
      wt[i * Cp + o] = w[e];
    }
  }
  __syncthreads();
  // stage x tile (forward: apply actnorm on the way in)
  for (int e = tid; e < C * Q; e += nthr) {
    const int c = e / Q, q = e - c * Q;
    const int64_t g = g0 + (int64_t)q * V;
    if (g < NP) {
      const int64_t n = g / HW, p = g - n * HW;
      const float* src = x + (n * C + c) * HW + p;
      if (V == 4) {
        float4 v = ld_stream4(src);
        if (!reverse && bias) {
          const float b = bs[c], s = sc[c];
          v.x = (v.x + b) * s; v.y = (v.y + b) * s; v.z = (v.z + b) * s; v.w = (v.w + b) * s;
        }
        *reinterpret_cast<float4*>(xs + (size_t)c * TP + q * 4) = v;
      } else {
        float v = *src;
        if (!reverse && bias) v = (v + bs[c]) * sc[c];
        xs[(size_t)c * TP + q] = v;
      }
    }
  }
  __syncthreads();

  const int q = tid % Q;
  const int64_t g = g0 + (int64_t)q * V;
  if (g >= NP) return;
  const int64_t n = g / HW, p = g - n * HW;
  float* dst = z + n * C * HW + p;
  const int ngrp_blk = nthr / Q;

  if (w == nullptr) {  // permutation: bit-exact gather
    for (int o = tid / Q; o < C; o += ngrp_blk) {
      const int s = (int)idx[o];
      if (V == 4) {
        float4 v = *reinterpret_cast<const float4*>(xs + (size_t)s * TP + q * 4);
        if (reverse && bias) {
          const float b = bs[o], sv = sc[o];
          v.x = v.x * sv - b; v.y = v.y * sv - b; v.z = v.z * sv - b; v.w = v.w * sv - b;
        }
        st_stream4(dst + (int64_t)o * HW, v);
      } else {
        float v = xs[(size_t)s * TP + q];
        if (reverse && bias) v = v * sc[o] - bs[o];
        dst[(int64_t)o * HW] = v;
      }
    }
    return;
  }

  const int ngrp = Cp / 4;
  for (int og = tid / Q; og < ngrp; og += ngrp_blk) {
    float acc[4][V];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[a][v] = 0.f;
    for (int i = 0; i < C; ++i) {
      const float4 wv = *reinterpret_cast<const float4*>(wt + i * Cp + og * 4);
      float xv[V];
      if (V == 4) {
        const float4 t = *reinterpret_cast<const float4*>(xs + (size_t)i * TP + q * 4);
        xv[0] = t.x; xv[1 % V] = t.y; xv[2 % V] = t.z; xv[3 % V] = t.w;
      } else {
        xv[0] = xs[(size_t)i * TP + q];
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        acc[0][v] = fmaf(wv.x, xv[v], acc[0][v]);
        acc[1][v] = fmaf(wv.y, xv[v], acc[1][v]);
        acc[2][v] = fmaf(wv.z, xv[v], acc[2][v]);
        acc[3][v] = fmaf(wv.w, xv[v], acc[3][v]);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int o = og * 4 + a;
      if (o < C) {
        if (reverse && bias) {
          const float b = bs[o], sv = sc[o];
#pragma unroll
          for (int v = 0; v < V; ++v) acc[a][v] = acc[a][v] * sv - b;
        }
        if (V == 4) {
          st_stream4(dst + (int64_t)o * HW, make_float4(acc[a][0], acc[a][1 % V], acc[a][2 % V], acc[a][3 % V]));
        } else {
          dst[(int64_t)o * HW] = acc[a][0];
        }
      }
    }
  }
}

// Fallback for very wide C (W does not fit in shared memory): one thread per output element.
__global__ void actnorm_mix_wide_kernel(const float* __restrict__ x, float* __restrict__ z,
                                        const float* __restrict__ w, const int64_t* __restrict__ idx,
                                        const float* __restrict__ bias, const float* __restrict__ logs,
                                        float f, int64_t N, int C, int64_t HW, int reverse) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * C * HW) return;
  const int64_t p = e % HW;
  const int o = (int)((e / HW) % C);
  const int64_t n = e / (HW * C);
  const float* xb = x + n * C * HW + p;
  float r;
  if (w) {
    r = 0.f;
    for (int i = 0; i < C; ++i) {
      float v = xb[(int64_t)i * HW];
      if (!reverse && bias) v = (v + bias[i]) * expf(logs[i] * f);
      r = fmaf(w[(int64_t)o * C + i], v, r);
    }
  } else {
    const int s = (int)idx[o];
    r = xb[(int64_t)s * HW];
    if (!reverse && bias) r = (r + bias[s]) * expf(logs[s] * f);
  }
  if (reverse && bias) r = r * expf(-logs[o] * f) - bias[o];
  z[e] = r;
}

// ------------------------------------------------------------------------------------------
// Squeeze2d / unsqueeze (module.py:551-591): pure index map, bit-exact.
// ------------------------------------------------------------------------------------------
__global__ void squeeze_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total,
                               int64_t C, int64_t H, int64_t W, int64_t sN, int f, int reverse) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  if (!reverse) {
    // output index e over [N, C*f*f, H/f, W/f]
    const int64_t Wo = W / f, Ho = H / f, Co = C * f * f;
    const int64_t j = e % Wo, i = (e / Wo) % Ho, co = (e / (Wo * Ho)) % Co, n = e / (Wo * Ho * Co);
    const int64_t c = co / (f * f), r = co % (f * f), fh = r / f, fw = r % f;
    y[e] = x[n * sN + (c * H + i * f + fh) * W + j * f + fw];
  } else {
    // x: [N, C, H, W] -> y: [N, C/f^2, H*f, W*f]; output index e
    const int64_t Wo = W * f, Ho = H * f, Co = C / (f * f);
    const int64_t jj = e % Wo, ii = (e / Wo) % Ho, c = (e / (Wo * Ho)) % Co, n = e / (Wo * Ho * Co);
    const int64_t i = ii / f, fh = ii % f, j = jj / f, fw = jj % f;
    y[e] = x[n * sN + ((c * f * f + fh * f + fw) * H + i) * W + j];
  }
}

// ------------------------------------------------------------------------------------------
// im2col for SAME-padded kxk conv: rows matrix with taps folded into K.
// One thread per 8 destination columns (16 B of bf16 / 32 B of fp32).
// ------------------------------------------------------------------------------------------
template <typename T, bool ROWS_SRC>
__global__ void im2col_kernel(const float* __restrict__ src, int64_t ld_src, int64_t NP, int64_t Ctot,
                              int64_t c0, int Cin, int H, int W, int ks, int flip,
                              T* __restrict__ dst, int64_t ld) {
  const int chunks = (int)(ld / 8);
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NP * chunks) return;
  const int64_t pix = e / chunks;
  const int k0 = (int)(e - pix * chunks) * 8;
  const int HW = H * W;
  const int64_t n = pix / HW;
  const int p = (int)(pix - n * HW);
  const int yy = p / W, xx = p - yy * W;
  const int pad = (ks - 1) / 2, T2 = ks * ks, K = T2 * Cin;
  // walk (tap, ci) incrementally: one division per thread instead of two per element
  int tap = k0 / Cin, ci = k0 - tap * Cin;
  int cur_tap = -1;
  bool inb = false;
  const float* base = src;
  float v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float r = 0.f;
    if (k0 + u < K) {
      if (tap != cur_tap) {
        const int tt = flip ? T2 - 1 - tap : tap;
        const int ky = ks == 3 ? tt / 3 : 0, kx = tt - ky * ks;
        const int sy = yy + ky - pad, sx = xx + kx - pad;
        inb = sy >= 0 && sy < H && sx >= 0 && sx < W;
        // Ctot carries the batch stride of an NCHW source
        base = ROWS_SRC ? src + (n * HW + (int64_t)sy * W + sx) * ld_src + c0
                        : src + n * Ctot + (c0 * (int64_t)H + sy) * W + sx;
        cur_tap = tap;
      }
      if (inb) r = ROWS_SRC ? base[ci] : base[(int64_t)ci * HW];
    }
    v[u] = r;
    if (++ci == Cin) { ci = 0; ++tap; }
  }
  T* d = dst + pix * ld + k0;
  if (sizeof(T) == 2) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    *reinterpret_cast<uint4*>(d) = *reinterpret_cast<uint4*>(h);
  } else {
    float* df = reinterpret_cast<float*>(d);
    *reinterpret_cast<float4*>(df) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(df + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// im2col of a 3x3 SAME conv from a pixel-major fp32 source into bf16 rows, ONE WARP PER PIXEL: lane l owns
// V consecutive output columns k = (pass*32 + l)*V .. +V-1 (V in {2,4,8} divides Cin, so the V values of a lane
// belong to one tap and are contiguous on both sides: one vector load, one vector store, and the pixel's whole
// row is written as coalesced 32*V*2-byte segments).  Pixel coordinates are warp-uniform; columns >= 9*Cin are
// written as zeros by the same instruction stream (no divergent padding loop).
template <int V>
__global__ void __launch_bounds__(256)
im2col_rows_warp_kernel(const float* __restrict__ src, int ld_src, int NP, int c0, int Cin, int H, int W,
                        FastDiv divW, FastDiv divHW, FastDiv divCin, int flip, __nv_bfloat16* __restrict__ dst, int ld) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * 8 + warp;
  if (pix >= NP) return;
  const int HW = H * W;
  const int n = fdiv(pix, divHW);
  const int p = pix - n * HW;
  const int yy = fdiv(p, divW), xx = p - yy * W;
  const int K = 9 * Cin;
  const float* sbase = src + (int64_t)pix * ld_src + c0;            // this pixel; taps are +-W, +-1 rows away
  __nv_bfloat16* drow = dst + (int64_t)pix * ld;
  for (int k = lane * V; k < ld; k += 32 * V) {
    float v[V];
#pragma unroll
    for (int u = 0; u < V; ++u) v[u] = 0.f;
    if (k < K) {
      const int tap = fdiv(k, divCin), ci = k - tap * Cin;
      const int tt = flip ? 8 - tap : tap;
      const int ky = tt / 3, kx = tt - ky * 3;
      const int sy = yy + ky - 1, sx = xx + kx - 1;
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
        const float* sp = sbase + ((ky - 1) * W + (kx - 1)) * ld_src + ci;
        if (V == 2) {
          const float2 t = *reinterpret_cast<const float2*>(sp);
          v[0] = t.x; v[1] = t.y;
        } else {
#pragma unroll
          for (int q = 0; q < V / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(sp + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
          }
        }
      }
    }
    __nv_bfloat162 h[V / 2];
#pragma unroll
    for (int u = 0; u < V / 2; ++u) h[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    if (V == 2) *reinterpret_cast<__nv_bfloat162*>(drow + k) = h[0];
    else if (V == 4) *reinterpret_cast<uint2*>(drow + k) = *reinterpret_cast<uint2*>(h);
    else *reinterpret_cast<uint4*>(drow + k) = *reinterpret_cast<uint4*>(h);
  }
}

// Same operation, ONE THREAD PER 16-BYTE OUTPUT PIECE (8 bf16 columns), the piece index fixed per thread: the
// (tap, channel) decode of its 8/V source groups is done once, then the thread walks `iters` pixels of a
// contiguous pixel run (per pixel: two fast divisions, 8/V border predicates + vector loads, one 16-byte store;
// consecutive threads write consecutive pieces, consecutive slots consecutive pixels).  The warp-per-pixel
// variant above is issue-bound (ncu: 76 % SM throughput for 80 MB at level 1); this one does ~4x fewer
// instructions per byte.  Results are identical (same loads, same float -> bf16 rounding).
template <int V>
__global__ void __launch_bounds__(256)
im2col_rows_piece_kernel(const float* __restrict__ src, int ld_src, int NP, int c0, int Cin, int H, int W,
                         FastDiv divW, FastDiv divHW, FastDiv divChunks, int flip, __nv_bfloat16* __restrict__ dst,
                         int ld, int ppb, int iters, int ones_col) {
  pdl_trigger();
  pdl_wait();
  constexpr int NG = 8 / V;
  const int chunks = (int)divChunks.d;
  const int slot = fdiv((int)threadIdx.x, divChunks), j = (int)threadIdx.x - slot * chunks;
  if (slot >= ppb) return;
  const int K = 9 * Cin, HW = H * W;
  int off[NG], dy[NG], dx[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int k = 8 * j + g * V;
    if (k < K) {
      const int tap = k / Cin, ci = k - tap * Cin;
      const int tt = flip ? 8 - tap : tap;
      const int ky = tt / 3, kx = tt - ky * 3;
      dy[g] = ky - 1; dx[g] = kx - 1;
      off[g] = ((ky - 1) * W + (kx - 1)) * ld_src + c0 + ci;
    } else {
      dy[g] = -(1 << 20); dx[g] = 0; off[g] = 0;      // zero padding columns: never in bounds
    }
  }
  // ones_col >= 9*Cin (a zero-padding column, -1 = none) is written as 1.0: the weight-gradient GEMM dW = d^T a1
  // then delivers sum_p d[p][n] -- the bias gradient of the ActNorm after the conv -- in that column for free,
  // while the conv itself (zero weight there) and its dgrad are unaffected
  const int ones_u = ones_col >= 0 ? ones_col - 8 * j : -1;
  const int pix0 = blockIdx.x * (ppb * iters) + slot;
  for (int it = 0; it < iters; ++it) {
    const int pix = pix0 + it * ppb;
    if (pix >= NP) break;
    const int n = fdiv(pix, divHW);
    const int p = pix - n * HW;
    const int yy = fdiv(p, divW), xx = p - yy * W;
    const float* sbase = src + (int64_t)pix * ld_src;
    float v[8];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const bool ok = (unsigned)(yy + dy[g]) < (unsigned)H && (unsigned)(xx + dx[g]) < (unsigned)W;
      if (V == 2) {
        float2 t = make_float2(0.f, 0.f);
        if (ok) t = *reinterpret_cast<const float2*>(sbase + off[g]);
        v[2 * g] = t.x; v[2 * g + 1] = t.y;
      } else {
#pragma unroll
        for (int q = 0; q < V / 4; ++q) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) t = *reinterpret_cast<const float4*>(sbase + off[g] + 4 * q);
          v[g * V + 4 * q] = t.x; v[g * V + 4 * q + 1] = t.y; v[g * V + 4 * q + 2] = t.z; v[g * V + 4 * q + 3] = t.w;
        }
      }
    }
    if (ones_u >= 0 && ones_u < 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (u == ones_u) ? 1.f : v[u];
    }
    __nv_bfloat162 h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    *reinterpret_cast<uint4*>(dst + (int64_t)pix * ld + 8 * j) = *reinterpret_cast<uint4*>(h);
  }
}

template <typename T>
__global__ void rows_to_nchw_kernel(const T* __restrict__ rows, int64_t ld, float* __restrict__ dst,
                                    int64_t total, int64_t C, int64_t HW) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t p = e % HW, c = (e / HW) % C, n = e / (HW * C);
  dst[e] = to_f32<T>(rows[(n * HW + p) * ld + c]);
}

__global__ void tapsum_kernel(const float* __restrict__ P, int64_t ldp, float* __restrict__ dst,
                              int64_t total, int64_t Ctot, int64_t c0, int C, int H, int W, int flip,
                              int accumulate) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int HW = H * W;
  const int p = (int)(e % HW);
  const int c = (int)((e / HW) % C);
  const int64_t n = e / ((int64_t)HW * C);
  const int yy = p / W, xx = p - yy * W;
  float r = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int tap = flip ? 8 - t : t;
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W)
      r += P[(n * HW + (int64_t)sy * W + sx) * ldp + t * C + c];
  }
  float* d = dst + (n * Ctot + c0 + c) * HW + p;
  *d = accumulate ? (*d + r) : r;
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, int O, int I, int ks, int layout,
                                   T* __restrict__ dst, int64_t rows, int64_t ld) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * ld) return;
  const int64_t r = e / ld, col = e - r * ld;
  const int T2 = ks * ks;
  int o = -1, i = -1, tap = 0;
  if (layout == 0) {            // dst[o][tap*I + i]
    if (r < O && col < (int64_t)T2 * I) { o = (int)r; tap = (int)(col / I); i = (int)(col - (int64_t)tap * I); }
  } else if (layout == 1) {     // dst[tap*O + o][i]
    if (r < (int64_t)T2 * O && col < I) { tap = (int)(r / O); o = (int)(r - (int64_t)tap * O); i = (int)col; }
  } else if (layout == 2) {     // dst[tap*I + i][o]
    if (r < (int64_t)T2 * I && col < O) { tap = (int)(r / I); i = (int)(r - (int64_t)tap * I); o = (int)col; }
  } else {                      // dst[i][tap*O + o]
    if (r < I && col < (int64_t)T2 * O) { i = (int)r; tap = (int)(col / O); o = (int)(col - (int64_t)tap * O); }
  }
  float v = 0.f;
  if (o >= 0) v = w[((int64_t)o * I + i) * T2 + tap];
  dst[e] = from_f32<T>(v);
}

// ------------------------------------------------------------------------------------------
// Coupling (model.py:105-115 fwd, 131-140 rev) fused with the tap gather-sum that finishes
// Conv2dZeros (module.py:295-296).  grid = (nblk, N); thread = pixel; per-CTA partial of
// sum(log scale) written to partials[n][blk] (deterministic two-stage reduction).
// ------------------------------------------------------------------------------------------
__global__ void coupling_kernel(const float* __restrict__ P, int64_t ldp, const float* __restrict__ bias3,
                                const float* __restrict__ logs3, float f, float* __restrict__ z,
                                float* __restrict__ partials, float* __restrict__ h_save, int C, int H,
                                int W, int affine, int reverse) {
  __shared__ float red[32];
  const int HW = H * W;
  const int64_t n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int Ch = C / 2;
  const int Cout = affine ? C : Ch;
  float lsum = 0.f;
  if (p < HW) {
    const int yy = p / W, xx = p - yy * W;
    const float* nb[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int sy = yy + t / 3 - 1, sx = xx + t % 3 - 1;
      nb[t] = (sy >= 0 && sy < H && sx >= 0 && sx < W)
                  ? P + (n * HW + (int64_t)sy * W + sx) * ldp + t * Cout : nullptr;
    }
    float* z2 = z + (n * C + Ch) * HW + p;
    float* hs = h_save ? h_save + (n * HW + p) * (int64_t)Cout : nullptr;
    for (int j = 0; j < Ch; ++j) {
      if (affine) {
        float u0 = 0.f, u1 = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t)
          if (nb[t]) {
            const float2 v = *reinterpret_cast<const float2*>(nb[t] + 2 * j);
            u0 += v.x; u1 += v.y;
          }
        const float shift = (u0 + bias3[2 * j]) * expf(logs3[2 * j] * f);
        const float hsc = (u1 + bias3[2 * j + 1]) * expf(logs3[2 * j + 1] * f);
        const float scale = 1.f / (1.f + expf(-(hsc + 2.f)));   // F.sigmoid(scale + 2.)
        float v = z2[(int64_t)j * HW];
        if (!reverse) { v = (v + shift) * scale; lsum += logf(scale); }
        else { v = v / scale - shift; lsum -= logf(scale); }
        z2[(int64_t)j * HW] = v;
        if (hs) { hs[2 * j] = shift; hs[2 * j + 1] = hsc; }
      } else {
        float u = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t)
          if (nb[t]) u += nb[t][j];
        const float h = (u + bias3[j]) * expf(logs3[j] * f);
        float v = z2[(int64_t)j * HW];
        v = reverse ? v - h : v + h;
        z2[(int64_t)j * HW] = v;
        if (hs) hs[j] = h;
      }
    }
  }
  if (partials) {
    const float tot = block_sum(lsum, red);
    if (threadIdx.x == 0) partials[n * gridDim.x + blockIdx.x] = tot;
  }
}

__global__ void logdet_finish_kernel(const float* __restrict__ logdet_in, float* __restrict__ logdet_out,
                                     const float* __restrict__ logs, int C, float f,
                                     const float* __restrict__ logabsdet, const float* __restrict__ partials,
                                     int nblk, float hw, float sign, int64_t N) {
  __shared__ float red[32];
  __shared__ float s_term;
  float a = 0.f;
  if (logs)
    for (int c = threadIdx.x; c < C; c += blockDim.x) a += logs[c] * f;
  const float tot = block_sum(a, red);
  if (threadIdx.x == 0) {
    float t = 0.f;
    if (logs) t += tot * hw;                     // torch.sum(logs)*HW   (module.py:78-80)
    s_term = t;
  }
  __syncthreads();
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
    float v = logdet_in ? logdet_in[n] : 0.f;
    v += sign * s_term;
    if (logabsdet) v += sign * (logabsdet[0] * hw);  // log|det W| * HW  (module.py:357)
    if (partials) {
      float s = 0.f;
      for (int b = 0; b < nblk; ++b) s += partials[n * nblk + b];
      v += s;
    }
    logdet_out[n] = v;
  }
}

// GaussianDiag.logp (module.py:437-467): one CTA per sample, fixed order.
__global__ void gaussian_logp_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ x,
                                     int64_t C, int64_t HW, int64_t c0, int64_t Cz,
                                     const float* __restrict__ logdet_in, float* __restrict__ logdet_out) {
  __shared__ float red[32];
  const int64_t n = blockIdx.x;
  const float log2pi = 1.8378770664093453f;
  float acc = 0.f;
  const int64_t cnt = Cz * HW;
  for (int64_t e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int64_t j = e / HW, p = e - j * HW;
    const float v = x[(n * C + c0 + j) * HW + p];
    float mean = 0.f, lg = 0.f;
    if (h) {
      const float2 ml = *reinterpret_cast<const float2*>(h + (n * HW + p) * ldh + 2 * j);
      mean = ml.x; lg = ml.y;
    }
    const float d = v - mean;
    acc += -0.5f * (log2pi + 2.f * lg + (d * d) / expf(2.f * lg));
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) logdet_out[n] = tot + (logdet_in ? logdet_in[n] : 0.f);
}

__global__ void split2d_sample_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ z1,
                                      const float* __restrict__ eps, float* __restrict__ out,
                                      int64_t total, int64_t Ch, int64_t HW) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t p = e % HW, c = (e / HW) % (2 * Ch), n = e / (HW * 2 * Ch);
  float v;
  if (c < Ch) {
    v = z1[(n * Ch + c) * HW + p];
  } else {
    const int64_t j = c - Ch;
    const float2 ml = *reinterpret_cast<const float2*>(h + (n * HW + p) * ldh + 2 * j);
    v = ml.x + expf(ml.y) * eps[(n * Ch + j) * HW + p];   // mean + exp(logs)*eps (module.py:482-483)
  }
  out[e] = v;
}


// Loss head of Glow.normal_flow + Glow.generative_loss (network/model.py:425-427, 435-450, 496-498) with the plain
// N(0, I) top prior: per sample  objective = ld[n] + c0 + sum log N(z[n]; 0, I)  (c0 = -log(n_bins) * D_x),
// nll[n] = -objective / denom  (denom = ln2 * D_x), and loss = mean_n nll[n], summed in index order by the last CTA
// to finish (the ticket resets itself).  Summation order of the prior term is that of gaussian_logp_kernel.
__global__ void nll_head_kernel(const float* __restrict__ z, int64_t D, const float* __restrict__ ld, float c0,
                                float denom, int N, float* __restrict__ nll, float* __restrict__ loss,
                                unsigned* __restrict__ ticket) {
  __shared__ float red[32];
  __shared__ bool s_last;
  const int64_t n = blockIdx.x;
  const float log2pi = 1.8378770664093453f;
  float acc = 0.f;
  for (int64_t e = threadIdx.x; e < D; e += blockDim.x) {
    const float v = z[n * D + e];
    acc += -0.5f * (log2pi + v * v);
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float objective = tot + (ld ? ld[n] : 0.f) + c0;
    nll[n] = (-objective) / denom;
    s_last = false;
    if (loss) {
      __threadfence();
      s_last = atomicAdd(ticket, 1u) == (unsigned)(N - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float part = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) part += __ldcg(nll + i);
  const float sum = block_sum(part, red);
  if (threadIdx.x == 0) { loss[0] = sum / (float)N; *ticket = 0u; }
}

// Its adjoint.  g_loss = dL/dloss (device scalar), g_nll = dL/dnll [N], dz_in = dL/dz [N][D] -- each may be null; with
// both g's null the loss gradient is 1.  coef_n = (g_loss / N + g_nll[n]) / denom:
//   dz = dz_in + z * coef_n,   dld[n] = -coef_n.
__global__ void nll_head_bwd_kernel(const float* __restrict__ z, const float* __restrict__ g_loss,
                                    const float* __restrict__ g_nll, const float* __restrict__ dz_in, float denom,
                                    int N, int64_t D, float* __restrict__ dz, float* __restrict__ dld) {
  const int n = blockIdx.y;
  float g = (g_loss ? g_loss[0] : (g_nll ? 0.f : 1.f)) / (float)N;
  if (g_nll) g += g_nll[n];
  const float coef = g / denom;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < D) {
    const int64_t i = (int64_t)n * D + e;
    const float v = z[i] * coef;
    dz[i] = dz_in ? dz_in[i] + v : v;
  }
  if (e == 0) dld[n] = -coef;
}

}  // namespace glowk

using namespace glowk;

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" int glowk_actnorm(const float* x, float* y, const float* bias, const float* logs,
                             float logscale_factor, int64_t N, int64_t C, int64_t HW, int reverse,
                             void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && y && bias && logs, "glowk_actnorm: null pointer");
  GLOWK_CHECK_ARG(N >= 0 && C > 0 && HW > 0, "glowk_actnorm: bad shape N=%lld C=%lld HW=%lld",
                  (long long)N, (long long)C, (long long)HW);
  const int64_t total = N * C * HW;
  if (total == 0) return GLOWK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
  if (vec)
    actnorm_kernel<4><<<(unsigned)ceil_div(total / 4, 256), 256, 0, st>>>(x, y, bias, logs, logscale_factor, total, C, HW, reverse);
  else
    actnorm_kernel<1><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(x, y, bias, logs, logscale_factor, total, C, HW, reverse);
  GLOWK_CHECK_LAUNCH("glowk_actnorm");
  return GLOWK_OK;
}

extern "C" int glowk_actnorm_init(const void* x, int act_dtype, int64_t N, int64_t C, int64_t HW,
                                  int64_t sN, int64_t sC, int64_t sP, float scale, float logscale_factor,
                                  float* bias_out, float* logs_out, void* stream) {
  GLOWK_CHECK_ARG(x && bias_out && logs_out, "glowk_actnorm_init: null pointer");
  GLOWK_CHECK_ARG(act_dtype == GLOWK_F32, "glowk_actnorm_init: statistics input must be fp32");
  GLOWK_CHECK_ARG(N > 0 && C > 0 && HW > 0, "glowk_actnorm_init: empty batch");
  actnorm_init_kernel<<<(unsigned)C, 512, 0, (cudaStream_t)stream>>>(
      (const float*)x, N, HW, sN, sC, sP, scale, logscale_factor, bias_out, logs_out);
  GLOWK_CHECK_LAUNCH("glowk_actnorm_init");
  return GLOWK_OK;
}

extern "C" int glowk_actnorm_init_ex(const float* x, int64_t N, int64_t C, int64_t HW, int64_t sN, int64_t sC,
                                     int64_t sP, float scale, float logscale_factor, int batch_variance, int reverse,
                                     float* bias_out, float* logs_out, void* stream) {
  GLOWK_CHECK_ARG(x && bias_out && logs_out, "glowk_actnorm_init_ex: null pointer");
  GLOWK_CHECK_ARG(N > 0 && C > 0 && HW > 0 && C < (1 << 20), "glowk_actnorm_init_ex: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  double* stats = nullptr;
  GLOWK_CUDA(cudaMallocAsync((void**)&stats, (size_t)(2 * C) * sizeof(double), st));
  actnorm_moments_kernel<<<(unsigned)C, 512, 0, st>>>(x, N, HW, sN, sC, sP, reverse ? 0 : 1, stats, (int)C);
  actnorm_init_finish_kernel<<<1, 256, 0, st>>>(stats, (int)C, scale, logscale_factor, batch_variance, reverse,
                                                bias_out, logs_out);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(stats, st);
  if (e != cudaSuccess) return fail(GLOWK_ECUDA, "glowk_actnorm_init_ex: %s", cudaGetErrorString(e));
  return GLOWK_OK;
}

extern "C" int glowk_actnorm_mix(const float* x, float* z, const float* w, const int64_t* idx,
                                 const float* bias, const float* logs, float logscale_factor, int64_t N,
                                 int64_t C, int64_t HW, int reverse, void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && z, "glowk_actnorm_mix: null pointer");
  GLOWK_CHECK_ARG((w != nullptr) != (idx != nullptr), "glowk_actnorm_mix: exactly one of w / idx");
  GLOWK_CHECK_ARG((bias != nullptr) == (logs != nullptr), "glowk_actnorm_mix: bias and logs go together");
  GLOWK_CHECK_ARG(x != z, "glowk_actnorm_mix: in-place not supported");
  GLOWK_CHECK_ARG(N >= 0 && C > 0 && HW > 0, "glowk_actnorm_mix: bad shape");
  const int64_t NP = N * HW;
  if (NP == 0) return GLOWK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)x | (uintptr_t)z) % 16 == 0);
  const int V = vec ? 4 : 1;
  const int TP = 32 * V;
  const int Cp = ((int)C + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)C * TP + (w ? (size_t)C * Cp : 0) + 2 * (size_t)C);
  if (smem > 200 * 1024) {
    const int64_t total = N * C * HW;
    actnorm_mix_wide_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(x, z, w, idx, bias, logs, logscale_factor, N, (int)C, HW, reverse);
    GLOWK_CHECK_LAUNCH("glowk_actnorm_mix(wide)");
    return GLOWK_OK;
  }
  int groups = w ? Cp / 4 : (int)C;
  if (groups > 8) groups = 8;
  const int threads = 32 * groups;
  const unsigned grid = (unsigned)ceil_div(NP, TP);
  if (vec) {
    if (smem > 48 * 1024) GLOWK_CUDA(cudaFuncSetAttribute(actnorm_mix_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    actnorm_mix_kernel<4><<<grid, threads, smem, st>>>(x, z, w, idx, bias, logs, logscale_factor, NP, (int)C, HW, reverse, TP);
  } else {
    if (smem > 48 * 1024) GLOWK_CUDA(cudaFuncSetAttribute(actnorm_mix_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    actnorm_mix_kernel<1><<<grid, threads, smem, st>>>(x, z, w, idx, bias, logs, logscale_factor, NP, (int)C, HW, reverse, TP);
  }
  GLOWK_CHECK_LAUNCH("glowk_actnorm_mix");
  return GLOWK_OK;
}

extern "C" int glowk_squeeze2d(const float* x, float* y, int64_t N, int64_t C, int64_t H, int64_t W,
                               int64_t sN, int factor, int reverse, void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && y && x != y, "glowk_squeeze2d: bad pointers");
  GLOWK_CHECK_ARG(factor >= 1, "glowk_squeeze2d: factor must be >= 1");
  GLOWK_CHECK_ARG(sN >= C * H * W, "glowk_squeeze2d: batch stride smaller than one sample");
  if (!reverse) GLOWK_CHECK_ARG(H % factor == 0 && W % factor == 0, "glowk_squeeze2d: H,W not divisible by factor");  // module.py:588
  else GLOWK_CHECK_ARG(C >= factor * factor && C % (factor * factor) == 0, "glowk_squeeze2d: C not divisible by factor^2");  // module.py:566
  const int64_t total = N * C * H * W;
  if (total == 0) return GLOWK_OK;
  squeeze_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, total, C, H, W, sN, factor, reverse);
  GLOWK_CHECK_LAUNCH("glowk_squeeze2d");
  return GLOWK_OK;
}

static int im2col_common(const float* src, int64_t ld_src, bool rows_src, int64_t N, int64_t Ctot, int64_t c0,
                         int64_t Cin, int64_t H, int64_t W, int ksize, int flip, void* dst, int act_dtype,
                         int64_t ld, void* stream, int64_t ones_col = -1) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(src && dst, "glowk_im2col: null pointer");
  GLOWK_CHECK_ARG(ksize == 1 || ksize == 3, "glowk_im2col: kernel size must be 1 or 3");
  GLOWK_CHECK_ARG(ld % 8 == 0 && ld >= (int64_t)ksize * ksize * Cin, "glowk_im2col: ld=%lld must be a multiple of 8 and >= k*k*Cin", (long long)ld);
  GLOWK_CHECK_ARG(c0 >= 0 && (rows_src ? c0 + Cin <= ld_src : (c0 + Cin) * H * W <= Ctot), "glowk_im2col: channel window out of range");
  GLOWK_CHECK_ARG(((uintptr_t)dst) % 16 == 0, "glowk_im2col: dst must be 16-byte aligned");
  const int64_t NP = N * H * W;
  if (NP == 0) return GLOWK_OK;
  const int64_t total = NP * (ld / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (rows_src && ksize == 3 && act_dtype == GLOWK_BF16 && Cin % 2 == 0 && c0 % 2 == 0 && ld_src % 2 == 0 &&
      ((uintptr_t)src) % 16 == 0 && NP * (ld > ld_src ? ld : ld_src) < (1ll << 31)) {
    const FastDiv dW_ = make_fastdiv(W), dHW = make_fastdiv(H * W), dC = make_fastdiv(Cin);
    __nv_bfloat16* d = (__nv_bfloat16*)dst;
    const int chunks = (int)(ld / 8);
    static const bool warp_variant = getenv("GLOWK_IM2COL_WARP") != nullptr;     // A/B switch for profiling
    GLOWK_CHECK_ARG(ones_col < 0 || (ones_col >= 9 * Cin && ones_col < ld && chunks <= 256), "glowk_im2col_rows_ones: ones_col=%lld must be a padding column", (long long)ones_col);
    if (chunks <= 256 && (!warp_variant || ones_col >= 0)) {
      const int ppb = 256 / chunks;
      int64_t iters = ceil_div(NP, ppb) / (16 * (int64_t)sm_count());             // >= 16 CTAs per SM before CTAs loop
      iters = iters < 1 ? 1 : (iters > 8 ? 8 : iters);
      const unsigned gp = (unsigned)ceil_div(NP, ppb * iters);
      const FastDiv dCh = make_fastdiv(chunks);
#define GLOWK_IM2COL_PIECE(V_)                                                                                          \
  GLOWK_CUDA(launch_pdl(im2col_rows_piece_kernel<V_>, gp, 256, 0, st, src, (int)ld_src, (int)NP, (int)c0, (int)Cin,     \
                        (int)H, (int)W, dW_, dHW, dCh, flip, d, (int)ld, ppb, (int)iters, (int)ones_col))
      if (Cin % 8 == 0 && c0 % 4 == 0 && ld_src % 4 == 0) GLOWK_IM2COL_PIECE(8);
      else if (Cin % 4 == 0 && c0 % 4 == 0 && ld_src % 4 == 0) GLOWK_IM2COL_PIECE(4);
      else GLOWK_IM2COL_PIECE(2);
#undef GLOWK_IM2COL_PIECE
      GLOWK_CHECK_LAUNCH("glowk_im2col_rows(piece)");
      return GLOWK_OK;
    }
    const unsigned gw = (unsigned)ceil_div(NP, 8);
    if (Cin % 8 == 0 && c0 % 4 == 0 && ld_src % 4 == 0)
      GLOWK_CUDA(launch_pdl(im2col_rows_warp_kernel<8>, gw, 256, 0, st, src, (int)ld_src, (int)NP, (int)c0, (int)Cin, (int)H, (int)W, dW_, dHW, dC, flip, d, (int)ld));
    else if (Cin % 4 == 0 && c0 % 4 == 0 && ld_src % 4 == 0)
      GLOWK_CUDA(launch_pdl(im2col_rows_warp_kernel<4>, gw, 256, 0, st, src, (int)ld_src, (int)NP, (int)c0, (int)Cin, (int)H, (int)W, dW_, dHW, dC, flip, d, (int)ld));
    else
      GLOWK_CUDA(launch_pdl(im2col_rows_warp_kernel<2>, gw, 256, 0, st, src, (int)ld_src, (int)NP, (int)c0, (int)Cin, (int)H, (int)W, dW_, dHW, dC, flip, d, (int)ld));
    GLOWK_CHECK_LAUNCH("glowk_im2col_rows(warp)");
    return GLOWK_OK;
  }
  GLOWK_CHECK_ARG(ones_col < 0, "glowk_im2col_rows_ones: needs a pixel-major fp32 source, a 3x3 kernel, bf16 output and even Cin / c0 / pitch");
  const unsigned grid = (unsigned)ceil_div(total, 256);
#define LAUNCH_IM2COL(T, R) im2col_kernel<T, R><<<grid, 256, 0, st>>>(src, ld_src, NP, Ctot, c0, (int)Cin, (int)H, (int)W, ksize, flip, (T*)dst, ld)
  if (act_dtype == GLOWK_BF16) { if (rows_src) LAUNCH_IM2COL(__nv_bfloat16, true); else LAUNCH_IM2COL(__nv_bfloat16, false); }
  else if (act_dtype == GLOWK_F32) { if (rows_src) LAUNCH_IM2COL(float, true); else LAUNCH_IM2COL(float, false); }
  else return fail(GLOWK_EINVAL, "glowk_im2col: bad act_dtype %d", act_dtype);
#undef LAUNCH_IM2COL
  GLOWK_CHECK_LAUNCH("glowk_im2col");
  return GLOWK_OK;
}

extern "C" int glowk_im2col(const float* src, int64_t N, int64_t sN, int64_t c0, int64_t Cin, int64_t H,
                            int64_t W, int ksize, int flip, void* dst, int act_dtype, int64_t ld, void* stream) {
  return im2col_common(src, 0, false, N, sN, c0, Cin, H, W, ksize, flip, dst, act_dtype, ld, stream);
}

extern "C" int glowk_im2col_rows(const float* src, int64_t ld_src, int64_t N, int64_t c0, int64_t Cin, int64_t H,
                                 int64_t W, int ksize, int flip, void* dst, int act_dtype, int64_t ld, void* stream) {
  return im2col_common(src, ld_src, true, N, 0, c0, Cin, H, W, ksize, flip, dst, act_dtype, ld, stream);
}

extern "C" int glowk_im2col_rows_ones(const float* src, int64_t ld_src, int64_t N, int64_t c0, int64_t Cin, int64_t H,
                                      int64_t W, int ksize, int flip, void* dst, int act_dtype, int64_t ld,
                                      int64_t ones_col, void* stream) {
  return im2col_common(src, ld_src, true, N, 0, c0, Cin, H, W, ksize, flip, dst, act_dtype, ld, stream, ones_col);
}

extern "C" int glowk_rows_to_nchw(const void* rows, int act_dtype, int64_t ld, float* dst, int64_t N, int64_t C,
                                  int64_t HW, void* stream) {
  GLOWK_CHECK_ARG(rows && dst && C <= ld, "glowk_rows_to_nchw: bad arguments");
  const int64_t total = N * C * HW;
  if (total == 0) return GLOWK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div(total, 256);
  if (act_dtype == GLOWK_BF16) rows_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)rows, ld, dst, total, C, HW);
  else if (act_dtype == GLOWK_F32) rows_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float*)rows, ld, dst, total, C, HW);
  else return fail(GLOWK_EINVAL, "glowk_rows_to_nchw: bad act_dtype %d", act_dtype);
  GLOWK_CHECK_LAUNCH("glowk_rows_to_nchw");
  return GLOWK_OK;
}

extern "C" int glowk_tapsum_to_nchw(const float* P, int64_t ldp, float* dst, int64_t N, int64_t Ctot, int64_t c0,
                                    int64_t C, int64_t H, int64_t W, int flip, int accumulate, void* stream) {
  GLOWK_CHECK_ARG(P && dst && ldp >= 9 * C && c0 + C <= Ctot, "glowk_tapsum_to_nchw: bad arguments");
  const int64_t total = N * C * H * W;
  if (total == 0) return GLOWK_OK;
  tapsum_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(P, ldp, dst, total, Ctot, c0, (int)C, (int)H, (int)W, flip, accumulate);
  GLOWK_CHECK_LAUNCH("glowk_tapsum_to_nchw");
  return GLOWK_OK;
}

extern "C" int glowk_pack_conv_weight(const float* w, int64_t O, int64_t I, int ksize, int layout, void* dst,
                                      int act_dtype, int64_t rows, int64_t ld, void* stream) {
  GLOWK_CHECK_ARG(w && dst, "glowk_pack_conv_weight: null pointer");
  GLOWK_CHECK_ARG(ksize == 1 || ksize == 3, "glowk_pack_conv_weight: kernel size must be 1 or 3");
  GLOWK_CHECK_ARG(layout >= 0 && layout <= 3, "glowk_pack_conv_weight: bad layout");
  const int64_t T2 = ksize * ksize;
  const int64_t need_r[4] = {O, T2 * O, T2 * I, I}, need_c[4] = {T2 * I, I, O, T2 * O};
  GLOWK_CHECK_ARG(rows >= need_r[layout] && ld >= need_c[layout], "glowk_pack_conv_weight: dst too small");
  const int64_t total = rows * ld;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div(total, 256);
  if (act_dtype == GLOWK_BF16) pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(w, (int)O, (int)I, ksize, layout, (__nv_bfloat16*)dst, rows, ld);
  else if (act_dtype == GLOWK_F32) pack_weight_kernel<float><<<grid, 256, 0, st>>>(w, (int)O, (int)I, ksize, layout, (float*)dst, rows, ld);
  else return fail(GLOWK_EINVAL, "glowk_pack_conv_weight: bad act_dtype %d", act_dtype);
  GLOWK_CHECK_LAUNCH("glowk_pack_conv_weight");
  return GLOWK_OK;
}

static inline int coupling_threads(int64_t HW) {
  int t = 256;
  while (t > 32 && t / 2 >= HW) t /= 2;
  return t;
}

extern "C" int64_t glowk_coupling_nblk(int64_t HW) { return ceil_div(HW, coupling_threads(HW)); }

extern "C" int glowk_coupling(const float* P, int64_t ldp, const float* bias3, const float* logs3,
                              float logscale_factor, float* z, float* partials, float* h_save, int64_t N,
                              int64_t C, int64_t H, int64_t W, int affine, int reverse, void* stream) {
  GLOWK_CHECK_ARG(P && bias3 && logs3 && z, "glowk_coupling: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C % 2 == 0, "glowk_coupling: channel count must be even (model.py:169)");
  const int64_t Cout = affine ? C : C / 2;
  GLOWK_CHECK_ARG(ldp >= 9 * Cout && ldp % 2 == 0, "glowk_coupling: ldp=%lld too small for 9*Cout=%lld", (long long)ldp, (long long)(9 * Cout));
  GLOWK_CHECK_ARG(!affine || partials, "glowk_coupling: affine coupling needs a partials buffer");
  GLOWK_CHECK_ARG(N <= 65535, "glowk_coupling: batch too large for one launch");
  if (N == 0) return GLOWK_OK;
  const int64_t HW = H * W;
  const int threads = coupling_threads(HW);
  dim3 grid((unsigned)ceil_div(HW, threads), (unsigned)N);
  coupling_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(P, ldp, bias3, logs3, logscale_factor, z,
                                                               affine ? partials : nullptr, h_save, (int)C,
                                                               (int)H, (int)W, affine, reverse);
  GLOWK_CHECK_LAUNCH("glowk_coupling");
  return GLOWK_OK;
}

extern "C" int glowk_logdet_finish(const float* logdet_in, float* logdet_out, const float* logs, int64_t C,
                                   float logscale_factor, const float* logabsdet, const float* partials,
                                   int64_t nblk, int64_t HW, float sign, int64_t N, void* stream) {
  GLOWK_CHECK_ARG(logdet_out, "glowk_logdet_finish: null output");
  if (N == 0) return GLOWK_OK;
  logdet_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logdet_in, logdet_out, logs, (int)C, logscale_factor,
                                                             logabsdet, partials, (int)nblk, (float)HW, sign, N);
  GLOWK_CHECK_LAUNCH("glowk_logdet_finish");
  return GLOWK_OK;
}

extern "C" int glowk_gaussian_logp(const float* h, int64_t ldh, const float* x, int64_t N, int64_t C, int64_t HW,
                                   int64_t c0, int64_t Cz, const float* logdet_in, float* logdet_out, void* stream) {
  GLOWK_CHECK_ARG(x && logdet_out, "glowk_gaussian_logp: null pointer");
  GLOWK_CHECK_ARG(c0 >= 0 && c0 + Cz <= C, "glowk_gaussian_logp: channel window out of range");
  GLOWK_CHECK_ARG(!h || (ldh >= 2 * Cz && ldh % 2 == 0), "glowk_gaussian_logp: ldh too small");
  if (N == 0) return GLOWK_OK;
  gaussian_logp_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(h, ldh, x, C, HW, c0, Cz, logdet_in, logdet_out);
  GLOWK_CHECK_LAUNCH("glowk_gaussian_logp");
  return GLOWK_OK;
}

extern "C" int glowk_split2d_sample(const float* h, int64_t ldh, const float* z1, const float* eps, float* out,
                                    int64_t N, int64_t Chalf, int64_t HW, void* stream) {
  GLOWK_CHECK_ARG(h && z1 && eps && out, "glowk_split2d_sample: null pointer");
  GLOWK_CHECK_ARG(ldh >= 2 * Chalf && ldh % 2 == 0, "glowk_split2d_sample: ldh too small");
  const int64_t total = N * 2 * Chalf * HW;
  if (total == 0) return GLOWK_OK;
  split2d_sample_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(h, ldh, z1, eps, out, total, Chalf, HW);
  GLOWK_CHECK_LAUNCH("glowk_split2d_sample");
  return GLOWK_OK;
}

extern "C" int glowk_nll_head(const float* z, const float* ld, float c0, float denom, int64_t N, int64_t D, float* nll,
                              float* loss, void* ticket, void* stream) {
  GLOWK_CHECK_ARG(z && nll && (!loss || ticket), "glowk_nll_head: null pointer");
  GLOWK_CHECK_ARG(denom > 0.f && N < (1ll << 31), "glowk_nll_head: bad arguments");
  if (N == 0) return GLOWK_OK;
  nll_head_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(z, D, ld, c0, denom, (int)N, nll, loss, (unsigned*)ticket);
  GLOWK_CHECK_LAUNCH("glowk_nll_head");
  return GLOWK_OK;
}

extern "C" int glowk_nll_head_bwd(const float* z, const float* g_loss, const float* g_nll, const float* dz_in,
                                  float denom, int64_t N, int64_t D, float* dz, float* dld, void* stream) {
  GLOWK_CHECK_ARG(z && dz && dld, "glowk_nll_head_bwd: null pointer");
  GLOWK_CHECK_ARG(denom > 0.f && N <= 65535 && D > 0, "glowk_nll_head_bwd: bad arguments");
  if (N == 0) return GLOWK_OK;
  const dim3 grid((unsigned)ceil_div(D, 256), (unsigned)N);
  nll_head_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, g_loss, g_nll, dz_in, denom, (int)N, D, dz, dld);
  GLOWK_CHECK_LAUNCH("glowk_nll_head_bwd");
  return GLOWK_OK;
}
