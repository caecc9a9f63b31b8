// Backward kernels of the flow layers (training path) and the fused optimizer step, sm_100a.
// The gradient formulas are the analytic derivatives of the forward kernels in flow_kernels.cu;
// parity is checked against autograd through the CPU oracle (tests/test_gpu_backward.py).
#include "common.cuh"

namespace glowk {

// ------------------------------------------------------------------------------------------
// Coupling backward (model.py:105-115).  Inputs: y (step output, y2 = y[:, C/2:]), h rows (saved
// Conv2dZeros output), dy, dld[n] (grad of the loss wrt this step's per-sample logdet).
//   affine: shift=h[2j], scale=sigmoid(h[2j+1]+2), z2+shift = y2/scale
//       dz2 = dy2*scale ; dh[2j] = dy2*scale ; dh[2j+1] = (dy2*(y2/scale) + dld/scale)*scale*(1-scale)
//   additive: dz2 = dy2 ; dh[j] = dy2
// Conv2dZeros: h=(u+b3)*exp(f*logs3) -> du = dh*exp(f*logs3); db3 += sum du; dlogs3 += f*sum dh*h.
// Writes dz (all C channels: dz1 = dy1, to be completed by the conv dgrad), du rows [P][Cout].
// ------------------------------------------------------------------------------------------
__global__ void coupling_bwd_kernel(const float* __restrict__ y, const float* __restrict__ hrows,
                                    const float* __restrict__ dy, const float* __restrict__ dld,
                                    const float* __restrict__ logs3, float f, float* __restrict__ dz,
                                    float* __restrict__ du, float* __restrict__ dlogs3,
                                    float* __restrict__ dbias3, int C, int HW, int affine) {
  extern __shared__ float s_acc[];   // [2][Cout] block partials
  const int Ch = C / 2, Cout = affine ? C : Ch;
  const int64_t n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const float g = dld ? dld[n] : 0.f;
  const int lane = threadIdx.x & 31;
  const bool active = p < HW;
  const float* hr = hrows + (n * HW + (active ? p : 0)) * (int64_t)Cout;
  float* dur = du + (n * HW + (active ? p : 0)) * (int64_t)Cout;
  for (int j = 0; j < Ch; ++j) {
    float dh0 = 0.f, dh1 = 0.f, h0 = 0.f, h1 = 0.f;
    if (active) {
      const int64_t o1 = (n * C + j) * HW + p, o2 = (n * C + Ch + j) * HW + p;
      dz[o1] = dy[o1];
      const float dy2 = dy[o2];
      if (affine) {
        h0 = hr[2 * j]; h1 = hr[2 * j + 1];
        const float scale = 1.f / (1.f + expf(-(h1 + 2.f)));
        const float zs = y[o2] / scale;            // z2 + shift
        dz[o2] = dy2 * scale;
        dh0 = dy2 * scale;
        dh1 = (dy2 * zs + g / scale) * scale * (1.f - scale);
        dur[2 * j] = dh0 * expf(logs3[2 * j] * f);
        dur[2 * j + 1] = dh1 * expf(logs3[2 * j + 1] * f);
      } else {
        h0 = hr[j];
        dz[o2] = dy2;
        dh0 = dy2;
        dur[j] = dh0 * expf(logs3[j] * f);
      }
    }
    // per-channel reductions over the block's pixels
    float a0 = warp_sum(dh0 * h0), b0 = warp_sum(dh0);
    if (affine) {
      float a1 = warp_sum(dh1 * h1), b1 = warp_sum(dh1);
      if (lane == 0) {
        atomicAdd(&s_acc[2 * j], a0); atomicAdd(&s_acc[Cout + 2 * j], b0);
        atomicAdd(&s_acc[2 * j + 1], a1); atomicAdd(&s_acc[Cout + 2 * j + 1], b1);
      }
    } else if (lane == 0) {
      atomicAdd(&s_acc[j], a0); atomicAdd(&s_acc[Cout + j], b0);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
    atomicAdd(dlogs3 + c, f * s_acc[c]);
    atomicAdd(dbias3 + c, expf(logs3[c] * f) * s_acc[Cout + c]);
  }
}

// ------------------------------------------------------------------------------------------
// Split2d backward (module.py:526-530): logdet += sum logp(z2; mean, logs), (mean, logs) = cross(h).
//   dz2 = -g (z2-mean)/e^{2logs};  dmean = g (z2-mean)/e^{2logs};  dlogs = g ((z2-mean)^2/e^{2logs} - 1)
// dx[:, :C/2] = dz1_out (conv dgrad is accumulated afterwards), dx[:, C/2:] = dz2, du = dh*exp(f*logs_p).
// ------------------------------------------------------------------------------------------
__global__ void split2d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ hrows, int64_t ldh,
                                   const float* __restrict__ dz1, int64_t dz1_sN, const float* __restrict__ dld,
                                   const float* __restrict__ logs_p, float f, float* __restrict__ dx,
                                   float* __restrict__ du, int64_t ldu, float* __restrict__ dlogs_p,
                                   float* __restrict__ dbias_p, int C, int HW) {
  extern __shared__ float s_acc[];   // [2][C]
  const int Ch = C / 2;
  const int64_t n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const float g = dld[n];
  const int lane = threadIdx.x & 31;
  const bool active = p < HW;
  const float* hr = hrows + (n * HW + (active ? p : 0)) * ldh;
  float* dur = du + (n * HW + (active ? p : 0)) * ldu;
  for (int j = 0; j < Ch; ++j) {
    float dm = 0.f, dl = 0.f, mean = 0.f, lg = 0.f;
    if (active) {
      const int64_t o1 = (n * C + j) * HW + p, o2 = (n * C + Ch + j) * HW + p;
      dx[o1] = dz1 ? dz1[n * dz1_sN + (int64_t)j * HW + p] : 0.f;
      mean = hr[2 * j]; lg = hr[2 * j + 1];
      const float d = x[o2] - mean;
      const float iv = 1.f / expf(2.f * lg);
      dx[o2] = -g * d * iv;
      dm = g * d * iv;
      dl = g * (d * d * iv - 1.f);
      dur[2 * j] = dm * expf(logs_p[2 * j] * f);
      dur[2 * j + 1] = dl * expf(logs_p[2 * j + 1] * f);
    }
    const float a0 = warp_sum(dm * mean), b0 = warp_sum(dm), a1 = warp_sum(dl * lg), b1 = warp_sum(dl);
    if (lane == 0) {
      atomicAdd(&s_acc[2 * j], a0); atomicAdd(&s_acc[C + 2 * j], b0);
      atomicAdd(&s_acc[2 * j + 1], a1); atomicAdd(&s_acc[C + 2 * j + 1], b1);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dlogs_p + c, f * s_acc[c]);
    atomicAdd(dbias_p + c, expf(logs_p[c] * f) * s_acc[C + c]);
  }
}

// ------------------------------------------------------------------------------------------
// ActNorm + channel-mix backward (model.py:94-103).  a = (x+b)*s, z = W a  (or z[o] = a[idx[o]]).
//   da = W^T dz ; dx = da*s ; db += sum da*s ; dlogs += f*sum da*a ; dW += sum_p dz a^T
// Tile of TP=128 pixels staged in smem (rows padded by 4 floats so both the float4 row reads of phase 1
// and the column-strided reads of the dW phase are bank-conflict free).
// ------------------------------------------------------------------------------------------
constexpr int MB_TP = 128;
constexpr int MB_LD = MB_TP + 4;

__global__ void __launch_bounds__(256)
mix_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dz, const float* __restrict__ w,
               const int64_t* __restrict__ idx, const float* __restrict__ bias, const float* __restrict__ logs,
               float f, float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dlogs,
               float* __restrict__ dbias, int64_t NP, int C, int64_t HW) {
  extern __shared__ __align__(16) float smem[];
  float* as = smem;                         // [C][MB_LD]  a = actnorm(x)
  float* ds = as + (size_t)C * MB_LD;       // [C][MB_LD]  dz, later da
  float* ws = ds + (size_t)C * MB_LD;       // [C][C]      W (mix only)
  float* sc = ws + (w ? (size_t)C * C : 0); // [C] scale
  float* bs = sc + C;                       // [C] bias
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int64_t g0 = (int64_t)blockIdx.x * MB_TP;
  const bool has_an = bias != nullptr;
  for (int c = tid; c < C; c += nthr) {
    sc[c] = has_an ? expf(logs[c] * f) : 1.f;
    bs[c] = has_an ? bias[c] : 0.f;
  }
  if (w) for (int e = tid; e < C * C; e += nthr) ws[e] = w[e];
  __syncthreads();
  constexpr int Q = MB_TP / 4;
  for (int e = tid; e < C * Q; e += nthr) {
    const int c = e / Q, q = e - c * Q;
    const int64_t g = g0 + (int64_t)q * 4;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
    if (g < NP) {
      const int64_t n = g / HW, p = g - n * HW;
      xv = ld_stream4(x + (n * C + c) * HW + p);
      dv = ld_stream4(dz + (n * C + c) * HW + p);
      const float b = bs[c], s = sc[c];
      xv.x = (xv.x + b) * s; xv.y = (xv.y + b) * s; xv.z = (xv.z + b) * s; xv.w = (xv.w + b) * s;
    }
    *reinterpret_cast<float4*>(as + (size_t)c * MB_LD + q * 4) = xv;
    *reinterpret_cast<float4*>(ds + (size_t)c * MB_LD + q * 4) = dv;
  }
  __syncthreads();

  // ---- dW[o][i] += sum_p dz[o][p] a[i][p]
  if (w) {
    for (int e = tid; e < C * C; e += nthr) {
      const int o = e / C, i = e - o * C;
      const float* dr = ds + (size_t)o * MB_LD;
      const float* ar = as + (size_t)i * MB_LD;
      float acc = 0.f;
#pragma unroll 4
      for (int p = 0; p < MB_TP; p += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(dr + p);
        const float4 a4 = *reinterpret_cast<const float4*>(ar + p);
        acc = fmaf(d4.x, a4.x, acc); acc = fmaf(d4.y, a4.y, acc);
        acc = fmaf(d4.z, a4.z, acc); acc = fmaf(d4.w, a4.w, acc);
      }
      atomicAdd(dw + e, acc);
    }
    __syncthreads();
  }

  // ---- da = W^T dz (thread = pixel quad x input channel i), dx = da * s, channel reductions
  const int q = tid % Q;
  const int ngrp = nthr / Q;
  const int64_t g = g0 + (int64_t)q * 4;
  const bool active = g < NP;
  const int64_t n = active ? g / HW : 0, p = active ? g - n * HW : 0;
  for (int i = tid / Q; i < C; i += ngrp) {   // all lanes of a warp share i
    float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w) {
      for (int o = 0; o < C; ++o) {
        const float wv = ws[o * C + i];
        const float4 d4 = *reinterpret_cast<const float4*>(ds + (size_t)o * MB_LD + q * 4);
        da.x = fmaf(wv, d4.x, da.x); da.y = fmaf(wv, d4.y, da.y);
        da.z = fmaf(wv, d4.z, da.z); da.w = fmaf(wv, d4.w, da.w);
      }
    } else {
      // z[o] = a[idx[o]]  =>  da[i] = dz[o] for the o with idx[o] == i
      int o = 0;
      for (int k = 0; k < C; ++k) if ((int)idx[k] == i) o = k;
      da = *reinterpret_cast<const float4*>(ds + (size_t)o * MB_LD + q * 4);
    }
    const float s = sc[i];
    if (active) st_stream4(dx + (n * C + i) * HW + p, make_float4(da.x * s, da.y * s, da.z * s, da.w * s));
    if (has_an) {
      const float4 a4 = *reinterpret_cast<const float4*>(as + (size_t)i * MB_LD + q * 4);
      float sg = da.x + da.y + da.z + da.w;
      float sga = da.x * a4.x + da.y * a4.y + da.z * a4.z + da.w * a4.w;
      sg = warp_sum(sg); sga = warp_sum(sga);
      if ((tid & 31) == 0) { atomicAdd(dbias + i, sg * s); atomicAdd(dlogs + i, f * sga); }
    }
  }
}

// Fallback of mix_bwd_kernel for channel counts whose tiles do not fit in shared memory (C >= ~150: levels 5-6 of
// the 256x256 L=6 configuration).  Correct for any C, not tuned: (a) one thread per (n, i, p) for dx and the
// per-channel sums, (b) one thread per (o, i) for dW.
__global__ void mix_bwd_wide_dx_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                       const float* __restrict__ w, const int64_t* __restrict__ idx,
                                       const float* __restrict__ bias, const float* __restrict__ logs, float f,
                                       float* __restrict__ dx, float* __restrict__ dlogs, float* __restrict__ dbias,
                                       int64_t total, int C, int64_t HW) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t p = e % HW;
  const int i = (int)((e / HW) % C);
  const int64_t n = e / (HW * C);
  const bool has_an = bias != nullptr;
  const float s = has_an ? expf(logs[i] * f) : 1.f;
  const float a = has_an ? (x[e] + bias[i]) * s : x[e];
  const float* dzb = dz + n * C * HW + p;
  float da = 0.f;
  if (w) {
    for (int o = 0; o < C; ++o) da = fmaf(w[(int64_t)o * C + i], dzb[(int64_t)o * HW], da);
  } else {
    int o = 0;
    for (int k = 0; k < C; ++k) if ((int)idx[k] == i) o = k;
    da = dzb[(int64_t)o * HW];
  }
  dx[e] = da * s;
  if (has_an) { atomicAdd(dbias + i, da * s); atomicAdd(dlogs + i, f * da * a); }
}

__global__ void mix_bwd_wide_dw_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                       const float* __restrict__ bias, const float* __restrict__ logs, float f,
                                       float* __restrict__ dw, int64_t N, int C, int64_t HW) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)C * C) return;
  const int o = (int)(e / C), i = (int)(e - (int64_t)o * C);
  const bool has_an = bias != nullptr;
  const float s = has_an ? expf(logs[i] * f) : 1.f, b = has_an ? bias[i] : 0.f;
  float acc = 0.f;
  for (int64_t n = 0; n < N; ++n) {
    const float* xr = x + (n * C + i) * HW;
    const float* dr = dz + (n * C + o) * HW;
    for (int64_t p = 0; p < HW; ++p) acc = fmaf(dr[p], has_an ? (xr[p] + b) * s : xr[p], acc);
  }
  dw[e] += acc;
}

// Gradient of the sample-independent logdet terms wrt their parameters (module.py:78-82, 357-363):
//   logdet[n] += HW*(sum_c f*logs_c + log|det W|)  =>  dlogs_c += f*HW*G,  dW += HW*G*W^-T,  G = sum_n dld[n]
__global__ void logdet_param_grad_kernel(const float* __restrict__ dld, int64_t N, float hw, float f,
                                         float* __restrict__ dlogs, int C, const float* __restrict__ winv,
                                         float* __restrict__ dw) {
  __shared__ float red[32];
  __shared__ float s_g;
  float a = 0.f;
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) a += dld[n];
  const float tot = block_sum(a, red);
  if (threadIdx.x == 0) s_g = tot;
  __syncthreads();
  const float G = s_g * hw;
  // every CTA re-derives G (N floats) and owns a slice of dW; CTA 0 also updates dlogs.  (One CTA for the whole
  // C x C transpose-add took 180 us per call at C = 384.)
  if (dlogs && blockIdx.x == 0) for (int c = threadIdx.x; c < C; c += blockDim.x) dlogs[c] += f * G;
  if (dw) for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < C * C; e += gridDim.x * blockDim.x) {
    const int o = e / C, i = e - o * C;
    dw[e] += G * winv[i * C + o];
  }
}

// Inverse of pack_weight_kernel: grad[o][i][tap] (+)= src[...] for the four packed layouts.
__global__ void unpack_weight_grad_kernel(const float* __restrict__ src, int64_t ld, int O, int I, int ks,
                                          int layout, float* __restrict__ grad, int accumulate) {
  const int T2 = ks * ks;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)O * I * T2) return;
  const int tap = (int)(e % T2);
  const int i = (int)((e / T2) % I);
  const int o = (int)(e / ((int64_t)T2 * I));
  int64_t s;
  if (layout == 0) s = (int64_t)o * ld + (int64_t)tap * I + i;
  else if (layout == 1) s = ((int64_t)tap * O + o) * ld + i;
  else if (layout == 2) s = ((int64_t)tap * I + i) * ld + o;
  else s = (int64_t)i * ld + (int64_t)tap * O + o;
  grad[e] = accumulate ? grad[e] + src[s] : src[s];
}

// ------------------------------------------------------------------------------------------
// Fused optimizer step over a flat fp32 parameter arena (trainer.py:142-150 + torch.optim.Adam):
//   pass 1: g = clamp(g, -clip, clip) in place, per-CTA sum of squares (deterministic two-stage)
//   pass 2: total norm, coef = min(1, max_norm/(norm+1e-6))
//   pass 3: g *= coef ; Adam moment update ; parameter update
// ------------------------------------------------------------------------------------------
__global__ void clip_sumsq_kernel(float* __restrict__ g, int64_t n, float clip, float* __restrict__ partials) {
  __shared__ float red[32];
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 v = *reinterpret_cast<float4*>(g + i);
      if (clip > 0.f) {
        v.x = fminf(fmaxf(v.x, -clip), clip); v.y = fminf(fmaxf(v.y, -clip), clip);
        v.z = fminf(fmaxf(v.z, -clip), clip); v.w = fminf(fmaxf(v.w, -clip), clip);
        *reinterpret_cast<float4*>(g + i) = v;
      }
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (int64_t k = i; k < n; ++k) {
        float v = g[k];
        if (clip > 0.f) { v = fminf(fmaxf(v, -clip), clip); g[k] = v; }
        acc += v * v;
      }
    }
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

__global__ void norm_finish_kernel(const float* __restrict__ partials, int nparts, float max_norm,
                                   float* __restrict__ out /* [0]=norm, [1]=coef */) {
  __shared__ double red[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) a += (double)partials[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float norm = (float)sqrt(t);
    out[0] = norm;
    float coef = 1.f;
    if (max_norm > 0.f) { coef = max_norm / (norm + 1e-6f); if (coef > 1.f) coef = 1.f; }
    out[1] = coef;
  }
}

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, const float* __restrict__ norm_coef,
                            const float* __restrict__ sched, float lr_scalar, float beta1, float beta2, float eps,
                            float bc1_h, float bc2_sqrt_h) {
  const float coef = norm_coef ? norm_coef[1] : 1.f;
  // sched (device, [lr, 1-beta1^t, sqrt(1-beta2^t)]) lets a captured CUDA graph follow the LR schedule
  const float lr = sched ? sched[0] : lr_scalar;
  const float bc1 = sched ? sched[1] : bc1_h;
  const float bc2_sqrt = sched ? sched[2] : bc2_sqrt_h;
  const float step_size = lr / bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    g[i] = gi;
    const float mi = m[i] * beta1 + (1.f - beta1) * gi;
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// torch.optim.Adamax (the reference's other optimizer choice, network/builder.py:10-13):
//   m = beta1*m + (1-beta1)*g;  u = max(beta2*u, |g| + eps);  p -= lr/(1-beta1^t) * m/u
__global__ void adamax_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                              float* __restrict__ u, int64_t n, const float* __restrict__ norm_coef,
                              const float* __restrict__ sched, float lr_scalar, float beta1, float beta2, float eps,
                              float bc1_h) {
  const float coef = norm_coef ? norm_coef[1] : 1.f;
  const float lr = sched ? sched[0] : lr_scalar;
  const float bc1 = sched ? sched[1] : bc1_h;
  const float clr = lr / bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    g[i] = gi;
    const float mi = m[i] * beta1 + (1.f - beta1) * gi;
    const float ui = fmaxf(u[i] * beta2, fabsf(gi) + eps);
    m[i] = mi; u[i] = ui;
    p[i] = p[i] - clr * (mi / ui);
  }
}

}  // namespace glowk

using namespace glowk;

static inline int pix_threads(int64_t HW) {
  int t = 256;
  while (t > 32 && t / 2 >= HW) t /= 2;
  return t;
}

extern "C" int glowk_coupling_bwd(const float* y, const float* hrows, const float* dy, const float* dld,
                                  const float* logs3, float logscale_factor, float* dz, float* du, float* dlogs3,
                                  float* dbias3, int64_t N, int64_t C, int64_t H, int64_t W, int affine, void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(y && hrows && dy && logs3 && dz && du && dlogs3 && dbias3, "glowk_coupling_bwd: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C % 2 == 0 && N <= 65535, "glowk_coupling_bwd: bad shape");
  const int64_t HW = H * W;
  const int threads = pix_threads(HW);
  const int64_t Cout = affine ? C : C / 2;
  dim3 grid((unsigned)ceil_div(HW, threads), (unsigned)N);
  coupling_bwd_kernel<<<grid, threads, 2 * Cout * sizeof(float), (cudaStream_t)stream>>>(
      y, hrows, dy, dld, logs3, logscale_factor, dz, du, dlogs3, dbias3, (int)C, (int)HW, affine);
  GLOWK_CHECK_LAUNCH("glowk_coupling_bwd");
  return GLOWK_OK;
}

extern "C" int glowk_split2d_bwd(const float* x, const float* hrows, int64_t ldh, const float* dz1, int64_t dz1_sN,
                                 const float* dld, const float* logs_p, float logscale_factor, float* dx, float* du,
                                 int64_t ldu, float* dlogs_p, float* dbias_p, int64_t N, int64_t C, int64_t HW,
                                 void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && hrows && dld && logs_p && dx && du && dlogs_p && dbias_p, "glowk_split2d_bwd: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C % 2 == 0 && ldh >= C && ldu >= C && N <= 65535, "glowk_split2d_bwd: bad shape");
  const int threads = pix_threads(HW);
  dim3 grid((unsigned)ceil_div(HW, threads), (unsigned)N);
  split2d_bwd_kernel<<<grid, threads, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      x, hrows, ldh, dz1, dz1_sN, dld, logs_p, logscale_factor, dx, du, ldu, dlogs_p, dbias_p, (int)C, (int)HW);
  GLOWK_CHECK_LAUNCH("glowk_split2d_bwd");
  return GLOWK_OK;
}

extern "C" int glowk_actnorm_mix_bwd(const float* x, const float* dz, const float* w, const int64_t* idx,
                                     const float* bias, const float* logs, float logscale_factor, float* dx, float* dw,
                                     float* dlogs, float* dbias, int64_t N, int64_t C, int64_t HW, void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && dz && dx, "glowk_actnorm_mix_bwd: null pointer");
  GLOWK_CHECK_ARG((w != nullptr) != (idx != nullptr), "glowk_actnorm_mix_bwd: exactly one of w / idx");
  GLOWK_CHECK_ARG(!w || dw, "glowk_actnorm_mix_bwd: dw required with w");
  GLOWK_CHECK_ARG((bias != nullptr) == (logs != nullptr) && (!bias || (dlogs && dbias)), "glowk_actnorm_mix_bwd: actnorm args");
  const size_t smem = sizeof(float) * (2 * (size_t)C * MB_LD + (w ? (size_t)C * C : 0) + 2 * (size_t)C);
  if (smem > 220 * 1024 || HW % 4 != 0) {      // wide / odd shapes: generic fallback kernels
    const int64_t total = N * C * HW;
    cudaStream_t st = (cudaStream_t)stream;
    mix_bwd_wide_dx_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(x, dz, w, idx, bias, logs, logscale_factor, dx, dlogs, dbias,
                                                                         total, (int)C, HW);
    if (w) mix_bwd_wide_dw_kernel<<<(unsigned)ceil_div(C * C, 256), 256, 0, st>>>(x, dz, bias, logs, logscale_factor, dw, N, (int)C, HW);
    GLOWK_CHECK_LAUNCH("glowk_actnorm_mix_bwd(wide)");
    return GLOWK_OK;
  }
  if (smem > 48 * 1024) GLOWK_CUDA(cudaFuncSetAttribute(mix_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t NP = N * HW;
  int groups = (int)C < 8 ? (int)C : 8;
  mix_bwd_kernel<<<(unsigned)ceil_div(NP, MB_TP), 32 * groups, smem, (cudaStream_t)stream>>>(
      x, dz, w, idx, bias, logs, logscale_factor, dx, dw, dlogs, dbias, NP, (int)C, HW);
  GLOWK_CHECK_LAUNCH("glowk_actnorm_mix_bwd");
  return GLOWK_OK;
}

extern "C" int glowk_logdet_param_grad(const float* dld, int64_t N, int64_t HW, float logscale_factor, float* dlogs,
                                       int64_t C, const float* winv, float* dw, void* stream) {
  GLOWK_CHECK_ARG(dld && (dlogs || dw) && (!dw || winv), "glowk_logdet_param_grad: null pointer");
  const unsigned ctas = dw ? (unsigned)ceil_div(C * C, 1024) : 1u;
  logdet_param_grad_kernel<<<ctas < 1 ? 1 : (ctas > 296 ? 296 : ctas), 256, 0, (cudaStream_t)stream>>>(dld, N, (float)HW, logscale_factor, dlogs, (int)C, winv, dw);
  GLOWK_CHECK_LAUNCH("glowk_logdet_param_grad");
  return GLOWK_OK;
}

extern "C" int glowk_unpack_weight_grad(const float* src, int64_t ld, int64_t O, int64_t I, int ksize, int layout,
                                        float* grad, int accumulate, void* stream) {
  GLOWK_CHECK_ARG(src && grad && layout >= 0 && layout <= 3 && (ksize == 1 || ksize == 3), "glowk_unpack_weight_grad: bad arguments");
  const int64_t total = O * I * ksize * ksize;
  unpack_weight_grad_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(src, ld, (int)O, (int)I, ksize, layout, grad, accumulate);
  GLOWK_CHECK_LAUNCH("glowk_unpack_weight_grad");
  return GLOWK_OK;
}

extern "C" int64_t glowk_optim_workspace_floats(void) { return 1024 + 8; }

extern "C" int glowk_optim_clip_norm(float* grads, int64_t n, float clip_value, float max_norm, float* workspace,
                                     void* stream) {
  GLOWK_CHECK_ARG(grads && workspace && n > 0, "glowk_optim_clip_norm: bad arguments");
  GLOWK_CHECK_ARG(((uintptr_t)grads) % 16 == 0, "glowk_optim_clip_norm: grads must be 16-byte aligned");
  int blocks = (int)ceil_div(n, 256 * 4 * 4);
  if (blocks > 1024) blocks = 1024;
  clip_sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grads, n, clip_value, workspace + 8);
  norm_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(workspace + 8, blocks, max_norm, workspace);
  GLOWK_CHECK_LAUNCH("glowk_optim_clip_norm");
  return GLOWK_OK;
}

extern "C" int glowk_optim_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                const float* norm_coef, const float* sched_dev, float lr, float beta1, float beta2,
                                float eps, int64_t step, void* stream) {
  GLOWK_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "glowk_optim_adam: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int blocks = (int)ceil_div(n, 256 * 4);
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, norm_coef, sched_dev, lr,
                                                        beta1, beta2, eps, (float)bc1, (float)sqrt(bc2));
  GLOWK_CHECK_LAUNCH("glowk_optim_adam");
  return GLOWK_OK;
}

extern "C" int glowk_optim_adamax(float* params, float* grads, float* exp_avg, float* exp_inf, int64_t n,
                                  const float* norm_coef, const float* sched_dev, float lr, float beta1, float beta2,
                                  float eps, int64_t step, void* stream) {
  GLOWK_CHECK_ARG(params && grads && exp_avg && exp_inf && n > 0 && step >= 1, "glowk_optim_adamax: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  int blocks = (int)ceil_div(n, 256 * 4);
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  adamax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_inf, n, norm_coef, sched_dev, lr,
                                                          beta1, beta2, eps, (float)bc1);
  GLOWK_CHECK_LAUNCH("glowk_optim_adamax");
  return GLOWK_OK;
}

// Noam learning-rate schedule (misc/lr_scheduler.py:18-37) + Adam bias corrections from a DEVICE step counter, so a
// captured optimizer graph follows the schedule without reading host memory (a pinned-host source can be overwritten
// by the CPU before the GPU executes the copy).  One thread; double precision like the host arithmetic it replaces.
__global__ void optim_schedule_kernel(long long* __restrict__ step, float* __restrict__ sched, double base_lr,
                                      double warmup, double min_lr, double beta1, double beta2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long gs = *step;                       // completed iterations
  const double t = (double)gs + 1.0;
  double lr = base_lr;
  if (warmup > 0.0) {
    lr = base_lr * sqrt(warmup) * fmin(t * pow(warmup, -1.5), 1.0 / sqrt(t));
    if ((double)gs >= warmup && min_lr >= 0.0) lr = fmax(lr, min_lr);
  }
  sched[0] = (float)lr;
  sched[1] = (float)(1.0 - pow(beta1, t));
  sched[2] = (float)sqrt(1.0 - pow(beta2, t));
  *step = gs + 1;
}

extern "C" int glowk_optim_schedule(void* step_dev, float* sched_dev, float base_lr, int64_t warmup_steps, float min_lr,
                                    float beta1, float beta2, void* stream) {
  GLOWK_CHECK_ARG(step_dev && sched_dev && warmup_steps >= 0, "glowk_optim_schedule: bad arguments");
  optim_schedule_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((long long*)step_dev, sched_dev, (double)base_lr,
                                                            (double)warmup_steps, (double)min_lr, (double)beta1, (double)beta2);
  GLOWK_CHECK_LAUNCH("glowk_optim_schedule");
  return GLOWK_OK;
}
