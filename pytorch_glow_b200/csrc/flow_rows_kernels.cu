// Glow flow kernels on a PIXEL-MAJOR flow state ("rows": x[p][c], p = (n*H + y)*W + x, c contiguous), sm_100a.
//
// FlowModel.encode / decode keep the flow state in this layout between the NCHW tensors of the
// reference API (network/model.py:263-294): it is the layout of every coupling-network GEMM operand,
// so ActNorm + 1x1 mix, the im2col of z1, the tap gather-sum + coupling and all their adjoints read
// and write whole contiguous pixels (48..384 bytes) instead of C strided planes.  The arithmetic of
// each element (operation order included) is that of the NCHW kernels in flow_kernels.cu /
// flow_bwd_kernels.cu, so the two layouts agree bit for bit except for the order of the per-sample
// and per-channel reductions.
#include "common.cuh"

namespace glowk {

constexpr int ROWS_MAX_C = 96;   // shared-memory tiles of the mix kernels are sized for C <= 96 (levels 1..4 of every config)
// The coupling / Split2d kernels (no C x C weight in shared memory) run up to 384 channels -- levels 5 and 6 of the
// 256x256 L=6 configuration -- where ActNorm + the 1x1 conv and their adjoint are done by the fp32 GEMMs instead
// (rows_path.py: _mix_wide_*; glowk_rows_actnorm_bwd below).
constexpr int ROWS_MAX_C_WIDE = 384;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------
// ActNorm + channel mix / permutation (model.py:94-103 fwd, 142-152 rev).
// block = (G = C/4 output-channel groups, PPB pixels): thread (og, slot) produces 4 output channels of one
// pixel; the G lanes of a pixel read the same 16-byte chunks of x (one transaction) and W^T from shared memory.
// ------------------------------------------------------------------------------------------
// CT = compile-time channel count (12 / 24 / 48: the CelebA / CIFAR levels; loops fully unrolled, no predicates),
// 0 = run-time C.
template <bool PERM, int CT>
__global__ void __launch_bounds__(256, 4)
rows_mix_kernel(const float* __restrict__ x, float* __restrict__ z, const float* __restrict__ w,
                const int64_t* __restrict__ idx, const float* __restrict__ bias, const float* __restrict__ logs,
                float f, int P, int C_rt, int reverse, int iters) {
  pdl_trigger();
  pdl_wait();
  const int C = CT ? CT : C_rt;
  extern __shared__ __align__(16) float smem[];
  float* wt = smem;                               // [C][C]: wt[i*C + o] = W[o][i]   (mix only)
  float* sc = wt + (PERM ? 0 : C * C);            // [C] exp(+-f*logs)
  float* bs = sc + C;                             // [C] bias
  int* sidx = reinterpret_cast<int*>(bs + C);     // [C] (perm only)
  const int og = threadIdx.x, slot = threadIdx.y;
  const int nthr = blockDim.x * blockDim.y;
  const int tid = slot * blockDim.x + og;
  // the pixel rows of a pass (2*blockDim.y contiguous pixels = one contiguous block of x) are staged through a
  // double buffer with cp.async: the copy of pass it+1 runs under the FMAs of pass it.  (Straight global loads made
  // every pass of a CTA one exposed memory round trip: ~16 us per launch at EVERY level, whatever its size.)
  const int ppp = 2 * blockDim.y;                 // pixels per pass
  float* xbuf = reinterpret_cast<float*>(sidx + C);   // [2][ppp][C]   (16-byte aligned: C is a multiple of 4)
  const int64_t total4 = (int64_t)P * C / 4;
  auto stage_pass = [&](int it_) {
    const int64_t base4 = (int64_t)(blockIdx.x * iters + it_) * ppp * C / 4;
    float* dst = xbuf + (it_ & 1) * ppp * C;
    const int n4 = ppp * C / 4;
    for (int i = tid; i < n4; i += nthr)
      if (base4 + i < total4) cp_async16(dst + 4 * i, x + 4 * (base4 + i));
    cp_async_commit();
  };
  stage_pass(0);
  const bool has_an = bias != nullptr;
  for (int c = tid; c < C; c += nthr) {
    const float l = has_an ? logs[c] * f : 0.f;
    sc[c] = has_an ? expf(reverse ? -l : l) : 1.f;
    bs[c] = has_an ? bias[c] : 0.f;
    if (PERM) sidx[c] = (int)idx[c];
  }
  if (!PERM)
    for (int o = slot; o < C; o += blockDim.y)      // thread (og, slot) transposes 4 elements of row o
#pragma unroll
      for (int u = 0; u < 4; ++u) wt[(og * 4 + u) * C + o] = w[o * C + og * 4 + u];
  __syncthreads();
  // two pixels per thread and pass (pixA, pixB = pixA + blockDim.y): every W^T float4 read from shared memory
  // feeds 8 FMAs instead of 4 -- the kernel was bound by the shared-memory pipe (M*C^2/16 LDS.128 per launch, the
  // same count at every level), not by HBM.  Each pixel's FMA order is unchanged.
  for (int it = 0; it < iters; ++it) {
    if ((int64_t)(blockIdx.x * iters + it) * ppp >= P) break;       // (CTA-uniform)
    if (it + 1 < iters) { stage_pass(it + 1); cp_async_wait_group<1>(); } else cp_async_wait_group<0>();
    __syncthreads();
    const int pixA = (blockIdx.x * iters + it) * ppp + slot;
    const int pixB = pixA + blockDim.y;
    const bool hasA = pixA < P, hasB = pixB < P;
    const float* xb_ = xbuf + (it & 1) * ppp * C;
    const float* xrA = xb_ + (hasA ? slot : 0) * C;
    const float* xrB = xb_ + (hasB ? slot + (int)blockDim.y : (hasA ? slot : 0)) * C;
    float accA[4], accB[4];
    if (PERM) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int o = og * 4 + a, sx = sidx[o];
        float tA = xrA[sx], tB = xrB[sx];
        if (has_an) {
          tA = reverse ? (tA * sc[o] - bs[o]) : ((tA + bs[sx]) * sc[sx]);
          tB = reverse ? (tB * sc[o] - bs[o]) : ((tB + bs[sx]) * sc[sx]);
        }
        accA[a] = tA; accB[a] = tB;
      }
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a) { accA[a] = 0.f; accB[a] = 0.f; }
#pragma unroll
      for (int i0 = 0; i0 < C; i0 += 12) {
        float4 xa4[3], xb4[3];                   // 12 channels of both pixels in flight: one memory latency per batch
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const bool in = i0 + 4 * b < C;
          xa4[b] = in ? *reinterpret_cast<const float4*>(xrA + i0 + 4 * b) : make_float4(0.f, 0.f, 0.f, 0.f);
          xb4[b] = in ? *reinterpret_cast<const float4*>(xrB + i0 + 4 * b) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const int i = i0 + 4 * b;
          if (i < C) {
            float xa[4] = {xa4[b].x, xa4[b].y, xa4[b].z, xa4[b].w};
            float xb[4] = {xb4[b].x, xb4[b].y, xb4[b].z, xb4[b].w};
            if (!reverse && has_an) {
              const float4 b4 = *reinterpret_cast<const float4*>(bs + i);
              const float4 s4 = *reinterpret_cast<const float4*>(sc + i);
              xa[0] = (xa[0] + b4.x) * s4.x; xa[1] = (xa[1] + b4.y) * s4.y;
              xa[2] = (xa[2] + b4.z) * s4.z; xa[3] = (xa[3] + b4.w) * s4.w;
              xb[0] = (xb[0] + b4.x) * s4.x; xb[1] = (xb[1] + b4.y) * s4.y;
              xb[2] = (xb[2] + b4.z) * s4.z; xb[3] = (xb[3] + b4.w) * s4.w;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 wv = *reinterpret_cast<const float4*>(wt + (i + u) * C + og * 4);
              accA[0] = fmaf(wv.x, xa[u], accA[0]); accA[1] = fmaf(wv.y, xa[u], accA[1]);
              accA[2] = fmaf(wv.z, xa[u], accA[2]); accA[3] = fmaf(wv.w, xa[u], accA[3]);
              accB[0] = fmaf(wv.x, xb[u], accB[0]); accB[1] = fmaf(wv.y, xb[u], accB[1]);
              accB[2] = fmaf(wv.z, xb[u], accB[2]); accB[3] = fmaf(wv.w, xb[u], accB[3]);
            }
          }
        }
      }
      if (reverse && has_an) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          accA[a] = accA[a] * sc[og * 4 + a] - bs[og * 4 + a];
          accB[a] = accB[a] * sc[og * 4 + a] - bs[og * 4 + a];
        }
      }
    }
    if (hasA) *reinterpret_cast<float4*>(z + (int64_t)pixA * C + og * 4) = make_float4(accA[0], accA[1], accA[2], accA[3]);
    if (hasB) *reinterpret_cast<float4*>(z + (int64_t)pixB * C + og * 4) = make_float4(accB[0], accB[1], accB[2], accB[3]);
    __syncthreads();                                // this pass's buffer is refilled by the pass after next
  }
}

// ------------------------------------------------------------------------------------------
// Reverse pass of a whole FlowStep behind the coupling network (model.py:131-152) in ONE kernel:
//   coupling^-1 (tap gather-sum of P3, Conv2dZeros scale, z2 = z2/scale - shift | z2 - h)  ->  W^-1 mix | perm^-1
//   ->  ActNorm^-1.
// rows_mix_kernel stages the pixel rows of a pass in shared memory anyway; here the staged rows are transformed in
// place by the inverse coupling before the mix reads them, so the sampling pass launches one flow kernel per step
// instead of two (the two were ~16 us each at every level: launch / latency bound) and the coupled rows make no
// HBM round trip.  Arithmetic (operation order included) is that of rows_coupling_kernel followed by
// rows_mix_kernel: results are bit-identical to the two-launch path.
// ------------------------------------------------------------------------------------------
template <bool PERM, int CT>
__global__ void __launch_bounds__(256, 4)
rows_coupling_rev_mix_kernel(const float* __restrict__ P3, int ldp, const float* __restrict__ bias3,
                             const float* __restrict__ logs3, float f3, int affine, int H, int W, FastDiv divW,
                             FastDiv divHW, FastDiv divCh, const float* __restrict__ x, float* __restrict__ z,
                             const float* __restrict__ w, const int64_t* __restrict__ idx,
                             const float* __restrict__ bias, const float* __restrict__ logs, float f, int P, int C_rt,
                             int iters) {
  pdl_trigger();
  pdl_wait();
  const int C = CT ? CT : C_rt;
  const int Ch = C >> 1, Cout = affine ? C : Ch, HW = H * W;
  extern __shared__ __align__(16) float smem[];
  float* wt = smem;                               // [C][C]: wt[i*C + o] = W^-1[o][i]   (mix only)
  float* sc = wt + (PERM ? 0 : C * C);            // [C] exp(-f*logs)
  float* bs = sc + C;                             // [C] bias
  float* s_b3 = bs + C;                           // [C] Conv2dZeros bias
  float* s_e3 = s_b3 + C;                         // [C] exp(f3*logs3)
  int* sidx = reinterpret_cast<int*>(s_e3 + C);   // [C] (perm only)
  const int og = threadIdx.x, slot = threadIdx.y;
  const int nthr = blockDim.x * blockDim.y;
  const int tid = slot * blockDim.x + og;
  const int ppp = 2 * blockDim.y;                 // pixels per pass
  float* xbuf = reinterpret_cast<float*>(sidx + C);   // [2][ppp][C]
  const int64_t total4 = (int64_t)P * C / 4;
  auto stage_pass = [&](int it_) {
    const int64_t base4 = (int64_t)(blockIdx.x * iters + it_) * ppp * C / 4;
    float* dst = xbuf + (it_ & 1) * ppp * C;
    const int n4 = ppp * C / 4;
    for (int i = tid; i < n4; i += nthr)
      if (base4 + i < total4) cp_async16(dst + 4 * i, x + 4 * (base4 + i));
    cp_async_commit();
  };
  stage_pass(0);
  const bool has_an = bias != nullptr;
  for (int c = tid; c < C; c += nthr) {
    const float l = has_an ? logs[c] * f : 0.f;
    sc[c] = has_an ? expf(-l) : 1.f;
    bs[c] = has_an ? bias[c] : 0.f;
    if (c < Cout) { s_b3[c] = bias3[c]; s_e3[c] = expf(logs3[c] * f3); }
    if (PERM) sidx[c] = (int)idx[c];
  }
  if (!PERM)
    for (int o = slot; o < C; o += blockDim.y)
#pragma unroll
      for (int u = 0; u < 4; ++u) wt[(og * 4 + u) * C + o] = w[o * C + og * 4 + u];
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int pix0 = (blockIdx.x * iters + it) * ppp;
    if (pix0 >= P) break;                                           // (CTA-uniform)
    if (it + 1 < iters) { stage_pass(it + 1); cp_async_wait_group<1>(); } else cp_async_wait_group<0>();
    __syncthreads();
    float* xb_ = xbuf + (it & 1) * ppp * C;
    // ---- inverse coupling on the staged rows: item = (pixel of the pass, channel pair | channel)
    for (int i = tid; i < ppp * Ch; i += nthr) {
      const int pl = fdiv(i, divCh), j = i - pl * Ch;
      const int pixg = pix0 + pl;
      if (pixg >= P) break;
      const int n = fdiv(pixg, divHW), pix = pixg - n * HW;
      const int yy = fdiv(pix, divW), xx = pix - yy * W;
      const bool up = yy > 0, dn = yy < H - 1, lf = xx > 0, rt = xx < W - 1;
      float* zp = xb_ + pl * C + Ch + j;
      if (affine) {
        const float* Pp = P3 + (int64_t)pixg * ldp + 2 * j;
        float u0 = 0.f, u1 = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dx < 0 ? lf : (dx > 0 ? rt : true));
          if (ok) {
            const float2 v = *reinterpret_cast<const float2*>(Pp + (dy * W + dx) * ldp + t * Cout);
            u0 += v.x; u1 += v.y;
          }
        }
        // __fmul_rn: no FMA contraction with the consumers below -- rows_coupling_kernel materialises shift / hsc
        // (it can store them for the backward pass), and the two paths must agree bit for bit
        const float shift = __fmul_rn(u0 + s_b3[2 * j], s_e3[2 * j]);
        const float hsc = __fmul_rn(u1 + s_b3[2 * j + 1], s_e3[2 * j + 1]);
        const float scale = 1.f / (1.f + expf(-(hsc + 2.f)));   // F.sigmoid(scale + 2.)
        float v = *zp;
        v = v / scale - shift;
        *zp = v;
      } else {
        const float* Pp = P3 + (int64_t)pixg * ldp + j;
        float u = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dx < 0 ? lf : (dx > 0 ? rt : true));
          if (ok) u += Pp[(dy * W + dx) * ldp + t * Cout];
        }
        const float h = __fmul_rn(u + s_b3[j], s_e3[j]);
        *zp = *zp - h;
      }
    }
    __syncthreads();
    // ---- W^-1 mix / inverse permutation + ActNorm^-1 (rows_mix_kernel, reverse)
    const int pixA = pix0 + slot;
    const int pixB = pixA + blockDim.y;
    const bool hasA = pixA < P, hasB = pixB < P;
    const float* xrA = xb_ + (hasA ? slot : 0) * C;
    const float* xrB = xb_ + (hasB ? slot + (int)blockDim.y : (hasA ? slot : 0)) * C;
    float accA[4], accB[4];
    if (PERM) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int o = og * 4 + a, sx = sidx[o];
        float tA = xrA[sx], tB = xrB[sx];
        if (has_an) { tA = tA * sc[o] - bs[o]; tB = tB * sc[o] - bs[o]; }
        accA[a] = tA; accB[a] = tB;
      }
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a) { accA[a] = 0.f; accB[a] = 0.f; }
#pragma unroll
      for (int i0 = 0; i0 < C; i0 += 12) {
        float4 xa4[3], xb4[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const bool in = i0 + 4 * b < C;
          xa4[b] = in ? *reinterpret_cast<const float4*>(xrA + i0 + 4 * b) : make_float4(0.f, 0.f, 0.f, 0.f);
          xb4[b] = in ? *reinterpret_cast<const float4*>(xrB + i0 + 4 * b) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const int i = i0 + 4 * b;
          if (i < C) {
            const float xa[4] = {xa4[b].x, xa4[b].y, xa4[b].z, xa4[b].w};
            const float xb[4] = {xb4[b].x, xb4[b].y, xb4[b].z, xb4[b].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 wv = *reinterpret_cast<const float4*>(wt + (i + u) * C + og * 4);
              accA[0] = fmaf(wv.x, xa[u], accA[0]); accA[1] = fmaf(wv.y, xa[u], accA[1]);
              accA[2] = fmaf(wv.z, xa[u], accA[2]); accA[3] = fmaf(wv.w, xa[u], accA[3]);
              accB[0] = fmaf(wv.x, xb[u], accB[0]); accB[1] = fmaf(wv.y, xb[u], accB[1]);
              accB[2] = fmaf(wv.z, xb[u], accB[2]); accB[3] = fmaf(wv.w, xb[u], accB[3]);
            }
          }
        }
      }
      if (has_an) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          accA[a] = accA[a] * sc[og * 4 + a] - bs[og * 4 + a];
          accB[a] = accB[a] * sc[og * 4 + a] - bs[og * 4 + a];
        }
      }
    }
    if (hasA) *reinterpret_cast<float4*>(z + (int64_t)pixA * C + og * 4) = make_float4(accA[0], accA[1], accA[2], accA[3]);
    if (hasB) *reinterpret_cast<float4*>(z + (int64_t)pixB * C + og * 4) = make_float4(accB[0], accB[1], accB[2], accB[3]);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Tap gather-sum (second half of Conv2dZeros as nine pointwise GEMMs, module.py:295-296) + coupling
// (model.py:105-115 fwd, 131-140 rev) + this step's logdet (module.py:77-82, 357-367; model.py:114,140).
// grid = (nblk, N), block = (Ch, PPB): thread (j, slot) owns channel pair j of one pixel; the Ch lanes of a
// pixel read 8*Ch (affine) contiguous bytes per tap.  Tap addresses are base + uniform offsets; the four
// border predicates are computed once.  The last CTA of a sample (atomic ticket, self-resetting) sums the
// per-CTA partials with a fixed-shape tree:
//   ld_out[n] = ld_in[n] + sign*HW*(sum_c f*an_logs[c] + logabsdet[0]) + sum_b partial[n][b]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8)
rows_coupling_kernel(const float* __restrict__ P3, int ldp, const float* __restrict__ bias3,
                     const float* __restrict__ logs3, float f, float* __restrict__ z,
                     float* __restrict__ h_save, int C, int H, int W, FastDiv divW, FastDiv divCh, int ppb, int iters,
                     int affine, int reverse,
                     const float* __restrict__ ld_in, float* __restrict__ ld_out,
                     const float* __restrict__ an_logs, float an_f, const float* __restrict__ logabsdet,
                     float sign, float* __restrict__ partials, unsigned int* __restrict__ tickets) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  __shared__ float s_b[ROWS_MAX_C_WIDE], s_e[ROWS_MAX_C_WIDE];
  __shared__ int s_last;
  const int HW = H * W, Ch = C >> 1, Cout = affine ? C : Ch;
  const int n = blockIdx.y;
  // 256 threads = PPB pixels x Ch channel pairs (+ a few idle lanes, so that every warp is complete for the
  // shuffle reductions below)
  const int tid = threadIdx.x, nthr = 256;
  const int slot = fdiv(tid, divCh), j = tid - slot * Ch;
  for (int c = tid; c < Cout; c += nthr) { s_b[c] = bias3[c]; s_e[c] = expf(logs3[c] * f); }
  __syncthreads();
  float lsum = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int pix = (blockIdx.x * iters + it) * ppb + slot;
    if (slot >= ppb || pix >= HW) break;
    const int yy = fdiv(pix, divW), xx = pix - yy * W;
    const bool up = yy > 0, dn = yy < H - 1, lf = xx > 0, rt = xx < W - 1;
    const int64_t row = (int64_t)n * HW + pix;
    float* zp = z + row * C + Ch + j;
    if (affine) {
      const float* Pp = P3 + row * ldp + 2 * j;
      float u0 = 0.f, u1 = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dx < 0 ? lf : (dx > 0 ? rt : true));
        if (ok) {
          const float2 v = *reinterpret_cast<const float2*>(Pp + (dy * W + dx) * ldp + t * Cout);
          u0 += v.x; u1 += v.y;
        }
      }
      const float shift = (u0 + s_b[2 * j]) * s_e[2 * j];
      const float hsc = (u1 + s_b[2 * j + 1]) * s_e[2 * j + 1];
      const float scale = 1.f / (1.f + expf(-(hsc + 2.f)));   // F.sigmoid(scale + 2.)
      float v = *zp;
      if (!reverse) { v = (v + shift) * scale; lsum += logf(scale); }
      else { v = v / scale - shift; lsum -= logf(scale); }
      *zp = v;
      if (h_save) *reinterpret_cast<float2*>(h_save + row * Cout + 2 * j) = make_float2(shift, hsc);
    } else {
      const float* Pp = P3 + row * ldp + j;
      float u = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dx < 0 ? lf : (dx > 0 ? rt : true));
        if (ok) u += Pp[(dy * W + dx) * ldp + t * Cout];
      }
      const float h = (u + s_b[j]) * s_e[j];
      float v = *zp;
      v = reverse ? v - h : v + h;
      *zp = v;
      if (h_save) h_save[row * Cout + j] = h;
    }
  }
  if (!ld_out) return;
  const float tot = block_sum(lsum, red);
  const int nblk = gridDim.x;
  if (tid == 0) {
    partials[n * nblk + blockIdx.x] = tot;
    __threadfence();
    const unsigned int t = atomicAdd(tickets + n, 1u);
    s_last = (t == (unsigned int)(nblk - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float a = 0.f;
  if (an_logs)
    for (int c = tid; c < C; c += nthr) a += an_logs[c] * an_f;
  const float term = block_sum(a, red);
  float ps = 0.f;
  if (affine)
    for (int b = tid; b < nblk; b += nthr) ps += __ldcg(partials + n * nblk + b);
  const float psum = block_sum(ps, red);
  if (tid == 0) {
    float v = ld_in ? ld_in[n] : 0.f;
    v += sign * (term * (float)HW);                              // torch.sum(logs)*HW   (module.py:78-80)
    if (logabsdet) v += sign * (logabsdet[0] * (float)HW);       // log|det W| * HW      (module.py:357)
    v += psum;
    ld_out[n] = v;
    tickets[n] = 0u;                                             // ready for the next launch on this stream
  }
}

// ------------------------------------------------------------------------------------------
// Same operation with the P3 rows of the CTA's pixel run staged in shared memory.  In rows_coupling_kernel
// every lane gathers its nine 8-byte taps straight from global memory, from nine rows ldp*4 bytes apart: the
// kernel is bound by load latency and L1 wavefronts (ncu: 48 % of the stall samples on the first tap add, IPC
// 1.1), 59 % of the HBM roofline at level 1 and 20-30 % at levels 2-3.  Here a CTA owns a contiguous run of
// T pixels of one sample, copies rows [p0-W-1, p1+W+1) of P3 (a contiguous block of global memory) with 16-byte
// cp.async in one shot -- fully coalesced, every byte of P3 fetched by ~1.2 CTAs instead of through nine
// scattered sector reads -- and gathers the taps from shared memory (pitch ldp+4 floats: conflict-light).
// Arithmetic, summation order and the per-sample logdet reduction are those of rows_coupling_kernel.
// ------------------------------------------------------------------------------------------
// mbarrier + bulk (TMA engine, no tensor map) global -> shared copies: one instruction per contiguous run
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

__global__ void __launch_bounds__(256)
rows_coupling_win_kernel(const float* __restrict__ P3, int ldp, const float* __restrict__ bias3,
                         const float* __restrict__ logs3, float f, float* __restrict__ z,
                         float* __restrict__ h_save, int C, int H, int W, FastDiv divW, FastDiv divCh, FastDiv divVec,
                         int T, int pitch, int affine, int reverse,
                         const float* __restrict__ ld_in, float* __restrict__ ld_out,
                         const float* __restrict__ an_logs, float an_f, const float* __restrict__ logabsdet,
                         float sign, float* __restrict__ partials, unsigned int* __restrict__ tickets) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float win[];
  __shared__ float red[32];
  __shared__ float s_b[ROWS_MAX_C_WIDE], s_e[ROWS_MAX_C_WIDE];
  __shared__ int s_last;
  const int HW = H * W, Ch = C >> 1, Cout = affine ? C : Ch;
  const int n = blockIdx.y;
  const int tid = threadIdx.x, nthr = 256;
  const int p0 = blockIdx.x * T, p1 = min(p0 + T, HW);
  const int r0 = max(p0 - W - 1, 0), r1 = min(p1 + W + 1, HW);
  {
    const int vec = (int)divVec.d;                     // float4 per row that hold the 9*Cout tap columns
    const float* src = P3 + ((int64_t)n * HW + r0) * ldp;
    const int total = (r1 - r0) * vec;
    for (int i = tid; i < total; i += nthr) {
      const int r = fdiv(i, divVec), v = i - r * vec;
      cp_async16(win + r * pitch + 4 * v, src + (int64_t)r * ldp + 4 * v);
    }
  }
  for (int c = tid; c < Cout; c += nthr) { s_b[c] = bias3[c]; s_e[c] = expf(logs3[c] * f); }
  cp_async_wait_all();
  __syncthreads();
  float lsum = 0.f;
  const int items = (p1 - p0) * Ch;
  for (int i = tid; i < items; i += nthr) {
    const int pl = fdiv(i, divCh), j = i - pl * Ch;
    const int pix = p0 + pl;
    const int yy = fdiv(pix, divW), xx = pix - yy * W;
    const bool up = yy > 0, dn = yy < H - 1, lf = xx > 0, rt = xx < W - 1;
    const int64_t row = (int64_t)n * HW + pix;
    float* zp = z + row * C + Ch + j;
    float v = *zp;
    const float* wrow = win + (pix - r0) * pitch;
    if (affine) {
      const float* Pp = wrow + 2 * j;
      float u0 = 0.f, u1 = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dx < 0 ? lf : (dx > 0 ? rt : true));
        if (ok) {
          const float2 tv = *reinterpret_cast<const float2*>(Pp + (dy * W + dx) * pitch + t * Cout);
          u0 += tv.x; u1 += tv.y;
        }
      }
      const float shift = (u0 + s_b[2 * j]) * s_e[2 * j];
      const float hsc = (u1 + s_b[2 * j + 1]) * s_e[2 * j + 1];
      const float scale = 1.f / (1.f + expf(-(hsc + 2.f)));   // F.sigmoid(scale + 2.)
      if (!reverse) { v = (v + shift) * scale; lsum += logf(scale); }
      else { v = v / scale - shift; lsum -= logf(scale); }
      *zp = v;
      if (h_save) *reinterpret_cast<float2*>(h_save + row * Cout + 2 * j) = make_float2(shift, hsc);
    } else {
      const float* Pp = wrow + j;
      float u = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dx < 0 ? lf : (dx > 0 ? rt : true));
        if (ok) u += Pp[(dy * W + dx) * pitch + t * Cout];
      }
      const float h = (u + s_b[j]) * s_e[j];
      v = reverse ? v - h : v + h;
      *zp = v;
      if (h_save) h_save[row * Cout + j] = h;
    }
  }
  if (!ld_out) return;
  const float tot = block_sum(lsum, red);
  const int nblk = gridDim.x;
  if (tid == 0) {
    partials[n * nblk + blockIdx.x] = tot;
    __threadfence();
    const unsigned int t = atomicAdd(tickets + n, 1u);
    s_last = (t == (unsigned int)(nblk - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float a = 0.f;
  if (an_logs)
    for (int c = tid; c < C; c += nthr) a += an_logs[c] * an_f;
  const float term = block_sum(a, red);
  float ps = 0.f;
  if (affine)
    for (int b = tid; b < nblk; b += nthr) ps += __ldcg(partials + n * nblk + b);
  const float psum = block_sum(ps, red);
  if (tid == 0) {
    float v = ld_in ? ld_in[n] : 0.f;
    v += sign * (term * (float)HW);
    if (logabsdet) v += sign * (logabsdet[0] * (float)HW);
    v += psum;
    ld_out[n] = v;
    tickets[n] = 0u;
  }
}

// ------------------------------------------------------------------------------------------
// Coupling backward (see coupling_bwd_kernel in flow_bwd_kernels.cu for the formulas).
// block = (Ch, PPB): thread (j, slot) keeps j over `iters` pixel groups, so the per-channel sums for
// dlogs3 / dbias3 are accumulated in registers and reduced once per CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_coupling_bwd_kernel(const float* __restrict__ y, const float* __restrict__ hrows,
                         const float* __restrict__ dy, const float* __restrict__ dld,
                         const float* __restrict__ logs3, float f, float* __restrict__ dz,
                         float* __restrict__ du, float* __restrict__ dlogs3, float* __restrict__ dbias3,
                         int NP, int C, FastDiv divHW, int affine, int iters) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_part[4][256];
  __shared__ float s_e[ROWS_MAX_C_WIDE];
  const int Ch = C >> 1, Cout = affine ? C : Ch;
  const int j = threadIdx.x, slot = threadIdx.y, ppb = blockDim.y;
  const int tid = slot * Ch + j;
  for (int c = tid; c < Cout; c += Ch * ppb) s_e[c] = expf(logs3[c] * f);
  __syncthreads();
  float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int pix = (blockIdx.x * iters + it) * ppb + slot;
    if (pix >= NP) break;
    const float g = dld ? dld[fdiv(pix, divHW)] : 0.f;
    const int64_t o1 = (int64_t)pix * C + j, o2 = o1 + Ch;
    dz[o1] = dy[o1];
    const float dy2 = dy[o2];
    if (affine) {
      const float2 hv = *reinterpret_cast<const float2*>(hrows + (int64_t)pix * Cout + 2 * j);
      const float h0 = hv.x, h1 = hv.y;
      const float scale = 1.f / (1.f + expf(-(h1 + 2.f)));
      const float zs = y[o2] / scale;            // z2 + shift
      dz[o2] = dy2 * scale;
      const float dh0 = dy2 * scale;
      const float dh1 = (dy2 * zs + g / scale) * scale * (1.f - scale);
      *reinterpret_cast<float2*>(du + (int64_t)pix * Cout + 2 * j) = make_float2(dh0 * s_e[2 * j], dh1 * s_e[2 * j + 1]);
      a0 += dh0 * h0; b0 += dh0; a1 += dh1 * h1; b1 += dh1;
    } else {
      const float h0 = hrows[(int64_t)pix * Cout + j];
      dz[o2] = dy2;
      du[(int64_t)pix * Cout + j] = dy2 * s_e[j];
      a0 += dy2 * h0; b0 += dy2;
    }
  }
  s_part[0][tid] = a0; s_part[1][tid] = b0; s_part[2][tid] = a1; s_part[3][tid] = b1;
  __syncthreads();
  // thread (q, j): sum of quantity q over the pixel slots, then one global atomic per channel and CTA
  const int nq = affine ? 4 : 2;
  for (int t = tid; t < nq * Ch; t += Ch * ppb) {          // (more than one pass only for C > 128)
    const int q = t / Ch, jj = t - q * Ch;
    float s = 0.f;
    for (int k = 0; k < ppb; ++k) s += s_part[q][k * Ch + jj];
    const int c = affine ? (2 * jj + (q >> 1)) : jj;
    if ((q & 1) == 0) atomicAdd(dlogs3 + c, f * s);
    else atomicAdd(dbias3 + c, s_e[c] * s);
  }
}

// ------------------------------------------------------------------------------------------
// ActNorm + mix backward (model.py:94-103) with the conv1 dgrad tap gather-sum fused into the load and the
// gradient of the sample-independent logdet terms (module.py:78-82, 357-363) folded into CTA 0:
//   dz[p][c] += sum_tap dA1[nbr(p,tap)][tap*Cin + c]   (c < Cin; transposed conv => mirrored taps)
//   a = (x+b)*s ; da = W^T dz ; dx = da*s ; db += sum da*s ; dlogs += f*sum da*a ; dW += sum_p dz a^T
//   G = HW*sum_n dld[n] ; dlogs += f*G ; dW += G*W^-T
// A CTA walks tiles of TP (32..128) pixels staged in shared memory as [TP][C] (16-byte rows: every global and
// most shared accesses are float4).  dW is accumulated in registers over ALL the CTA's tiles as 4x4 blocks
// (x a split of the tile's pixels, so that all 256 threads work) and leaves the CTA once, like dbias / dlogs.
// ------------------------------------------------------------------------------------------
// RMB_MAXIT = 4x4 dW blocks per thread: 1 while (C/4)^2 <= 256 (C <= 64), 3 up to C = 96 (576 blocks)
template <bool PERM, int RMB_MAXIT>
__global__ void __launch_bounds__(256)
rows_mix_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dz, const float* __restrict__ dA1,
                    int ld_a1, int Cin, const float* __restrict__ w, const int64_t* __restrict__ idx,
                    const float* __restrict__ bias, const float* __restrict__ logs, float f,
                    float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dlogs,
                    float* __restrict__ dbias, int NP, int C, int H, int W, FastDiv divW, FastDiv divHW,
                    FastDiv divG, int TP, const float* __restrict__ dld, int Nld, const float* __restrict__ winv) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                           // [C][C]    W, later this CTA's dW (mix only)
  float* a_s = ws + (PERM ? 0 : C * C);       // [TP][C]   a = actnorm(x)
  float* d_s = a_s + TP * C;                  // [TP][C]   dz, later da
  float* sc = d_s + TP * C;                   // [C]
  float* bs = sc + C;                         // [C]
  float* s_red = bs + C;                      // [2][C]
  int* sinv = reinterpret_cast<int*>(s_red + 2 * C);   // [C] inverse permutation (perm only)
  float* red = reinterpret_cast<float*>(sinv + C);      // [32] + [1]
  const int tid = threadIdx.x;
  const bool has_an = bias != nullptr;
  const int HW = H * W, G = C >> 2;
  for (int c = tid; c < C; c += 256) {
    sc[c] = has_an ? expf(logs[c] * f) : 1.f;
    bs[c] = has_an ? bias[c] : 0.f;
    s_red[c] = 0.f; s_red[C + c] = 0.f;
    if (PERM) sinv[(int)idx[c]] = c;          // z[o] = a[idx[o]]  =>  da[i] = dz[o] with idx[o] == i
  }
  if (!PERM)
    for (int e = tid; e < C * C; e += 256) ws[e] = w[e];
  __syncthreads();
  // (pixel slot, channel quad) of this thread in the staging / copy-out passes
  const int slot = fdiv(tid, divG), quad = tid - slot * G;
  const int ppb = 256 / G;
  // 4x4 dW blocks of this thread: item = (block, pixel split); persistent register accumulators
  const int nb4 = G * G;
  int psplit = 256 / nb4;
  if (psplit < 1) psplit = 1;
  const int nitems = nb4 * psplit;
  float dwacc[RMB_MAXIT][16];
#pragma unroll
  for (int k = 0; k < RMB_MAXIT; ++k)
#pragma unroll
    for (int u = 0; u < 16; ++u) dwacc[k][u] = 0.f;
  // channel reductions: thread (channel i, pixel chunk ch) keeps its partial sums over all tiles
  const int chunks = 256 / C;
  const int red_ch = tid / C, red_i = tid - red_ch * C;
  float sg_acc = 0.f, sga_acc = 0.f;

  const int ntiles = (NP + TP - 1) / TP;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int g0 = tile * TP;
    // ---- stage a and dz (+ conv1 dgrad), one float4 of one pixel per thread and pass
    if (slot < ppb) {
      for (int p = slot; p < TP; p += ppb) {
        const int pix = g0 + p;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
        if (pix < NP) {
          xv = *reinterpret_cast<const float4*>(x + (int64_t)pix * C + quad * 4);
          dv = *reinterpret_cast<const float4*>(dz + (int64_t)pix * C + quad * 4);
          const int c0 = quad * 4;
          if (dA1 && c0 < Cin) {
            const int n = fdiv(pix, divHW);
            const int q = pix - n * HW;
            const int yy = fdiv(q, divW), xx = q - yy * W;
            const bool up = yy > 0, dn = yy < H - 1, lf = xx > 0, rt = xx < W - 1;
            const bool hi = c0 + 2 < Cin;                       // channels c0+2, c0+3 also belong to z1
            const float* ap = dA1 + (int64_t)pix * ld_a1 + c0;
            float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const int tap = 8 - t;
              const int dy = tap / 3 - 1, dxx = tap % 3 - 1;
              const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dxx < 0 ? lf : (dxx > 0 ? rt : true));
              if (ok) {
                const float* tp = ap + (dy * W + dxx) * ld_a1 + t * Cin;
                const float2 v0 = *reinterpret_cast<const float2*>(tp);
                r0 += v0.x; r1 += v0.y;
                if (hi) {
                  const float2 v1 = *reinterpret_cast<const float2*>(tp + 2);
                  r2 += v1.x; r3 += v1.y;
                }
              }
            }
            dv.x = dv.x + r0; dv.y = dv.y + r1;
            if (hi) { dv.z = dv.z + r2; dv.w = dv.w + r3; }
          }
          const float4 b4 = *reinterpret_cast<const float4*>(bs + quad * 4);
          const float4 s4 = *reinterpret_cast<const float4*>(sc + quad * 4);
          xv.x = (xv.x + b4.x) * s4.x; xv.y = (xv.y + b4.y) * s4.y;
          xv.z = (xv.z + b4.z) * s4.z; xv.w = (xv.w + b4.w) * s4.w;
        }
        *reinterpret_cast<float4*>(a_s + p * C + quad * 4) = xv;
        *reinterpret_cast<float4*>(d_s + p * C + quad * 4) = dv;
      }
    }
    __syncthreads();
    // ---- dW[o][i] += sum_p dz[p][o] a[p][i] on 4x4 register blocks
    if (!PERM) {
#pragma unroll
      for (int k = 0; k < RMB_MAXIT; ++k) {
        const int it = tid + k * 256;
        if (it < nitems) {
          const int ps = it / nb4, blk = it - ps * nb4;
          const int bo = fdiv(blk, divG), bi = blk - bo * G;
          for (int p = ps; p < TP; p += psplit) {
            const float4 d4 = *reinterpret_cast<const float4*>(d_s + p * C + bo * 4);
            const float4 a4 = *reinterpret_cast<const float4*>(a_s + p * C + bi * 4);
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v = 0; v < 4; ++v) dwacc[k][u * 4 + v] = fmaf(dd[u], aa[v], dwacc[k][u * 4 + v]);
          }
        }
      }
    }
    __syncthreads();                           // dW has read every dz element: da may now overwrite dz in place
    // ---- da[p][i] = sum_o W[o][i] dz[p][o].  The G threads of a pixel sit in ONE warp (32/G pixels per warp
    // pass), so a pixel's dz row is replaced by its da row between two __syncwarp()s.
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int ppw = 32 / G;
      const int pl = fdiv(lane, divG), ig = lane - pl * G;
      for (int p0 = warp * ppw; p0 < TP; p0 += 8 * ppw) {
        const int p = p0 + pl;
        const bool act = pl < ppw && p < TP;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (act) {
          if (PERM) {
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = d_s[p * C + sinv[ig * 4 + u]];
          } else {
            for (int o = 0; o < C; o += 4) {
              const float4 d4 = *reinterpret_cast<const float4*>(d_s + p * C + o);
              const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 wv = *reinterpret_cast<const float4*>(ws + (o + u) * C + ig * 4);
                acc[0] = fmaf(wv.x, dd[u], acc[0]); acc[1] = fmaf(wv.y, dd[u], acc[1]);
                acc[2] = fmaf(wv.z, dd[u], acc[2]); acc[3] = fmaf(wv.w, dd[u], acc[3]);
              }
            }
          }
        }
        __syncwarp();
        if (act) *reinterpret_cast<float4*>(d_s + p * C + ig * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- per-channel reductions over the tile's pixels (registers; flushed once per CTA)
    if (has_an && red_ch < chunks) {
      for (int p = red_ch; p < TP; p += chunks) {
        const float d = d_s[p * C + red_i];
        sg_acc += d; sga_acc = fmaf(d, a_s[p * C + red_i], sga_acc);
      }
    }
    // ---- dx = da * s, written back as whole pixels
    if (slot < ppb) {
      const float4 s4 = *reinterpret_cast<const float4*>(sc + quad * 4);
      for (int p = slot; p < TP; p += ppb) {
        const int pix = g0 + p;
        if (pix < NP) {
          const float4 d4 = *reinterpret_cast<const float4*>(d_s + p * C + quad * 4);
          *reinterpret_cast<float4*>(dx + (int64_t)pix * C + quad * 4) =
              make_float4(d4.x * s4.x, d4.y * s4.y, d4.z * s4.z, d4.w * s4.w);
        }
      }
    }
    __syncthreads();                           // the tile buffers are free for the next tile
  }
  // ---- one round of global atomics per CTA: partials are first combined in shared memory (W is dead now)
  if (has_an && red_ch < chunks) {
    atomicAdd(&s_red[red_i], sg_acc);
    atomicAdd(&s_red[C + red_i], sga_acc);
  }
  if (!PERM) {
    for (int e = tid; e < C * C; e += 256) ws[e] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RMB_MAXIT; ++k) {
      const int it = tid + k * 256;
      if (it < nitems) {
        const int ps = it / nb4, blk = it - ps * nb4;
        const int bo = fdiv(blk, divG), bi = blk - bo * G;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) atomicAdd(&ws[(bo * 4 + u) * C + bi * 4 + v], dwacc[k][u * 4 + v]);
      }
    }
    __syncthreads();
    for (int e = tid; e < C * C; e += 256) atomicAdd(dw + e, ws[e]);
  }
  __syncthreads();
  if (has_an)
    for (int c = tid; c < C; c += 256) {
      atomicAdd(dbias + c, s_red[c] * sc[c]);
      atomicAdd(dlogs + c, f * s_red[C + c]);
    }
  // ---- sample-independent logdet terms (one CTA)
  if (blockIdx.x == 0 && dld) {
    float a = 0.f;
    for (int n = tid; n < Nld; n += 256) a += dld[n];
    const float tot = block_sum(a, red);
    if (tid == 0) red[32] = tot * (float)HW;
    __syncthreads();
    const float Gs = red[32];
    if (has_an)
      for (int c = tid; c < C; c += 256) atomicAdd(dlogs + c, f * Gs);
    if (!PERM && winv)
      for (int e = tid; e < C * C; e += 256) {
        const int o = e / C, i = e - o * C;
        atomicAdd(dw + e, Gs * winv[i * C + o]);
      }
  }
}

// ------------------------------------------------------------------------------------------
// Same adjoint with the conv1-dgrad operand staged through shared memory.  In rows_mix_bwd_kernel every
// (pixel, channel quad) thread gathers its 9..18 taps of dA1 straight from global memory, from nine rows ld_a1*4
// bytes apart (ncu: 38 % of the stall samples wait on those loads, IPC 1.0, 30 % of warp slots occupied;
// 20-25 % of the HBM roofline).  Here the rows [g0-W-1, g0+TP+W+1) of dA1 -- contiguous in global memory -- are
// copied with 16-byte cp.async while x and dz are staged, and the taps are gathered from shared memory by
// (pixel, channel pair) items, which also balances the work (the quad mapping left a third of the threads idle).
// The window is filled by the TMA engine: one bulk copy per row (16 x ceil(9*Cin/4) bytes into a padded pitch),
// completion counted on an mbarrier every thread arrives on (a per-16-byte cp.async loop cost ~25 % of the
// kernel's instructions).
// Further changes: compile-time channel count CT (12/24/48; 0 = run time) and a conflict-free shared-memory
// combine of the register dW blocks instead of contended shared float atomics (CAS loops).
// Element arithmetic and summation order of dx are those of rows_mix_bwd_kernel.
// ------------------------------------------------------------------------------------------
template <bool PERM, int RMB_MAXIT, int CT, typename TA>
__global__ void __launch_bounds__(256, RMB_MAXIT == 1 ? 4 : 1)
rows_mix_bwd_win_kernel(const float* __restrict__ x, const float* __restrict__ dz, const TA* __restrict__ dA1,
                        int ld_a1, int Cin, const float* __restrict__ w, const int64_t* __restrict__ idx,
                        const float* __restrict__ bias, const float* __restrict__ logs, float f,
                        float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dlogs,
                        float* __restrict__ dbias, int NP, int C_rt, int H, int W, FastDiv divW, FastDiv divHW,
                        FastDiv divG, FastDiv divVec, FastDiv divPair, int TP, int wpitch, int win_floats,
                        const float* __restrict__ dld, int Nld, const float* __restrict__ winv) {
  pdl_trigger();
  pdl_wait();
  const int C = CT ? CT : C_rt;
  const int G = C >> 2;
  extern __shared__ __align__(16) float smem[];
  float* win = smem;                          // [TP + 2W + 2][wpitch] dA1 rows (TA); later the dW combine scratch
  TA* wina = reinterpret_cast<TA*>(smem);     // wpitch is in TA elements (a multiple of 16 bytes)
  float* ws = win + win_floats;               // [C][C]    W (mix only)
  float* a_s = ws + (PERM ? 0 : C * C);       // [TP][C]   a = actnorm(x)
  float* d_s = a_s + TP * C;                  // [TP][C]   dz, later da
  float* sc = d_s + TP * C;                   // [C]
  float* bs = sc + C;                         // [C]
  float* s_red = bs + C;                      // [2][C]
  int* sinv = reinterpret_cast<int*>(s_red + 2 * C);   // [C] inverse permutation (perm only)
  float* red = reinterpret_cast<float*>(sinv + C);      // [32] + [1]
  __shared__ uint64_t s_bar;                  // window-copy barrier: 256 arrivals + the rows' bytes per tile
  const int tid = threadIdx.x;
  const bool has_an = bias != nullptr;
  const int HW = H * W;
  if (tid == 0) {
    mbar_init(&s_bar, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t win_parity = 0;
  auto div_g = [&](int v) { return CT ? v / (CT ? CT / 4 : 1) : fdiv(v, divG); };
  for (int c = tid; c < C; c += 256) {
    sc[c] = has_an ? expf(logs[c] * f) : 1.f;
    bs[c] = has_an ? bias[c] : 0.f;
    s_red[c] = 0.f; s_red[C + c] = 0.f;
    if (PERM) sinv[(int)idx[c]] = c;
  }
  if (!PERM)
    for (int e = tid; e < C * C; e += 256) ws[e] = w[e];
  __syncthreads();
  const int slot = div_g(tid), quad = tid - slot * G;
  const int ppb = 256 / G;
  const int nb4 = G * G;
  int psplit = 256 / nb4;
  if (psplit < 1) psplit = 1;
  const int nitems = nb4 * psplit;
  float dwacc[RMB_MAXIT][16];
#pragma unroll
  for (int k = 0; k < RMB_MAXIT; ++k)
#pragma unroll
    for (int u = 0; u < 16; ++u) dwacc[k][u] = 0.f;
  const int chunks = 256 / C;
  const int red_ch = tid / C, red_i = tid - red_ch * C;
  float sg_acc = 0.f, sga_acc = 0.f;
  const int vec = (int)divVec.d;              // 16-byte pieces per dA1 row that hold its 9*Cin columns
  const int npair = Cin >> 1;

  const int ntiles = (NP + TP - 1) / TP;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int g0 = tile * TP;
    const int wr0 = max(g0 - W - 1, 0), wr1 = min(g0 + TP + W + 1, NP);
    // ---- (a) async copy of the dA1 window: thread r copies row wr0 + r (and r + 256, ...), everyone arrives once
    {
      const int nrows = wr1 - wr0;
      uint32_t my_bytes = 0;
      for (int r = tid; r < nrows; r += 256) my_bytes += (uint32_t)vec * 16u;
      if (my_bytes) mbar_arrive_expect_tx(&s_bar, my_bytes); else mbar_arrive(&s_bar);
      for (int r = tid; r < nrows; r += 256)
        bulk_g2s(wina + r * wpitch, dA1 + (int64_t)(wr0 + r) * ld_a1, (uint32_t)vec * 16u, &s_bar);
    }
    // ---- (b) stage a = actnorm(x) and dz, one float4 of one pixel per thread and pass
    if (slot < ppb) {
      for (int p = slot; p < TP; p += ppb) {
        const int pix = g0 + p;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
        if (pix < NP) {
          xv = *reinterpret_cast<const float4*>(x + (int64_t)pix * C + quad * 4);
          dv = *reinterpret_cast<const float4*>(dz + (int64_t)pix * C + quad * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(bs + quad * 4);
          const float4 s4 = *reinterpret_cast<const float4*>(sc + quad * 4);
          xv.x = (xv.x + b4.x) * s4.x; xv.y = (xv.y + b4.y) * s4.y;
          xv.z = (xv.z + b4.z) * s4.z; xv.w = (xv.w + b4.w) * s4.w;
        }
        *reinterpret_cast<float4*>(a_s + p * C + quad * 4) = xv;
        *reinterpret_cast<float4*>(d_s + p * C + quad * 4) = dv;
      }
    }
    mbar_wait(&s_bar, win_parity);
    win_parity ^= 1u;
    __syncthreads();
    // ---- (c) conv1 dgrad: dz[p][c] += sum_tap dA1[nbr(p, tap)][tap*Cin + c], c < Cin, taps in the order of
    // rows_mix_bwd_kernel; one (pixel, channel pair) item per thread and pass
    {
      const int total = TP * npair;
      for (int i = tid; i < total; i += 256) {
        const int p = fdiv(i, divPair), j = i - p * npair;
        const int pix = g0 + p;
        if (pix < NP) {
          const int n = fdiv(pix, divHW);
          const int q = pix - n * HW;
          const int yy = fdiv(q, divW), xx = q - yy * W;
          const bool up = yy > 0, dn = yy < H - 1, lf = xx > 0, rt = xx < W - 1;
          const TA* ap = wina + (pix - wr0) * wpitch + 2 * j;
          float r0 = 0.f, r1 = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int tap = 8 - t;
            const int dy = tap / 3 - 1, dxx = tap % 3 - 1;
            const bool ok = (dy < 0 ? up : (dy > 0 ? dn : true)) && (dxx < 0 ? lf : (dxx > 0 ? rt : true));
            if (ok) {
              float2 v0;
              if (sizeof(TA) == 4) v0 = *reinterpret_cast<const float2*>(ap + (dy * W + dxx) * wpitch + t * Cin);
              else v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ap + (dy * W + dxx) * wpitch + t * Cin));
              r0 += v0.x; r1 += v0.y;
            }
          }
          float2* dp = reinterpret_cast<float2*>(d_s + p * C + 2 * j);
          float2 dv = *dp;
          dv.x = dv.x + r0; dv.y = dv.y + r1;
          *dp = dv;
        }
      }
    }
    __syncthreads();
    // ---- dW[o][i] += sum_p dz[p][o] a[p][i] on 4x4 register blocks
    if (!PERM) {
#pragma unroll
      for (int k = 0; k < RMB_MAXIT; ++k) {
        const int it = tid + k * 256;
        if (it < nitems) {
          const int ps = it / nb4, blk = it - ps * nb4;
          const int bo = div_g(blk), bi = blk - bo * G;
          for (int p = ps; p < TP; p += psplit) {
            const float4 d4 = *reinterpret_cast<const float4*>(d_s + p * C + bo * 4);
            const float4 a4 = *reinterpret_cast<const float4*>(a_s + p * C + bi * 4);
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v = 0; v < 4; ++v) dwacc[k][u * 4 + v] = fmaf(dd[u], aa[v], dwacc[k][u * 4 + v]);
          }
        }
      }
    }
    __syncthreads();                           // dW has read every dz element: da may now overwrite dz in place
    // ---- da[p][i] = sum_o W[o][i] dz[p][o]; the G threads of a pixel sit in one warp
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int ppw = 32 / G;
      const int pl = div_g(lane), ig = lane - pl * G;
      for (int p0 = warp * ppw; p0 < TP; p0 += 8 * ppw) {
        const int p = p0 + pl;
        const bool act = pl < ppw && p < TP;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (act) {
          if (PERM) {
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = d_s[p * C + sinv[ig * 4 + u]];
          } else {
#pragma unroll 4
            for (int o = 0; o < C; o += 4) {
              const float4 d4 = *reinterpret_cast<const float4*>(d_s + p * C + o);
              const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 wv = *reinterpret_cast<const float4*>(ws + (o + u) * C + ig * 4);
                acc[0] = fmaf(wv.x, dd[u], acc[0]); acc[1] = fmaf(wv.y, dd[u], acc[1]);
                acc[2] = fmaf(wv.z, dd[u], acc[2]); acc[3] = fmaf(wv.w, dd[u], acc[3]);
              }
            }
          }
        }
        __syncwarp();
        if (act) *reinterpret_cast<float4*>(d_s + p * C + ig * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- per-channel reductions over the tile's pixels (registers; flushed once per CTA)
    if (has_an && red_ch < chunks) {
      for (int p = red_ch; p < TP; p += chunks) {
        const float d = d_s[p * C + red_i];
        sg_acc += d; sga_acc = fmaf(d, a_s[p * C + red_i], sga_acc);
      }
    }
    // ---- dx = da * s, written back as whole pixels
    if (slot < ppb) {
      const float4 s4 = *reinterpret_cast<const float4*>(sc + quad * 4);
      for (int p = slot; p < TP; p += ppb) {
        const int pix = g0 + p;
        if (pix < NP) {
          const float4 d4 = *reinterpret_cast<const float4*>(d_s + p * C + quad * 4);
          *reinterpret_cast<float4*>(dx + (int64_t)pix * C + quad * 4) =
              make_float4(d4.x * s4.x, d4.y * s4.y, d4.z * s4.z, d4.w * s4.w);
        }
      }
    }
    __syncthreads();                           // the tile buffers and the window are free for the next tile
  }
  // ---- one round of global atomics per CTA
  if (has_an && red_ch < chunks) {
    atomicAdd(&s_red[red_i], sg_acc);
    atomicAdd(&s_red[C + red_i], sga_acc);
  }
  if (!PERM) {
    if (RMB_MAXIT == 1) {
      // thread `it` parks its 4x4 block in the (dead) window: scratch[it][16]; then one thread per dW element sums
      // the psplit copies in a fixed order and issues the CTA's single global atomic for it
      if (tid < nitems) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<float4*>(win + tid * 16 + u * 4) =
              make_float4(dwacc[0][u * 4], dwacc[0][u * 4 + 1], dwacc[0][u * 4 + 2], dwacc[0][u * 4 + 3]);
      }
      __syncthreads();
      for (int e = tid; e < C * C; e += 256) {
        const int o = CT ? e / (CT ? CT : 1) : e / C, i = e - o * C;
        const int blk = (o >> 2) * G + (i >> 2), sub = (o & 3) * 4 + (i & 3);
        float acc = 0.f;
        for (int ps = 0; ps < psplit; ++ps) acc += win[(ps * nb4 + blk) * 16 + sub];
        atomicAdd(dw + e, acc);
      }
    } else {
      for (int e = tid; e < C * C; e += 256) ws[e] = 0.f;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < RMB_MAXIT; ++k) {
        const int it = tid + k * 256;
        if (it < nitems) {
          const int ps = it / nb4, blk = it - ps * nb4;
          const int bo = div_g(blk), bi = blk - bo * G;
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) atomicAdd(&ws[(bo * 4 + u) * C + bi * 4 + v], dwacc[k][u * 4 + v]);
        }
      }
      __syncthreads();
      for (int e = tid; e < C * C; e += 256) atomicAdd(dw + e, ws[e]);
    }
  }
  __syncthreads();
  if (has_an)
    for (int c = tid; c < C; c += 256) {
      atomicAdd(dbias + c, s_red[c] * sc[c]);
      atomicAdd(dlogs + c, f * s_red[C + c]);
    }
  // ---- sample-independent logdet terms (one CTA)
  if (blockIdx.x == 0 && dld) {
    float a = 0.f;
    for (int n = tid; n < Nld; n += 256) a += dld[n];
    const float tot = block_sum(a, red);
    if (tid == 0) red[32] = tot * (float)HW;
    __syncthreads();
    const float Gs = red[32];
    if (has_an)
      for (int c = tid; c < C; c += 256) atomicAdd(dlogs + c, f * Gs);
    if (!PERM && winv)
      for (int e = tid; e < C * C; e += 256) {
        const int o = e / C, i = e - o * C;
        atomicAdd(dw + e, Gs * winv[i * C + o]);
      }
  }
}

// ------------------------------------------------------------------------------------------
// GaussianDiag.logp on rows (module.py:437-467): one CTA per sample.  h == null => N(0, I).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_gaussian_logp_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ x, int64_t ldx,
                          int HW, int c0, int Cz, const float* __restrict__ logdet_in,
                          float* __restrict__ logdet_out) {
  pdl_wait();
  __shared__ float red[32];
  const int64_t n = blockIdx.x;
  const float log2pi = 1.8378770664093453f;
  float acc = 0.f;
  const int cnt = Cz * HW;
  for (int e = threadIdx.x; e < cnt; e += 256) {
    const int p = e / Cz, j = e - p * Cz;
    const float v = x[(n * HW + p) * ldx + c0 + j];
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

// Split2d reverse (module.py:482-483, 532-536): out[p] = cat(z1[p], mean + exp(logs) * eps); eps is NCHW
// (torch's generator fills it in the reference's tensor order).
__global__ void rows_split2d_sample_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ z1,
                                           int64_t ldz1, const float* __restrict__ eps, float* __restrict__ out,
                                           int64_t total, int Ch, int HW) {
  pdl_wait();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int C = 2 * Ch;
  const int64_t pix = e / C;
  const int c = (int)(e - pix * C);
  float v;
  if (c < Ch) {
    v = z1[pix * ldz1 + c];
  } else {
    const int j = c - Ch;
    const int64_t n = pix / HW, p = pix - n * HW;
    const float2 ml = *reinterpret_cast<const float2*>(h + pix * ldh + 2 * j);
    v = ml.x + expf(ml.y) * eps[(n * Ch + j) * HW + p];
  }
  out[e] = v;
}

// Split2d backward (module.py:526-530), channels C/2.. of dx and du rows; channels 0..C/2-1 of dx are
// written by the un-squeeze of the next level's gradient (glowk_rows_squeeze) before this kernel.
__global__ void __launch_bounds__(256)
rows_split2d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ hrows, int64_t ldh,
                        const float* __restrict__ dld, const float* __restrict__ logs_p, float f,
                        float* __restrict__ dx, float* __restrict__ du, int64_t ldu,
                        float* __restrict__ dlogs_p, float* __restrict__ dbias_p, int64_t NP, int C, int HW,
                        int iters) {
  pdl_wait();
  __shared__ float s_part[4][256];
  __shared__ float s_e[ROWS_MAX_C_WIDE];
  const int Ch = C >> 1;
  const int tid = threadIdx.x;
  const int ppb = 256 / Ch;
  for (int c = tid; c < C; c += 256) s_e[c] = expf(logs_p[c] * f);
  __syncthreads();
  const int j = tid % Ch, slot = tid / Ch;
  float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
  if (slot < ppb) {
    for (int it = 0; it < iters; ++it) {
      const int64_t pix = ((int64_t)blockIdx.x * iters + it) * ppb + slot;
      if (pix >= NP) break;
      const float g = dld[pix / HW];
      const float2 hv = *reinterpret_cast<const float2*>(hrows + pix * ldh + 2 * j);
      const float mean = hv.x, lg = hv.y;
      const float d = x[pix * C + Ch + j] - mean;
      const float iv = 1.f / expf(2.f * lg);
      dx[pix * C + Ch + j] = -g * d * iv;
      const float dm = g * d * iv;
      const float dl = g * (d * d * iv - 1.f);
      *reinterpret_cast<float2*>(du + pix * ldu + 2 * j) = make_float2(dm * s_e[2 * j], dl * s_e[2 * j + 1]);
      a0 += dm * mean; b0 += dm; a1 += dl * lg; b1 += dl;
    }
  }
  s_part[0][tid] = a0; s_part[1][tid] = b0; s_part[2][tid] = a1; s_part[3][tid] = b1;
  __syncthreads();
  for (int t = tid; t < 4 * Ch; t += 256) {
    const int q = t / Ch, jj = t - q * Ch;
    float s = 0.f;
    for (int k = 0; k < ppb; ++k) s += s_part[q][k * Ch + jj];
    const int c = 2 * jj + (q >> 1);
    if ((q & 1) == 0) atomicAdd(dlogs_p + c, f * s);
    else atomicAdd(dbias_p + c, s_e[c] * s);
  }
}

// ------------------------------------------------------------------------------------------
// Squeeze2d / unsqueeze (module.py:551-591, bit-exact) between any two layouts.  Logical map (factor f):
//   squeezed[n, c*f*f + fh*f + fw, i, j] = full[n, c, i*f+fh, j*f+fw]
// `full` is [N,C,H,W], `squeezed` is [N,C*f*f,H/f,W/f]; each side is NCHW (layout 0, batch stride
// `ld` elements) or rows (layout 1, row pitch `ld` floats: only the first channels of wider rows are
// touched, which is how Split2d's z1 / dz1 halves are read and written in place).  reverse = 0 reads
// `full` and writes `squeezed`; reverse = 1 the other way.  One thread per element of the destination.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t layout_off(int layout, int64_t ld, int64_t n, int c, int y, int x, int Cc, int Hh, int Ww) {
  return layout == 0 ? n * ld + ((int64_t)c * Hh + y) * Ww + x
                     : ((n * Hh + y) * (int64_t)Ww + x) * ld + c;
}

__global__ void rows_squeeze_kernel(const float* __restrict__ src, int src_layout, int64_t src_ld,
                                    float* __restrict__ dst, int dst_layout, int64_t dst_ld, int total,
                                    int C, int H, int W, int f, int reverse, const float* __restrict__ add) {
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int Cs = C * f * f, Hs = H / f, Ws = W / f;
  // decode e in the DESTINATION's own element order so that writes are coalesced (32-bit arithmetic)
  int n, c, y, x;     // logical coordinates in the destination tensor
  const int Cd = reverse ? C : Cs, Hd = reverse ? H : Hs, Wd = reverse ? W : Ws;
  int r = e;
  if (dst_layout == 0) {
    x = r % Wd; r /= Wd; y = r % Hd; r /= Hd; c = r % Cd; n = r / Cd;
  } else {
    c = r % Cd; r /= Cd; x = r % Wd; r /= Wd; y = r % Hd; n = r / Hd;
  }
  int64_t so, dofs;
  if (!reverse) {   // destination = squeezed (c = cf*f*f + fh*f + fw), source = full
    const int cf = c / (f * f), rr = c - cf * f * f, fh = rr / f, fw = rr - fh * f;
    so = layout_off(src_layout, src_ld, n, cf, y * f + fh, x * f + fw, C, H, W);
    dofs = layout_off(dst_layout, dst_ld, n, c, y, x, Cs, Hs, Ws);
  } else {          // destination = full, source = squeezed
    const int i = y / f, fh = y - i * f, jx = x / f, fw = x - jx * f;
    so = layout_off(src_layout, src_ld, n, c * f * f + fh * f + fw, i, jx, Cs, Hs, Ws);
    dofs = layout_off(dst_layout, dst_ld, n, c, y, x, C, H, W);
  }
  dst[dofs] = add ? src[so] + add[so] : src[so];      // add: the dequantisation noise (network/model.py:421-423)
}

// ------------------------------------------------------------------------------------------
// Batched conv-weight packing / gradient unpacking: one launch for every coupling network of a
// FlowModel (a train step otherwise launches ~6 pack and ~2 unpack kernels per FlowStep).
// ------------------------------------------------------------------------------------------
struct PackJob {           // 48 bytes, mirrored by pytorch_glow_b200/rows_path.py
  const float* w;          // fp32 [O][I][k][k]  (unpack: the gradient, accumulated into)
  void* packed;            // GEMM-layout copy   (unpack: fp32 source)
  int32_t O, I, ks, layout;
  int32_t rows, ld;
  int64_t block0;          // first CTA of this job (pack: one CTA per 32x32 channel tile; unpack: 256 elements per CTA)
};

__device__ __forceinline__ int find_job(const PackJob* jobs, int njobs, int64_t blk) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ bool packed_coords(int layout, int O, int I, int T2, int64_t r, int64_t col, int* o, int* i, int* tap) {
  if (layout == 0) {            // packed[o][tap*I + i]
    if (r < O && col < (int64_t)T2 * I) { *o = (int)r; *tap = (int)(col / I); *i = (int)(col - (int64_t)*tap * I); return true; }
  } else if (layout == 1) {     // packed[tap*O + o][i]
    if (r < (int64_t)T2 * O && col < I) { *tap = (int)(r / O); *o = (int)(r - (int64_t)*tap * O); *i = (int)col; return true; }
  } else if (layout == 2) {     // packed[tap*I + i][o]
    if (r < (int64_t)T2 * I && col < O) { *tap = (int)(r / I); *i = (int)(r - (int64_t)*tap * I); *o = (int)col; return true; }
  } else {                      // packed[i][tap*O + o]
    if (r < I && col < (int64_t)T2 * O) { *i = (int)r; *tap = (int)(col / O); *o = (int)(col - (int64_t)*tap * O); return true; }
  }
  return false;
}

// One CTA packs a 32 (out channels) x 32 (in channels) x k*k tile: coalesced reads of the fp32 parameter (the
// 32*k*k values of an output channel's tile are contiguous), staged in shared memory, written as 64-byte runs
// of whichever index is contiguous in the destination layout.  Padding rows / columns of the destination are
// never written: the arena they live in is zero-filled once (rows_path.PackPlan).
constexpr int PK_T = 32;

template <typename T>
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const PackJob* __restrict__ jobs, int njobs) {
  pdl_wait();
  __shared__ int s_job;
  __shared__ float tile[PK_T * (PK_T * 9 + 1)];
  if (threadIdx.x == 0) s_job = find_job(jobs, njobs, blockIdx.x);
  __syncthreads();
  const PackJob jb = jobs[s_job];
  const int T2 = jb.ks * jb.ks;
  const int tiles_i = (jb.I + PK_T - 1) / PK_T;
  const int tb = (int)(blockIdx.x - jb.block0);
  const int o0 = (tb / tiles_i) * PK_T, i0 = (tb % tiles_i) * PK_T;
  const int no = min(PK_T, jb.O - o0), ni = min(PK_T, jb.I - i0);
  const int run = ni * T2;                          // contiguous source floats per output channel
  const int ostride = PK_T * T2 + 1;                // +1: conflict-free when o is the fast index below
  for (int e = threadIdx.x; e < no * run; e += 256) {
    const int o = e / run, r = e - o * run;
    tile[o * ostride + r] = jb.w[((int64_t)(o0 + o) * jb.I + i0) * T2 + r];
  }
  __syncthreads();
  T* dst = reinterpret_cast<T*>(jb.packed);
  const int64_t ld = jb.ld;
  if (jb.layout == 0 || jb.layout == 1) {           // i contiguous: dst[o][tap*I + i] / dst[tap*O + o][i]
    for (int e = threadIdx.x; e < no * T2 * ni; e += 256) {
      const int i = e % ni, q = e / ni, tap = q % T2, o = q / T2;
      const float v = tile[o * ostride + i * T2 + tap];
      const int64_t d = jb.layout == 0 ? (int64_t)(o0 + o) * ld + (int64_t)tap * jb.I + i0 + i
                                       : ((int64_t)tap * jb.O + o0 + o) * ld + i0 + i;
      dst[d] = from_f32<T>(v);
    }
  } else {                                          // o contiguous: dst[tap*I + i][o] / dst[i][tap*O + o]
    for (int e = threadIdx.x; e < no * T2 * ni; e += 256) {
      const int o = e % no, q = e / no, tap = q % T2, i = q / T2;
      const float v = tile[o * ostride + i * T2 + tap];
      const int64_t d = jb.layout == 2 ? ((int64_t)tap * jb.I + i0 + i) * ld + o0 + o
                                       : (int64_t)(i0 + i) * ld + (int64_t)tap * jb.O + o0 + o;
      dst[d] = from_f32<T>(v);
    }
  }
}

__global__ void unpack_grads_batched_kernel(const PackJob* __restrict__ jobs, int njobs) {
  pdl_wait();
  __shared__ int s_job;
  if (threadIdx.x == 0) s_job = find_job(jobs, njobs, blockIdx.x);
  __syncthreads();
  const PackJob jb = jobs[s_job];
  const int64_t e = ((int64_t)blockIdx.x - jb.block0) * 256 + threadIdx.x;
  const int T2 = jb.ks * jb.ks;
  if (e >= (int64_t)jb.O * jb.I * T2) return;
  const int tap = (int)(e % T2);
  const int i = (int)((e / T2) % jb.I);
  const int o = (int)(e / ((int64_t)T2 * jb.I));
  int64_t s;
  if (jb.layout == 0) s = (int64_t)o * jb.ld + (int64_t)tap * jb.I + i;
  else if (jb.layout == 1) s = ((int64_t)tap * jb.O + o) * jb.ld + i;
  else if (jb.layout == 2) s = ((int64_t)tap * jb.I + i) * jb.ld + o;
  else s = (int64_t)i * jb.ld + (int64_t)tap * jb.O + o;
  float* g = const_cast<float*>(jb.w);
  g[e] += reinterpret_cast<const float*>(jb.packed)[s];
}

// ------------------------------------------------------------------------------------------
// ActNorm-after-conv gradient finish (Conv2d = conv -> ActNorm -> ReLU, module.py:188-260), batched over layers.
// With y = (W a + b) * s on the active set and v = g * s the gradient the dgrad epilogue stores (dW = v^T a,
// db = sum v), the logs gradient sum_m g*y is EXACTLY  <W[n,:], dW[n,:]> + b[n]*db[n]  per output channel n.
// The tcgen05 ReLU-backward epilogue is issue-bound; dropping its second transposing column sum (g*y) and
// recovering dlogs here costs one pass over W and this step's dW.  W is the bf16 copy the forward GEMM multiplied
// by, dW / db this backward pass's own contributions (scratch, zeroed per pass), so gradient accumulation over
// several passes stays correct:  dbias[n] += db[n];  dlogs[n] += f * (<W[n,:], dW[n,:]> + b[n]*db[n]).
// ------------------------------------------------------------------------------------------
struct FinishJob {           // 72 bytes, mirrored by pytorch_glow_b200/rows_path.py
  const __nv_bfloat16* w;    // [N][ldw] bf16, k-order of dw
  const float* dw;           // [N][lddw] fp32, this pass only
  const float* bias;         // [N] ActNorm bias
  const float* db;           // [N] this pass's sum of v
  float* dbias;              // [N] accumulated into
  float* dlogs;              // [N] accumulated into
  int32_t N, K, ldw, lddw;
  float f;
  int32_t db_stride;         // db[n * db_stride]: 1 for a vector, lddw when db is a (ones-)column of dw
};

__global__ void __launch_bounds__(256)
conv_actnorm_finish_kernel(const FinishJob* __restrict__ jobs) {
  pdl_wait();
  const FinishJob jb = jobs[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y * 8 + warp;
  if (n >= jb.N) return;
  const __nv_bfloat16* wr = jb.w + (int64_t)n * jb.ldw;
  const float* dr = jb.dw + (int64_t)n * jb.lddw;
  float acc = 0.f;
  for (int k = lane * 2; k < jb.K; k += 64) {              // K, ldw, lddw are even (GEMM pitches)
    const float2 wv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(wr + k));
    const float2 dv = *reinterpret_cast<const float2*>(dr + k);
    acc = fmaf(wv.x, dv.x, acc);
    acc = fmaf(wv.y, dv.y, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float db = jb.db[(int64_t)n * jb.db_stride];
    jb.dbias[n] += db;
    jb.dlogs[n] += jb.f * (acc + jb.bias[n] * db);
  }
}

// ------------------------------------------------------------------------------------------
// ActNorm adjoint on rows for the wide levels (C > ROWS_MAX_C), where the 1x1 conv and its adjoint run as fp32
// GEMMs:  dx = da*s ; dbias += sum_p da*s ; dlogs += f * sum_p da*(x+b)*s   (module.py:34-84).  dx may alias da.
// grid = (channel blocks of 256, row chunks); one atomic per channel, quantity and CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_actnorm_bwd_kernel(const float* __restrict__ da, const float* __restrict__ x, const float* __restrict__ bias,
                        const float* __restrict__ logs, float f, float* dx, float* __restrict__ dlogs,
                        float* __restrict__ dbias, int P, int C, int rows_per_cta) {
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const float s = expf(logs[c] * f), b = bias[c];
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, P);
  float sg = 0.f, sga = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int64_t e = (int64_t)r * C + c;
    const float v = da[e];
    const float a = (x[e] + b) * s;
    dx[e] = v * s;
    sg += v; sga = fmaf(v, a, sga);
  }
  atomicAdd(dbias + c, sg * s);
  atomicAdd(dlogs + c, f * sga);
}

}  // namespace glowk

using namespace glowk;

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" int glowk_rows_actnorm_bwd(const float* da, const float* x, const float* bias, const float* logs,
                                      float logscale_factor, float* dx, float* dlogs, float* dbias, int64_t P,
                                      int64_t C, void* stream) {
  if (P == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(da && x && bias && logs && dx && dlogs && dbias, "glowk_rows_actnorm_bwd: null pointer");
  GLOWK_CHECK_ARG(C > 0 && P > 0 && P * C < (1ll << 31), "glowk_rows_actnorm_bwd: bad shape");
  int rows_per_cta = 32;
  while (ceil_div(P, rows_per_cta) > 65535) rows_per_cta *= 2;
  const dim3 grid((unsigned)ceil_div(C, 256), (unsigned)ceil_div(P, rows_per_cta));
  rows_actnorm_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(da, x, bias, logs, logscale_factor, dx, dlogs, dbias,
                                                                  (int)P, (int)C, rows_per_cta);
  GLOWK_CHECK_LAUNCH("glowk_rows_actnorm_bwd");
  return GLOWK_OK;
}

extern "C" int glowk_conv_actnorm_finish_batched(const void* jobs, int64_t njobs, int64_t max_n, void* stream) {
  if (njobs == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(jobs && njobs > 0 && njobs < 65536 && max_n > 0 && max_n < (1 << 20), "glowk_conv_actnorm_finish_batched: bad arguments");
  const dim3 grid((unsigned)njobs, (unsigned)ceil_div(max_n, 8));
  GLOWK_CUDA(launch_pdl(conv_actnorm_finish_kernel, grid, 256, 0, (cudaStream_t)stream, (const FinishJob*)jobs));
  GLOWK_CHECK_LAUNCH("glowk_conv_actnorm_finish_batched");
  return GLOWK_OK;
}

// pixel groups per CTA of the adjoint kernels: one resident wave whose CTAs loop, rounded UP so that no partial
// second wave is left
static inline int bwd_iters(int64_t NP, int ppb, int resident) {
  int64_t it = ceil_div(ceil_div(NP, ppb), resident);
  if (it < 1) it = 1;
  if (it > 32) it = 32;
  return (int)it;
}

extern "C" int glowk_rows_max_channels(void) { return ROWS_MAX_C; }
extern "C" int glowk_rows_max_channels_wide(void) { return ROWS_MAX_C_WIDE; }

extern "C" int glowk_rows_actnorm_mix(const float* x, float* z, const float* w, const int64_t* idx,
                                      const float* bias, const float* logs, float logscale_factor, int64_t P,
                                      int64_t C, int reverse, void* stream) {
  if (P == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && z && x != z, "glowk_rows_actnorm_mix: bad pointers");
  GLOWK_CHECK_ARG((w != nullptr) != (idx != nullptr), "glowk_rows_actnorm_mix: exactly one of w / idx");
  GLOWK_CHECK_ARG((bias != nullptr) == (logs != nullptr), "glowk_rows_actnorm_mix: bias and logs go together");
  GLOWK_CHECK_ARG(C > 0 && C % 4 == 0 && C <= ROWS_MAX_C, "glowk_rows_actnorm_mix: C=%lld must be a multiple of 4, <= %d", (long long)C, ROWS_MAX_C);
  GLOWK_CHECK_ARG((((uintptr_t)x | (uintptr_t)z) & 15) == 0, "glowk_rows_actnorm_mix: rows must be 16-byte aligned");
  GLOWK_CHECK_ARG(P * C < (1ll << 31), "glowk_rows_actnorm_mix: tensor too large for 32-bit indexing");
  const int G = (int)C / 4, ppb = 256 / G;
  const dim3 block((unsigned)G, (unsigned)ppb);
  cudaStream_t st = (cudaStream_t)stream;
  // one resident wave whose CTAs loop: rounding the loop count DOWN (and assuming 8 CTAs per SM for every
  // instance) left a second, nearly empty wave that doubled the kernel time (1234 CTAs on 1184 slots at C = 12)
#define GLOWK_MIX_LAUNCH(PERM_, CT_, SMEM_)                                                                          \
  do {                                                                                                               \
    auto kern = rows_mix_kernel<PERM_, CT_>;                                                                         \
    if (SMEM_ > 40 * 1024) GLOWK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
    const int64_t passes = ceil_div(P, 2 * ppb);             /* two pixels per thread and pass */                   \
    int iters = (int)ceil_div(passes, resident_ctas((const void*)kern, G * ppb, SMEM_));                             \
    iters = iters < 1 ? 1 : (iters > 32 ? 32 : iters);                                                               \
    const unsigned grid = (unsigned)ceil_div(passes, iters);                                                         \
    GLOWK_CUDA(launch_pdl(kern, grid, block, SMEM_, st, x, z, w, idx, bias, logs, logscale_factor, (int)P, (int)C,   \
                          reverse, iters));                                                                          \
  } while (0)
  const size_t xbuf_bytes = sizeof(float) * 2 * (size_t)(2 * ppb) * C;      // double-buffered pixel rows of a pass
  if (w) {
    const size_t smem = sizeof(float) * ((size_t)C * C + 3 * C) + xbuf_bytes;
    if (C == 12) GLOWK_MIX_LAUNCH(false, 12, smem);
    else if (C == 24) GLOWK_MIX_LAUNCH(false, 24, smem);
    else if (C == 48) GLOWK_MIX_LAUNCH(false, 48, smem);
    else GLOWK_MIX_LAUNCH(false, 0, smem);
  } else {
    const size_t smem = sizeof(float) * (3 * (size_t)C) + xbuf_bytes;
    GLOWK_MIX_LAUNCH(true, 0, smem);
  }
#undef GLOWK_MIX_LAUNCH
  GLOWK_CHECK_LAUNCH("glowk_rows_actnorm_mix");
  return GLOWK_OK;
}

extern "C" int glowk_rows_coupling_rev_mix(const float* P3, int64_t ldp, const float* bias3, const float* logs3,
                                           float logscale_factor3, const float* x, float* z, const float* w,
                                           const int64_t* idx, const float* bias, const float* logs,
                                           float logscale_factor, int64_t N, int64_t C, int64_t H, int64_t W, int affine,
                                           void* stream) {
  const int64_t P = N * H * W;
  if (P == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(P3 && bias3 && logs3 && x && z && x != z, "glowk_rows_coupling_rev_mix: bad pointers");
  GLOWK_CHECK_ARG((w != nullptr) != (idx != nullptr), "glowk_rows_coupling_rev_mix: exactly one of w / idx");
  GLOWK_CHECK_ARG((bias != nullptr) == (logs != nullptr), "glowk_rows_coupling_rev_mix: bias and logs go together");
  GLOWK_CHECK_ARG(C > 0 && C % 4 == 0 && C <= ROWS_MAX_C, "glowk_rows_coupling_rev_mix: C=%lld must be a multiple of 4, <= %d", (long long)C, ROWS_MAX_C);
  const int64_t Cout = affine ? C : C / 2;
  GLOWK_CHECK_ARG(ldp >= 9 * Cout && ldp % 2 == 0, "glowk_rows_coupling_rev_mix: ldp=%lld too small for 9*Cout=%lld", (long long)ldp, (long long)(9 * Cout));
  GLOWK_CHECK_ARG((((uintptr_t)x | (uintptr_t)z) & 15) == 0 && (((uintptr_t)P3) & 7) == 0, "glowk_rows_coupling_rev_mix: rows must be 16-byte aligned");
  GLOWK_CHECK_ARG(P * C < (1ll << 31) && P * ldp < (1ll << 31), "glowk_rows_coupling_rev_mix: tensor too large for 32-bit indexing");
  const int G = (int)C / 4, ppb = 256 / G;
  const dim3 block((unsigned)G, (unsigned)ppb);
  cudaStream_t st = (cudaStream_t)stream;
  const FastDiv dW = make_fastdiv(W), dHW = make_fastdiv(H * W), dCh = make_fastdiv(C / 2);
#define GLOWK_CRM_LAUNCH(PERM_, CT_, SMEM_)                                                                          \
  do {                                                                                                               \
    auto kern = rows_coupling_rev_mix_kernel<PERM_, CT_>;                                                            \
    if (SMEM_ > 40 * 1024) GLOWK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
    const int64_t passes = ceil_div(P, 2 * ppb);                                                                     \
    int iters = (int)ceil_div(passes, resident_ctas((const void*)kern, G * ppb, SMEM_));                             \
    iters = iters < 1 ? 1 : (iters > 32 ? 32 : iters);                                                               \
    const unsigned grid = (unsigned)ceil_div(passes, iters);                                                         \
    GLOWK_CUDA(launch_pdl(kern, grid, block, SMEM_, st, P3, (int)ldp, bias3, logs3, logscale_factor3, affine, (int)H, \
                          (int)W, dW, dHW, dCh, x, z, w, idx, bias, logs, logscale_factor, (int)P, (int)C, iters));  \
  } while (0)
  const size_t xbuf_bytes = sizeof(float) * 2 * (size_t)(2 * ppb) * C;
  if (w) {
    const size_t smem = sizeof(float) * ((size_t)C * C + 5 * C) + xbuf_bytes;
    if (C == 12) GLOWK_CRM_LAUNCH(false, 12, smem);
    else if (C == 24) GLOWK_CRM_LAUNCH(false, 24, smem);
    else if (C == 48) GLOWK_CRM_LAUNCH(false, 48, smem);
    else GLOWK_CRM_LAUNCH(false, 0, smem);
  } else {
    const size_t smem = sizeof(float) * (5 * (size_t)C) + xbuf_bytes;
    GLOWK_CRM_LAUNCH(true, 0, smem);
  }
#undef GLOWK_CRM_LAUNCH
  GLOWK_CHECK_LAUNCH("glowk_rows_coupling_rev_mix");
  return GLOWK_OK;
}

// pixel groups per CTA: enough to amortise the per-CTA reduction tail, few enough to keep >= ~4 CTAs per SM busy
static inline int coupling_iters(int64_t HW, int64_t C) {
  const int64_t groups = ceil_div(HW, 256 / (C / 2));
  return groups >= 16 ? 4 : (groups >= 4 ? 2 : 1);
}
extern "C" int64_t glowk_rows_coupling_nblk(int64_t HW, int64_t C) {
  return ceil_div(ceil_div(HW, 256 / (C / 2)), coupling_iters(HW, C));
}

extern "C" int glowk_rows_coupling(const float* P3, int64_t ldp, const float* bias3, const float* logs3,
                                   float logscale_factor, float* z, float* h_save, int64_t N, int64_t C, int64_t H,
                                   int64_t W, int affine, int reverse, const float* ld_in, float* ld_out,
                                   const float* an_logs, float an_logscale_factor, const float* logabsdet,
                                   float sign, float* partials, void* tickets, void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(P3 && bias3 && logs3 && z, "glowk_rows_coupling: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C % 2 == 0 && C <= ROWS_MAX_C_WIDE, "glowk_rows_coupling: bad channel count %lld", (long long)C);
  const int64_t Cout = affine ? C : C / 2;
  GLOWK_CHECK_ARG(ldp >= 9 * Cout && ldp % 2 == 0, "glowk_rows_coupling: ldp=%lld too small for 9*Cout=%lld", (long long)ldp, (long long)(9 * Cout));
  GLOWK_CHECK_ARG(!ld_out || (partials && tickets), "glowk_rows_coupling: logdet output needs partials and tickets");
  GLOWK_CHECK_ARG(N <= 65535 && H * W * C < (1ll << 30), "glowk_rows_coupling: shape out of range");
  GLOWK_CHECK_ARG(N * H * W * ldp < (1ll << 31), "glowk_rows_coupling: P3 too large for 32-bit tap offsets");
  const int Ch = (int)C / 2, ppb = 256 / Ch;
  const int64_t nblk_max = glowk_rows_coupling_nblk(H * W, C);
  // (a) shared-memory window kernel: T pixels per CTA, window of min(T + 2W + 2, HW) rows of ldp+4 floats, sized for
  // two CTAs per SM; taken when the window fits and the caller's `partials` covers the CTAs per sample
  // Measured on B200 (B = 512, gpurun_out/h3_*): SLOWER than the direct gather (62 vs 42 us per launch averaged
  // over the three levels): with ~100 KB of window only two CTAs fit per SM and each one waits for its whole window
  // before computing, so the copy and the arithmetic do not overlap.  Opt-in (GLOWK_COUPLING_WIN=1).
  static const bool use_win = getenv("GLOWK_COUPLING_WIN") != nullptr;
  if (use_win && ldp % 4 == 0 && (((uintptr_t)P3) & 15) == 0) {
    const int64_t HWl = H * W, pitch = ldp + 4;
    const int64_t rows_max = (100 * 1024) / (pitch * 4);
    const int64_t t_max = rows_max - 2 * W - 2;
    if (t_max >= 16 || rows_max >= HWl) {
      int64_t chunks = rows_max >= HWl ? 1 : ceil_div(HWl, t_max);
      // whole samples per CTA leave the SMs short of CTAs when the batch is small
      while (chunks < nblk_max && N * chunks < 2 * (int64_t)sm_count() && ceil_div(HWl, chunks + 1) >= 16) ++chunks;
      const int64_t T = ceil_div(HWl, chunks);
      chunks = ceil_div(HWl, T);
      const int64_t rows = (T + 2 * W + 2) < HWl ? (T + 2 * W + 2) : HWl;
      const size_t smem = (size_t)rows * pitch * 4;
      if (chunks <= nblk_max || !ld_out) {
        const int vec = (int)((9 * Cout + 3) / 4);
        // always the same value (>= any window this launcher builds): safe when several host threads launch
        GLOWK_CUDA(cudaFuncSetAttribute(rows_coupling_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        const dim3 gridw((unsigned)chunks, (unsigned)N);
        GLOWK_CUDA(launch_pdl(rows_coupling_win_kernel, gridw, 256, smem, (cudaStream_t)stream, P3, (int)ldp, bias3, logs3,
                              logscale_factor, z, h_save, (int)C, (int)H, (int)W, make_fastdiv(W), make_fastdiv(Ch),
                              make_fastdiv(vec), (int)T, (int)pitch, affine, reverse, ld_in, ld_out, an_logs,
                              an_logscale_factor, logabsdet, sign, partials, (unsigned int*)tickets));
        GLOWK_CHECK_LAUNCH("glowk_rows_coupling(window)");
        return GLOWK_OK;
      }
    }
  }
  // (b) direct-gather kernel.  CTAs per sample: the kernel is latency-bound (every pixel group is one round of
  // dependent loads), so its time is ~ waves x (groups per CTA + a constant for the per-CTA set-up and reduction
  // tail): pick the split that minimises that.
  const int64_t groups = ceil_div(H * W, ppb);
  const int64_t resident = resident_ctas((const void*)rows_coupling_kernel, 256, 0);
  int64_t nblk = 1, best_cost = -1;
  for (int64_t b = 1; b <= nblk_max; ++b) {
    const int64_t it = ceil_div(groups, b), eff = ceil_div(groups, it);
    const int64_t cost = ceil_div(N * eff, resident) * (it + 2);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; nblk = eff; }
  }
  const int iters = (int)ceil_div(groups, nblk);
  const dim3 grid((unsigned)nblk, (unsigned)N);
  GLOWK_CUDA(launch_pdl(rows_coupling_kernel, grid, 256, 0, (cudaStream_t)stream, P3, (int)ldp, bias3, logs3, logscale_factor, z, h_save,
                        (int)C, (int)H, (int)W, make_fastdiv(W), make_fastdiv(Ch), ppb, iters, affine, reverse, ld_in, ld_out,
                        an_logs, an_logscale_factor, logabsdet, sign, partials, (unsigned int*)tickets));
  GLOWK_CHECK_LAUNCH("glowk_rows_coupling");
  return GLOWK_OK;
}

extern "C" int glowk_rows_coupling_bwd(const float* y, const float* hrows, const float* dy, const float* dld,
                                       const float* logs3, float logscale_factor, float* dz, float* du,
                                       float* dlogs3, float* dbias3, int64_t N, int64_t C, int64_t HW, int affine,
                                       void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(y && hrows && dy && logs3 && dz && du && dlogs3 && dbias3, "glowk_rows_coupling_bwd: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C % 2 == 0 && C <= ROWS_MAX_C_WIDE, "glowk_rows_coupling_bwd: bad channel count");
  const int64_t NP = N * HW;
  GLOWK_CHECK_ARG(NP * C < (1ll << 31), "glowk_rows_coupling_bwd: tensor too large for 32-bit indexing");
  const int Ch = (int)(C / 2), ppb = 256 / Ch;
  const int iters = bwd_iters(NP, ppb, resident_ctas((const void*)rows_coupling_bwd_kernel, Ch * ppb, 0));
  const unsigned grid = (unsigned)ceil_div(NP, (int64_t)ppb * iters);
  GLOWK_CUDA(launch_pdl(rows_coupling_bwd_kernel, grid, dim3((unsigned)Ch, (unsigned)ppb), 0, (cudaStream_t)stream, y, hrows, dy, dld,
                        logs3, logscale_factor, dz, du, dlogs3, dbias3, (int)NP, (int)C, make_fastdiv(HW), affine, iters));
  GLOWK_CHECK_LAUNCH("glowk_rows_coupling_bwd");
  return GLOWK_OK;
}

extern "C" int glowk_rows_actnorm_mix_bwd(const float* x, const float* dz, const float* dA1, int64_t ld_a1,
                                          int64_t Cin, const float* w, const int64_t* idx, const float* bias,
                                          const float* logs, float logscale_factor, float* dx, float* dw,
                                          float* dlogs, float* dbias, int64_t N, int64_t C, int64_t H, int64_t W,
                                          const float* dld, const float* winv, void* stream) {
  return glowk_rows_actnorm_mix_bwd_ex(x, dz, dA1, GLOWK_F32, ld_a1, Cin, w, idx, bias, logs, logscale_factor, dx, dw,
                                       dlogs, dbias, N, C, H, W, dld, winv, stream);
}

extern "C" int glowk_rows_actnorm_mix_bwd_ex(const float* x, const float* dz, const void* dA1v, int da1_dtype,
                                             int64_t ld_a1, int64_t Cin, const float* w, const int64_t* idx,
                                             const float* bias, const float* logs, float logscale_factor, float* dx,
                                             float* dw, float* dlogs, float* dbias, int64_t N, int64_t C, int64_t H,
                                             int64_t W, const float* dld, const float* winv, void* stream) {
  const float* dA1 = (const float*)dA1v;
  GLOWK_CHECK_ARG(da1_dtype == GLOWK_F32 || da1_dtype == GLOWK_BF16, "glowk_rows_actnorm_mix_bwd: bad da1_dtype");
  const bool a_bf16 = dA1v && da1_dtype == GLOWK_BF16;
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && dz && dx, "glowk_rows_actnorm_mix_bwd: null pointer");
  GLOWK_CHECK_ARG((w != nullptr) != (idx != nullptr), "glowk_rows_actnorm_mix_bwd: exactly one of w / idx");
  GLOWK_CHECK_ARG(!w || dw, "glowk_rows_actnorm_mix_bwd: dw required with w");
  GLOWK_CHECK_ARG((bias != nullptr) == (logs != nullptr) && (!bias || (dlogs && dbias)), "glowk_rows_actnorm_mix_bwd: actnorm args");
  GLOWK_CHECK_ARG(C > 0 && C % 4 == 0 && C <= ROWS_MAX_C, "glowk_rows_actnorm_mix_bwd: bad channel count");
  GLOWK_CHECK_ARG(!dA1 || (Cin > 0 && Cin <= C && ld_a1 >= 9 * Cin), "glowk_rows_actnorm_mix_bwd: bad conv1 dgrad operand");
  GLOWK_CHECK_ARG(!dld || !w || winv, "glowk_rows_actnorm_mix_bwd: the logdet gradient of a 1x1 conv needs W^-1");
  GLOWK_CHECK_ARG(N < (1ll << 31), "glowk_rows_actnorm_mix_bwd: batch too large");
  const int64_t NP = N * H * W;
  GLOWK_CHECK_ARG(NP * C < (1ll << 31), "glowk_rows_actnorm_mix_bwd: tensor too large for 32-bit indexing");
  int TP = 128;                                                  // smaller tiles until every SM has two CTAs
  while (TP > 32 && ceil_div(NP, TP) < 2 * sm_count()) TP >>= 1;
  const size_t smem = sizeof(float) * (2 * (size_t)TP * C + (w ? (size_t)C * C : 0) + 5 * (size_t)C + 33);
  int64_t tiles = ceil_div(NP, TP);
  cudaStream_t st = (cudaStream_t)stream;
  GLOWK_CHECK_ARG(NP * (ld_a1 > C ? ld_a1 : C) < (1ll << 31), "glowk_rows_actnorm_mix_bwd: operands too large for 32-bit offsets");
  GLOWK_CHECK_ARG(!dA1 || (Cin % 2 == 0 && ld_a1 % 2 == 0 && ((uintptr_t)dA1) % 8 == 0), "glowk_rows_actnorm_mix_bwd: dA1 needs even Cin / pitch");
  GLOWK_CHECK_ARG((((uintptr_t)x | (uintptr_t)dz | (uintptr_t)dx) & 15) == 0, "glowk_rows_actnorm_mix_bwd: rows must be 16-byte aligned");
  const FastDiv dW_ = make_fastdiv(W), dHW = make_fastdiv(H * W), dG = make_fastdiv(C / 4);
  const bool big = (C / 4) * (C / 4) > 256;
  // (a) conv1-dgrad operand staged through a shared-memory window (see rows_mix_bwd_win_kernel)
  static const bool no_win = getenv("GLOWK_MIXBWD_NOWIN") != nullptr;        // A/B switch for profiling
  const int esz = a_bf16 ? 2 : 4;
  if (dA1 && (!no_win || a_bf16) && (ld_a1 * esz) % 16 == 0 && (((uintptr_t)dA1) & 15) == 0) {
    // window rows: `vec` 16-byte pieces of data, pitch an ODD number of 16-byte units (rows of neighbouring pixels
    // then start 4 banks apart instead of 0 / 16)
    const int vec = (int)((9 * Cin * esz + 15) / 16);
    const int pitch16 = (vec + 1) | 1;
    const int wpitch = pitch16 * 16 / esz;                                     // in dA1 elements
    int win_floats = 0;
    auto smem_for = [&](int tp) {
      win_floats = (int)((tp + 2 * W + 2) * pitch16 * 4);
      if (win_floats < 4096) win_floats = 4096;                            // also the dW combine scratch (256 x 16)
      return sizeof(float) * ((size_t)win_floats + 2 * (size_t)tp * C + (w ? (size_t)C * C : 0) + 5 * (size_t)C + 33);
    };
    // <= 56 KB per CTA and 64 registers: four CTAs per SM.  A/B at B = 512 (profiles/r1j_mixbwd_occupancy_ab.txt):
    // 61.4 / 42.2 / 43.1 us at levels 1 / 2 / 3 with three CTAs of <= 72 KB, 60.4 / 40.1 / 37.6 us with four.
    static const size_t smem_cap = []() { const char* e = getenv("GLOWK_MIXBWD_SMEM_KB"); return (size_t)(e ? atoi(e) : 56) * 1024; }();
    int TPw = 128;                          // smaller tiles until the cap holds and every SM has two CTAs
    while (TPw > 32 && (smem_for(TPw) > smem_cap || ceil_div(NP, TPw) < 2 * sm_count())) TPw >>= 1;
    const size_t smem_w = smem_for(TPw);
    if (smem_w <= 160 * 1024) {
      const int64_t tiles_w = ceil_div(NP, TPw);
      const FastDiv dVec = make_fastdiv(vec), dPair = make_fastdiv(Cin / 2);
#define GLOWK_RMBW_LAUNCH_T(PERM_, IT_, CT_, TA_)                                                                        \
  do {                                                                                                                   \
    auto kern = rows_mix_bwd_win_kernel<PERM_, IT_, CT_, TA_>;                                                           \
    GLOWK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));                     \
    const int64_t cap = resident_ctas((const void*)kern, 256, smem_w);                                                   \
    const unsigned grid = (unsigned)(tiles_w < cap ? tiles_w : cap);                                                     \
    GLOWK_CUDA(launch_pdl(kern, grid, 256, smem_w, st, x, dz, (const TA_*)dA1v, (int)ld_a1, (int)Cin, w, idx, bias,      \
                          logs, logscale_factor, dx, dw, dlogs, dbias, (int)NP, (int)C, (int)H, (int)W, dW_, dHW, dG,    \
                          dVec, dPair, TPw, wpitch, win_floats, dld, (int)N, winv));                                     \
  } while (0)
#define GLOWK_RMBW_LAUNCH(PERM_, IT_, CT_)                                                                               \
  do {                                                                                                                   \
    if (a_bf16) GLOWK_RMBW_LAUNCH_T(PERM_, IT_, CT_, __nv_bfloat16); else GLOWK_RMBW_LAUNCH_T(PERM_, IT_, CT_, float);   \
  } while (0)
      if (!w) { if (big) GLOWK_RMBW_LAUNCH(true, 3, 0); else GLOWK_RMBW_LAUNCH(true, 1, 0); }
      else if (big) GLOWK_RMBW_LAUNCH(false, 3, 0);
      else if (C == 12) GLOWK_RMBW_LAUNCH(false, 1, 12);
      else if (C == 24) GLOWK_RMBW_LAUNCH(false, 1, 24);
      else if (C == 48) GLOWK_RMBW_LAUNCH(false, 1, 48);
      else GLOWK_RMBW_LAUNCH(false, 1, 0);
#undef GLOWK_RMBW_LAUNCH
#undef GLOWK_RMBW_LAUNCH_T
      GLOWK_CHECK_LAUNCH("glowk_rows_actnorm_mix_bwd(window)");
      return GLOWK_OK;
    }
  }
  // (b) direct-gather kernel (fp32 operand only)
  if (a_bf16) return fail(GLOWK_EUNSUP, "glowk_rows_actnorm_mix_bwd: a bf16 conv1-dgrad operand needs the shared-memory window path (16-byte aligned rows, window <= 160 KB)");
#define GLOWK_RMB_LAUNCH(PERM_, IT_)                                                                                     \
  do {                                                                                                                   \
    auto kern = rows_mix_bwd_kernel<PERM_, IT_>;                                                                         \
    if (smem > 48 * 1024) GLOWK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int occ = 1;                                                                                                         \
    GLOWK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));                                    \
    if (occ < 1) occ = 1;                                                                                                \
    const int64_t cap = (int64_t)occ * sm_count();            /* persistent: every CTA resident, walks several tiles */ \
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);                                                         \
    GLOWK_CUDA(launch_pdl(kern, grid, 256, smem, st, x, dz, dA1, (int)ld_a1, (int)Cin, w, idx, bias, logs,               \
                          logscale_factor, dx, dw, dlogs, dbias, (int)NP, (int)C, (int)H, (int)W, dW_, dHW, dG, TP, dld, \
                          (int)N, winv));                                                                                \
  } while (0)
  if (w) { if (big) GLOWK_RMB_LAUNCH(false, 3); else GLOWK_RMB_LAUNCH(false, 1); }
  else { if (big) GLOWK_RMB_LAUNCH(true, 3); else GLOWK_RMB_LAUNCH(true, 1); }
#undef GLOWK_RMB_LAUNCH
  GLOWK_CHECK_LAUNCH("glowk_rows_actnorm_mix_bwd");
  return GLOWK_OK;
}

extern "C" int glowk_rows_gaussian_logp(const float* h, int64_t ldh, const float* x, int64_t ldx, int64_t N,
                                        int64_t HW, int64_t c0, int64_t Cz, const float* logdet_in,
                                        float* logdet_out, void* stream) {
  GLOWK_CHECK_ARG(x && logdet_out, "glowk_rows_gaussian_logp: null pointer");
  GLOWK_CHECK_ARG(c0 >= 0 && c0 + Cz <= ldx, "glowk_rows_gaussian_logp: channel window out of range");
  GLOWK_CHECK_ARG(!h || (ldh >= 2 * Cz && ldh % 2 == 0), "glowk_rows_gaussian_logp: ldh too small");
  GLOWK_CHECK_ARG(HW * Cz < (1ll << 31), "glowk_rows_gaussian_logp: sample too large");
  if (N == 0) return GLOWK_OK;
  rows_gaussian_logp_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(h, ldh, x, ldx, (int)HW, (int)c0, (int)Cz,
                                                                            logdet_in, logdet_out);
  GLOWK_CHECK_LAUNCH("glowk_rows_gaussian_logp");
  return GLOWK_OK;
}

extern "C" int glowk_rows_split2d_sample(const float* h, int64_t ldh, const float* z1, int64_t ldz1, const float* eps,
                                         float* out, int64_t N, int64_t Chalf, int64_t HW, void* stream) {
  GLOWK_CHECK_ARG(h && z1 && eps && out, "glowk_rows_split2d_sample: null pointer");
  GLOWK_CHECK_ARG(ldh >= 2 * Chalf && ldh % 2 == 0 && ldz1 >= Chalf, "glowk_rows_split2d_sample: pitches too small");
  const int64_t total = N * HW * 2 * Chalf;
  if (total == 0) return GLOWK_OK;
  rows_split2d_sample_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(h, ldh, z1, ldz1, eps, out, total, (int)Chalf, (int)HW);
  GLOWK_CHECK_LAUNCH("glowk_rows_split2d_sample");
  return GLOWK_OK;
}

extern "C" int glowk_rows_split2d_bwd(const float* x, const float* hrows, int64_t ldh, const float* dld,
                                      const float* logs_p, float logscale_factor, float* dx, float* du, int64_t ldu,
                                      float* dlogs_p, float* dbias_p, int64_t N, int64_t C, int64_t HW, void* stream) {
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(x && hrows && dld && logs_p && dx && du && dlogs_p && dbias_p, "glowk_rows_split2d_bwd: null pointer");
  GLOWK_CHECK_ARG(C > 0 && C % 2 == 0 && C <= ROWS_MAX_C_WIDE && ldh >= C && ldu >= C && ldu % 2 == 0 && ldh % 2 == 0, "glowk_rows_split2d_bwd: bad shape");
  const int64_t NP = N * HW;
  const int ppb = 256 / (int)(C / 2);
  const int iters = bwd_iters(NP, ppb, resident_ctas((const void*)rows_split2d_bwd_kernel, 256, 0));
  const unsigned grid = (unsigned)ceil_div(NP, (int64_t)ppb * iters);
  rows_split2d_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, hrows, ldh, dld, logs_p, logscale_factor, dx, du, ldu,
                                                                   dlogs_p, dbias_p, NP, (int)C, (int)HW, iters);
  GLOWK_CHECK_LAUNCH("glowk_rows_split2d_bwd");
  return GLOWK_OK;
}

static int rows_squeeze_launch(const char* who, const float* src, const float* add, int src_layout, int64_t src_ld,
                               float* dst, int dst_layout, int64_t dst_ld, int64_t N, int64_t C, int64_t H, int64_t W,
                               int factor, int reverse, void* stream) {
  (void)who;
  if (N == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(src && dst && src != dst, "glowk_rows_squeeze: bad pointers");
  GLOWK_CHECK_ARG(factor >= 1 && H % factor == 0 && W % factor == 0, "glowk_rows_squeeze: H,W not divisible by factor");   // module.py:588
  GLOWK_CHECK_ARG((src_layout | dst_layout) >= 0 && src_layout <= 1 && dst_layout <= 1, "glowk_rows_squeeze: layout must be 0 (NCHW) or 1 (rows)");
  const int64_t Cs = C * factor * factor;
  const int64_t c_src = reverse ? Cs : C, c_dst = reverse ? C : Cs;
  GLOWK_CHECK_ARG(src_ld >= (src_layout ? c_src : C * H * W) && dst_ld >= (dst_layout ? c_dst : C * H * W), "glowk_rows_squeeze: pitch too small");
  const int64_t total = N * C * H * W;
  if (total == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(total < (1ll << 31), "glowk_rows_squeeze: tensor too large for 32-bit indexing");
  rows_squeeze_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      src, src_layout, src_ld, dst, dst_layout, dst_ld, (int)total, (int)C, (int)H, (int)W, factor, reverse, add);
  GLOWK_CHECK_LAUNCH("glowk_rows_squeeze");
  return GLOWK_OK;
}

extern "C" int glowk_rows_squeeze(const float* src, int src_layout, int64_t src_ld, float* dst, int dst_layout,
                                  int64_t dst_ld, int64_t N, int64_t C, int64_t H, int64_t W, int factor, int reverse,
                                  void* stream) {
  return rows_squeeze_launch("glowk_rows_squeeze", src, nullptr, src_layout, src_ld, dst, dst_layout, dst_ld, N, C, H, W,
                             factor, reverse, stream);
}

extern "C" int glowk_rows_squeeze_add(const float* src, const float* add, int src_layout, int64_t src_ld, float* dst,
                                      int dst_layout, int64_t dst_ld, int64_t N, int64_t C, int64_t H, int64_t W,
                                      int factor, void* stream) {
  GLOWK_CHECK_ARG(add, "glowk_rows_squeeze_add: null pointer");
  return rows_squeeze_launch("glowk_rows_squeeze_add", src, add, src_layout, src_ld, dst, dst_layout, dst_ld, N, C, H, W,
                             factor, 0, stream);
}

extern "C" int glowk_pack_conv_weights_batched(const void* jobs, int64_t njobs, int64_t total_blocks, int act_dtype,
                                               void* stream) {
  if (njobs == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(jobs && njobs > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "glowk_pack_conv_weights_batched: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == GLOWK_BF16) pack_weights_batched_kernel<__nv_bfloat16><<<(unsigned)total_blocks, 256, 0, st>>>((const PackJob*)jobs, (int)njobs);
  else if (act_dtype == GLOWK_F32) pack_weights_batched_kernel<float><<<(unsigned)total_blocks, 256, 0, st>>>((const PackJob*)jobs, (int)njobs);
  else return fail(GLOWK_EINVAL, "glowk_pack_conv_weights_batched: bad act_dtype %d", act_dtype);
  GLOWK_CHECK_LAUNCH("glowk_pack_conv_weights_batched");
  return GLOWK_OK;
}

extern "C" int glowk_unpack_weight_grads_batched(const void* jobs, int64_t njobs, int64_t total_blocks, void* stream) {
  if (njobs == 0) return GLOWK_OK;
  GLOWK_CHECK_ARG(jobs && njobs > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "glowk_unpack_weight_grads_batched: bad arguments");
  unpack_grads_batched_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>((const PackJob*)jobs, (int)njobs);
  GLOWK_CHECK_LAUNCH("glowk_unpack_weight_grads_batched");
  return GLOWK_OK;
}

// 3x3 tap gather-sum on rows: dst[p][c0 + c] (+)= sum_tap P[nbr(p, tap)][tap*C + c]  (Split2d's conv dgrad).
namespace glowk {
__global__ void rows_tapsum_kernel(const float* __restrict__ P, int64_t ldp, float* __restrict__ dst, int64_t ld_dst,
                                   int64_t total, int c0, int C, int H, int W, int flip, int accumulate) {
  pdl_wait();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t pix = e / C;
  const int c = (int)(e - pix * C);
  const int HW = H * W;
  const int64_t n = pix / HW;
  const int q = (int)(pix - n * HW);
  const int yy = q / W, xx = q - yy * W;
  float r = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int tap = flip ? 8 - t : t;
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) r += P[(n * HW + (int64_t)sy * W + sx) * ldp + t * C + c];
  }
  float* d = dst + pix * ld_dst + c0 + c;
  *d = accumulate ? (*d + r) : r;
}
}  // namespace glowk

extern "C" int glowk_rows_tapsum(const float* P, int64_t ldp, float* dst, int64_t ld_dst, int64_t c0, int64_t C,
                                 int64_t N, int64_t H, int64_t W, int flip, int accumulate, void* stream) {
  GLOWK_CHECK_ARG(P && dst && ldp >= 9 * C && c0 >= 0 && c0 + C <= ld_dst, "glowk_rows_tapsum: bad arguments");
  const int64_t total = N * H * W * C;
  if (total == 0) return GLOWK_OK;
  glowk::rows_tapsum_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      P, ldp, dst, ld_dst, total, (int)c0, (int)C, (int)H, (int)W, flip, accumulate);
  GLOWK_CHECK_LAUNCH("glowk_rows_tapsum");
  return GLOWK_OK;
}
