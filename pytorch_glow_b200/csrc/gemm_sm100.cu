// tcgen05 / TMEM / TMA GEMMs for the coupling network (bf16 operands, fp32 accumulate), sm_100a.
//
//   gemm_tc_kernel : out[M][N] = epilogue(A[M][K] . B[N][K]^T)      A, B K-major (forward, dgrad)
//   wgrad_tc_kernel: dW[Mo][No] += A[P][Mo]^T . B[P][No]            A, B MN-major (weight gradients)
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor 2-D loads into a SWIZZLE_128B smem ring,
//                                completion on `full` mbarriers
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma (M=128, N<=256, K=16) into one of two
//                                TMEM accumulator stages; tcgen05.commit releases smem slots / publishes
//                                the accumulator
//   warps 2..5  epilogue      : tcgen05.ld (32 lanes x 32 columns) -> registers -> fused epilogue ->
//                                swizzled smem staging -> TMA store (or red.global for wgrad)
// The two accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>

#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace glowk {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // one 128-byte swizzle span of bf16
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int ACC_STAGE_COLS = 256;  // TMEM columns per accumulator stage (2 stages = 512 columns)
constexpr int MAX_STAGES = 8;
constexpr int STAGING_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)

struct Shared {
  uint64_t full_bar[MAX_STAGES];
  uint64_t empty_bar[MAX_STAGES];
  uint64_t tmem_full_bar[2];
  uint64_t tmem_empty_bar[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a launch error (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3ff) == 0 && globaltimer_ns() - t0 > 4000000000ull) {
      printf("glowk: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (sm_100).
//   K-major  operand: rows of 128 B (64 bf16 along K); SBO = 1024 B between 8-row groups; LBO unused (1).
//   MN-major operand: rows of 128 B (64 bf16 along M/N), one row per k; SBO = 1024 B between 8-k groups;
//                     LBO = byte distance between consecutive 64-element M/N blocks.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor: D=f32, A=B=bf16, M=128, N=n; a_major/b_major: 0 = K-major, 1 = MN-major.
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// Sum v[lane'] over the 32 lanes for every index: on return v[0] of lane l holds sum over lanes of v[l].
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

template <typename OutT>
struct OutBox {  // one TMA-store box = 32 rows x 128 bytes
  static constexpr int COLS = 128 / (int)sizeof(OutT);
};

// ---------------------------------------------------------------------------------------------
template <int EPI, typename OutT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_o, int M, int N, int K, int block_n, int num_stages,
               EpiParams ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = BLOCK_M * BLOCK_K * 2;
  const uint32_t b_bytes = (uint32_t)block_n * BLOCK_K * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* staging = smem + (size_t)num_stages * stage_bytes;
  Shared* sh = reinterpret_cast<Shared*>(staging + STAGING_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (N + block_n - 1) / block_n;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_a); prefetch_tensormap(&tm_b); prefetch_tensormap(&tm_o);
    for (int s = 0; s < num_stages; ++s) { mbar_init(&sh->full_bar[s], 1); mbar_init(&sh->empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh->tmem_full_bar[s], 1); mbar_init(&sh->tmem_empty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 512 columns (two accumulator stages), allocated and freed by this warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&sh->empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_arrive_expect_tx(&sh->full_bar[stage], stage_bytes);
          tma_load_2d(&tm_a, &sh->full_bar[stage], sa, kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(&tm_b, &sh->full_bar[stage], sa + a_bytes, kb * BLOCK_K, n_blk * block_n);
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(block_n, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&sh->tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_STAGE_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&sh->full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sa + a_bytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle span (>>4 => +2)
            tcgen05_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          tcgen05_commit(&sh->empty_bar[stage]);        // smem slot reusable once these MMAs retire
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&sh->tmem_full_bar[acc]);        // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4) =====================
    const int quarter = warp & 3;
    constexpr int BOX = OutBox<OutT>::COLS;
    uint8_t* my_stage = staging + (size_t)quarter * 2 * 4096;
    int buf = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
      const int row0 = m_blk * BLOCK_M + quarter * 32;
      const int64_t m = row0 + lane;
      mbar_wait(&sh->tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * ACC_STAGE_COLS;
      for (int c0 = 0; c0 < block_n; c0 += BOX) {
        const int ncol0 = n_blk * block_n + c0;
        if (ncol0 >= N) break;
        uint8_t* sbuf = my_stage + buf * 4096;
        // the TMA store that last read this staging buffer must have finished reading it
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int sub = 0; sub < BOX / 32; ++sub) {
          float v[32];
          tmem_ld32(t_row + (uint32_t)(c0 + sub * 32), v);
          const int nc = ncol0 + sub * 32;
          if (EPI == GLOWK_EPI_RELU_BWD) {
            float ga[32], gb[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { ga[j] = 0.f; gb[j] = 0.f; }
            if (m < M) epilogue_apply<EPI, 32>(ep, m, nc, N, v, ga, gb);
            const float sa = warp_colsum32(ga, lane);
            const float sb = warp_colsum32(gb, lane);
            if (nc + lane < N) epilogue_commit_colsums(ep, nc + lane, sa, sb);
          } else {
            float d0[32], d1[32];
            if (m < M) epilogue_apply<EPI, 32>(ep, m, nc, N, v, d0, d1);
          }
          // write this lane's row segment into the 128B-swizzled staging box
          uint8_t* srow = sbuf + lane * 128;
          if (sizeof(OutT) == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int chunk = j ^ (lane & 7);
              *reinterpret_cast<float4*>(srow + chunk * 16) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int chunk = (sub * 4 + j) ^ (lane & 7);
              __nv_bfloat162 h[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(v[8 * j + 2 * u], v[8 * j + 2 * u + 1]);
              *reinterpret_cast<uint4*>(srow + chunk * 16) = *reinterpret_cast<uint4*>(h);
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_o, sbuf, ncol0, row0);   // rows >= M and columns >= N are clipped by TMA
          tma_store_commit();
        }
        buf ^= 1;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient: dW[Mo][No] += sum_p A[p][mo] * B[p][no].  Both operands are read exactly as the
// forward pass stored them ([pixels][channels]) => MN-major UMMA operands.  Work item = (output tile,
// pixel chunk); partial tiles are accumulated with red.global.add.f32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int P,
                int Mo, int No, int block_n, int num_stages, int kb_per_item, float* __restrict__ dW, int64_t lddw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = BLOCK_M * BLOCK_K * 2;                 // two [64 k][64 m] boxes
  const uint32_t b_bytes = (uint32_t)block_n * BLOCK_K * 2;       // block_n/64 boxes
  const uint32_t stage_bytes = a_bytes + b_bytes;
  Shared* sh = reinterpret_cast<Shared*>(smem + (size_t)num_stages * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (Mo + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (No + block_n - 1) / block_n;
  const int total_kb = (P + BLOCK_K - 1) / BLOCK_K;
  const int k_items = (total_kb + kb_per_item - 1) / kb_per_item;
  const int num_items = m_blocks * n_blocks * k_items;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_a); prefetch_tensormap(&tm_b);
    for (int s = 0; s < num_stages; ++s) { mbar_init(&sh->full_bar[s], 1); mbar_init(&sh->empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh->tmem_full_bar[s], 1); mbar_init(&sh->tmem_empty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int tile = item / k_items, ki = item - tile * k_items;
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        const int kb0 = ki * kb_per_item;
        const int kb1 = min(kb0 + kb_per_item, total_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sh->empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_arrive_expect_tx(&sh->full_bar[stage], stage_bytes);
          for (int h = 0; h < BLOCK_M / 64; ++h)
            tma_load_2d(&tm_a, &sh->full_bar[stage], sa + h * 8192, m_blk * BLOCK_M + h * 64, kb * BLOCK_K);
          for (int h = 0; h < block_n / 64; ++h)
            tma_load_2d(&tm_b, &sh->full_bar[stage], sa + a_bytes + h * 8192, n_blk * block_n + h * 64, kb * BLOCK_K);
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(block_n, 1, 1);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int tile = item / k_items, ki = item - tile * k_items;
        const int kb0 = ki * kb_per_item;
        const int kb1 = min(kb0 + kb_per_item, total_kb);
        (void)tile;
        mbar_wait(&sh->tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_STAGE_COLS;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sh->full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          // MN-major: 64-element M/N blocks are 8192 B apart (LBO); 8-k groups 1024 B apart (SBO)
          const uint64_t adesc = make_smem_desc(sa, 8192, 1024);
          const uint64_t bdesc = make_smem_desc(sa + a_bytes, 8192, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 k-rows = 2048 bytes (>>4 => +128)
            tcgen05_mma_bf16(d_tmem, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc,
                             (kb > kb0) || (k != 0));
          }
          tcgen05_commit(&sh->empty_bar[stage]);
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&sh->tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / k_items;
      const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
      const int mrow = m_blk * BLOCK_M + quarter * 32 + lane;
      mbar_wait(&sh->tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * ACC_STAGE_COLS;
      for (int c0 = 0; c0 < block_n; c0 += 32) {
        const int nc = n_blk * block_n + c0;
        if (nc >= No) break;
        float v[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        if (mrow < Mo) {
          float* d = dW + (int64_t)mrow * lddw + nc;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nc + j < No) atomicAdd(d + j, v[j]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D row-major tensor [rows][cols] with leading dimension ld (elements); box = [box_rows][box_cols].
static int make_map_2d(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int elsize, uint64_t cols,
                       uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GLOWK_EUNSUP, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * (uint64_t)elsize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(GLOWK_ECUDA, "cuTensorMapEncodeTiled failed (%d): base=%p cols=%llu rows=%llu ld=%llu box=%ux%u", (int)r,
                base, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld, box_cols, box_rows);
  return GLOWK_OK;
}

static int pick_stages(size_t stage_bytes, size_t extra) {
  const size_t budget = 227 * 1024 - 1024 /*align slack*/ - sizeof(Shared) - 64 - extra;
  int s = (int)(budget / stage_bytes);
  if (s > MAX_STAGES) s = MAX_STAGES;
  return s;
}

template <int EPI, typename OutT>
static int launch_gemm_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, int M, int N, int K,
                          int block_n, int stages, const EpiParams& ep, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)stages * (BLOCK_M * BLOCK_K * 2 + (size_t)block_n * BLOCK_K * 2) + STAGING_BYTES + sizeof(Shared) + 64;
  auto kern = gemm_tc_kernel<EPI, OutT>;
  GLOWK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (int)(ceil_div(M, BLOCK_M) * ceil_div(N, block_n));
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, NUM_THREADS, smem, st>>>(ta, tb, to, M, N, K, block_n, stages, ep);
  GLOWK_CHECK_LAUNCH("glowk_gemm(tcgen05)");
  return GLOWK_OK;
}

}  // namespace tc

bool tc_available() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
  return major == 10 && tc::encode_fn() != nullptr;
}

int gemm_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                 int epilogue, const EpiParams& ep, void* out, int out_dtype, int64_t ldo, cudaStream_t st) {
  using namespace tc;
  GLOWK_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "glowk_gemm(bf16): lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
  GLOWK_CHECK_ARG(((uintptr_t)A | (uintptr_t)B | (uintptr_t)out) % 16 == 0, "glowk_gemm(bf16): operands must be 16-byte aligned");
  GLOWK_CHECK_ARG(M < (1ll << 31) && N < (1 << 20) && K < (1 << 24), "glowk_gemm(bf16): shape out of range");
  const int elo = out_dtype == GLOWK_BF16 ? 2 : 4;
  GLOWK_CHECK_ARG((ldo * elo) % 16 == 0, "glowk_gemm(bf16): output row pitch must be a multiple of 16 bytes");
  const int box_cols = 128 / elo;
  const int npad = (int)ceil_div(N, 16) * 16;
  const int n_blocks = (int)ceil_div(npad, 256);
  int block_n = n_blocks == 1 ? npad : (int)ceil_div(ceil_div(npad, n_blocks), box_cols) * box_cols;
  GLOWK_CHECK_ARG(block_n <= 256 && block_n % 16 == 0, "glowk_gemm(bf16): cannot tile N=%lld", (long long)N);
  const size_t stage_bytes = BLOCK_M * BLOCK_K * 2 + (size_t)block_n * BLOCK_K * 2;
  const int stages = pick_stages(stage_bytes, STAGING_BYTES);
  GLOWK_CHECK_ARG(stages >= 2, "glowk_gemm(bf16): not enough shared memory for a 2-stage pipeline");

  CUtensorMap ta, tb, to;
  int rc;
  if ((rc = make_map_2d(&ta, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BLOCK_K, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&tb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BLOCK_K, (uint32_t)block_n))) return rc;
  if ((rc = make_map_2d(&to, out, out_dtype == GLOWK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        elo, (uint64_t)N, (uint64_t)M, (uint64_t)ldo, (uint32_t)box_cols, 32))) return rc;

#define GLOWK_TC_CASE(E)                                                                                         \
  case E:                                                                                                        \
    return out_dtype == GLOWK_BF16                                                                               \
               ? launch_gemm_tc<E, __nv_bfloat16>(ta, tb, to, (int)M, (int)N, (int)K, block_n, stages, ep, st)   \
               : launch_gemm_tc<E, float>(ta, tb, to, (int)M, (int)N, (int)K, block_n, stages, ep, st);
  switch (epilogue) {
    GLOWK_TC_CASE(GLOWK_EPI_STORE)
    GLOWK_TC_CASE(GLOWK_EPI_ACTNORM_RELU)
    GLOWK_TC_CASE(GLOWK_EPI_ACTNORM)
    GLOWK_TC_CASE(GLOWK_EPI_ZEROS)
    GLOWK_TC_CASE(GLOWK_EPI_RELU_BWD)
  }
#undef GLOWK_TC_CASE
  return fail(GLOWK_EINVAL, "glowk_gemm: unknown epilogue %d", epilogue);
}

int wgrad_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t P, int64_t Mo, int64_t No,
                  float* dW, int64_t lddw, cudaStream_t st) {
  using namespace tc;
  // tensor-core path needs 64-wide channel blocks; anything else goes to the CUDA-core kernel
  if (Mo % 64 != 0 || No % 64 != 0 || lda % 8 != 0 || ldb % 8 != 0 || P >= (1ll << 31))
    return wgrad_simt(A, lda, B, ldb, GLOWK_BF16, P, Mo, No, dW, lddw, st);
  const int n_blocks = (int)ceil_div(No, 256);
  const int block_n = (int)ceil_div(ceil_div(No, n_blocks), 64) * 64;
  const size_t stage_bytes = BLOCK_M * BLOCK_K * 2 + (size_t)block_n * BLOCK_K * 2;
  const int stages = pick_stages(stage_bytes, 0);
  const int tiles = (int)(ceil_div(Mo, BLOCK_M) * n_blocks);
  const int total_kb = (int)ceil_div(P, BLOCK_K);
  int k_items = (int)ceil_div(2 * sm_count(), tiles);
  if (k_items > total_kb) k_items = total_kb;
  int kb_per_item = (int)ceil_div(total_kb, k_items);
  if (kb_per_item < 4) kb_per_item = total_kb < 4 ? total_kb : 4;
  k_items = (int)ceil_div(total_kb, kb_per_item);
  CUtensorMap ta, tb;
  int rc;
  // [P][Mo] row-major: inner dimension = channels (64-wide boxes), outer = pixels (64 per stage)
  if ((rc = make_map_2d(&ta, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)Mo, (uint64_t)P, (uint64_t)lda, 64, BLOCK_K))) return rc;
  if ((rc = make_map_2d(&tb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)No, (uint64_t)P, (uint64_t)ldb, 64, BLOCK_K))) return rc;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + sizeof(Shared) + 64;
  GLOWK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = tiles * k_items;
  const int grid = items < sm_count() ? items : sm_count();
  wgrad_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(ta, tb, (int)P, (int)Mo, (int)No, block_n, stages, kb_per_item, dW, lddw);
  GLOWK_CHECK_LAUNCH("glowk_gemm_wgrad(tcgen05)");
  return GLOWK_OK;
}

}  // namespace glowk
