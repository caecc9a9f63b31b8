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
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace glowk {
namespace tc {

constexpr int GEMM_THREADS = 320;         // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int EPI_WARPS = 8;
// The ReLU-backward epilogue is ~1200 dependent instructions per 32x64 chunk (mask, scale, two transposing
// shuffle reductions): with 8 warps it bounds the dgrad GEMMs (the MMA warp idles 76 % of the time at K = 128).
// That instance therefore runs 16 epilogue warps (4 per TMEM lane quarter) on 32-column chunks (<= 113 registers).
constexpr int EPI_WARPS_RB = 16;
constexpr int GEMM_THREADS_RB = 64 + 32 * EPI_WARPS_RB;
template <int EPI> struct EpiCfg { static constexpr int WARPS = EPI_WARPS, THREADS = GEMM_THREADS; };
template <> struct EpiCfg<GLOWK_EPI_RELU_BWD> { static constexpr int WARPS = EPI_WARPS_RB, THREADS = GEMM_THREADS_RB; };
constexpr int ACC_STAGE_COLS = 256;  // TMEM columns per accumulator stage (2 stages = 512 columns)
constexpr int MAX_STAGES = 8;
constexpr int STAGING_BYTES = 8 * 4096;      // one (32 rows x 128 B) TMA-store box per epilogue warp

struct Shared {
  uint64_t full_bar[MAX_STAGES];
  uint64_t empty_bar[MAX_STAGES];
  uint64_t tmem_full_bar[2];
  uint64_t tmem_empty_bar[2];
  uint64_t y_bar[16];
  uint32_t tmem_base;
};

// Profiling aid (GLOWK_GEMM_DEBUG bit 64): per-role wait-cycle counters of CTA 0, read back by glowk_debug_gemm_trace.
//   [0] producer: cycles waiting for free slots   [1] producer: total
//   [2] MMA: waiting for operands (full)          [3] MMA: waiting for a free accumulator   [4] MMA: total
//   [5] epilogue warp 0: waiting for accumulators [6] ... for the y box   [7] ... for the staging box (store read)
//   [8] epilogue warp 0: total                    [9] tiles of CTA 0
__device__ unsigned long long g_gemm_trace[16];
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, bool on, unsigned long long& acc) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}


constexpr int EPI_NMAX = 1024;   // column sums of RELU_BWD are kept in shared memory up to this N

// ---------------------------------------------------------------------------------------------
// out[M][N] = epilogue(A[M][K] . B[N][K]^T).  Optional thread-block cluster of CM x CN CTAs working on
// CM m-tiles x CN n-tiles of one "super tile": the A tile of an m-tile is needed by the CN CTAs of that
// row and the B tile of an n-tile by the CM CTAs of that column, so every CTA loads 1/CN of its A tile
// and 1/CM of its B tile and TMA-multicasts the slice to its row / column mates (L2 -> SM traffic per CTA
// drops from A+B to A/CN + B/CM).
// ---------------------------------------------------------------------------------------------
template <int EPI, typename OutT, bool PAIR>
__global__ void __launch_bounds__(EpiCfg<EPI>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_y, int M, int N,
               int K, int block_n, int num_stages, int CM, int CN, int dbg, EpiParams ep) {
  // PAIR is a template parameter on purpose: a kernel that contains cta_group::2 instructions can only be launched
  // as a cluster of an even number of CTAs ("cluster misconfiguration" otherwise)
  constexpr bool pair = PAIR;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t a_bytes = BLOCK_M * BLOCK_K * 2;
  // pair mode (CM = 2, CN = 1, cta_group::2): this CTA keeps only ITS half of the B tile (block_n/2 rows)
  const uint32_t b_bytes = (uint32_t)(pair ? block_n / 2 : block_n) * BLOCK_K * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* staging = smem + (size_t)num_stages * stage_bytes;
  uint8_t* ybuf = staging + STAGING_BYTES;                                        // RELU_BWD: one 4 KB y box per epilogue warp
  float* s_scale = reinterpret_cast<float*>(ybuf + (EPI == GLOWK_EPI_RELU_BWD ? STAGING_BYTES : 0));
  float* s_shift = s_scale + 256;
  float* s_gy = s_shift + 256;                                                    // RELU_BWD: [EPI_NMAX] x 2
  float* s_g = s_gy + EPI_NMAX;
  Shared* sh = reinterpret_cast<Shared*>(s_gy + (EPI == GLOWK_EPI_RELU_BWD ? 2 * EPI_NMAX : 0));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int csize = CM * CN;
  const uint32_t crank = csize > 1 ? cluster_ctarank() : 0;
  const int cm = (int)crank % CM, cn = (int)crank / CM;
  const int cluster_id = blockIdx.x / csize, num_clusters = gridDim.x / csize;
  const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (N + block_n - 1) / block_n;
  const int sup_n = (n_blocks + CN - 1) / CN;
  const int num_super = ((m_blocks + CM - 1) / CM) * sup_n;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  // multicast masks: row mates (same cm) receive my A slice, column mates (same cn) my B slice
  uint16_t mask_a = 0, mask_b = 0;
  for (int j = 0; j < CN; ++j) mask_a |= (uint16_t)(1u << (cm + CM * j));
  for (int i = 0; i < CM; ++i) mask_b |= (uint16_t)(1u << (i + CM * cn));
  const uint16_t mask_all = mask_a | mask_b;
  const bool small_n = N <= EPI_NMAX;

  pdl_trigger_entry();                       // the next kernel may be scheduled as SMs free up (it waits for us)
  if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("glowk: dynamic smem not 1024-aligned\n"); __trap(); }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_a); prefetch_tensormap(&tm_b); prefetch_tensormap(&tm_o);
    if (EPI == GLOWK_EPI_RELU_BWD) prefetch_tensormap(&tm_y);
    // pair mode: one multicast tcgen05.commit frees a slot in both CTAs; the leader's accumulator is free once the
    // epilogue warps of BOTH CTAs have drained it
    for (int s = 0; s < num_stages; ++s) { mbar_init(&sh->full_bar[s], 1); mbar_init(&sh->empty_bar[s], pair ? 1u : (uint32_t)(CM + CN - 1)); }
    constexpr int EW = EpiCfg<EPI>::WARPS;
    for (int s = 0; s < 2; ++s) { mbar_init(&sh->tmem_full_bar[s], 1); mbar_init(&sh->tmem_empty_bar[s], pair ? 2 * EW : EW); }
    for (int q = 0; q < EW; ++q) mbar_init(&sh->y_bar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 512 columns (two accumulator stages), allocated and freed by this warp
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (EPI == GLOWK_EPI_RELU_BWD)
    for (int i = threadIdx.x; i < 2 * EPI_NMAX; i += EpiCfg<EPI>::THREADS) s_gy[i] = 0.f;
  tcgen05_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();        // peers' barriers must be initialised before any remote arrive / multicast
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  pdl_wait();                                // everything above touched no global memory (PDL, common.cuh)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int a_rows = BLOCK_M / CN, b_rows = block_n / CM;
      const bool tr = (dbg & 64) && blockIdx.x == 0;
      unsigned long long w_empty = 0;
      const long long t_begin = clock64();
      for (int st = cluster_id; st < num_super && !(dbg & 8); st += num_clusters) {
        const int m_blk = (st / sup_n) * CM + cm, n_blk = (st % sup_n) * CN + cn;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_timed(&sh->empty_bar[stage], phase ^ 1, tr, w_empty);     // every CTA I write into has drained this slot
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if ((dbg & 1) && (st != cluster_id || kb >= num_stages)) {     // profiling aid: no operand traffic
            mbar_arrive(&sh->full_bar[stage]);
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
            continue;
          }
          if constexpr (PAIR) {
            // both CTAs' bytes are counted on the LEADER's barrier; the leader alone arms it (2 x stage_bytes)
            const uint32_t lead_full = mapa_cluster(&sh->full_bar[stage], 0);
            if (crank == 0) mbar_arrive_expect_tx(&sh->full_bar[stage], 2 * stage_bytes);
            tma_load_2d_pair(&tm_a, lead_full, sa, kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d_pair(&tm_b, lead_full, sa + a_bytes, kb * BLOCK_K, n_blk * block_n + cm * (block_n / 2));
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&sh->full_bar[stage], stage_bytes);
          if (csize == 1) {
            tma_load_2d(&tm_a, &sh->full_bar[stage], sa, kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(&tm_b, &sh->full_bar[stage], sa + a_bytes, kb * BLOCK_K, n_blk * block_n);
          } else {
            tma_load_2d_mc(&tm_a, &sh->full_bar[stage], sa + (size_t)cn * a_rows * 128, kb * BLOCK_K,
                           m_blk * BLOCK_M + cn * a_rows, mask_a);
            tma_load_2d_mc(&tm_b, &sh->full_bar[stage], sa + a_bytes + (size_t)cm * b_rows * 128, kb * BLOCK_K,
                           n_blk * block_n + cm * b_rows, mask_b);
          }
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
      }
      pdl_trigger_drain();
      if (tr) { g_gemm_trace[0] = w_empty; g_gemm_trace[1] = (unsigned long long)(clock64() - t_begin); }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (warp-uniform control flow: descriptors stay in uniform registers, no R2UR + vote
    // loop around every UTCHMMA); one elected lane issues.  Pair mode: the leader CTA issues for both.
    if (!(pair && crank != 0)) {
      // M field: 128 for one CTA, 256 for the pair
      const uint32_t idesc = make_idesc(block_n, 0, 0) + (pair ? ((uint32_t)(BLOCK_M >> 4) << 24) : 0u);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const bool tr = (dbg & 64) && blockIdx.x == 0;
      unsigned long long w_full = 0, w_acc = 0, ntile = 0;
      const long long t_begin = clock64();
      for (int st = cluster_id; st < num_super; st += num_clusters) {
        ++ntile;
        mbar_wait_timed(&sh->tmem_empty_bar[acc], acc_phase ^ 1, tr, w_acc);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_STAGE_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!(dbg & 8)) mbar_wait_timed(&sh->full_bar[stage], phase, tr, w_full);   // dbg 8: free-running MMA issue (raw tensor rate)
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sa + a_bytes, 16, 1024);
          if (elect_one_sync()) {
            if constexpr (PAIR) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                tcgen05_mma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              tcgen05_commit_pair(&sh->empty_bar[stage]);            // frees the slot in both CTAs
            } else {
              if (!(dbg & 2)) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle span (>>4 => +2)
                  tcgen05_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                }
              }
              // smem slot reusable once these MMAs retire: tell every CTA that writes into it
              if (dbg & 8) {}
              else if (csize == 1) tcgen05_commit(&sh->empty_bar[stage]);
              else tcgen05_commit_mc(&sh->empty_bar[stage], mask_all);
            }
          }
          __syncwarp();
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) {
          if constexpr (PAIR) tcgen05_commit_pair(&sh->tmem_full_bar[acc]);        // accumulator complete (both CTAs' halves)
          else tcgen05_commit(&sh->tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (tr && lane == 0) { g_gemm_trace[2] = w_full; g_gemm_trace[3] = w_acc; g_gemm_trace[4] = (unsigned long long)(clock64() - t_begin); g_gemm_trace[9] = ntile; }
    }
  } else if constexpr (EPI == GLOWK_EPI_RELU_BWD) {
    // ===================== ReLU-backward epilogue: 16 warps, 32-column chunks =====================
    // warp -> (TMEM lane quarter = warp % 4, column group = (warp-2)/4 in 0..3); chunk c0 = grp*32 + k*128.
    // y / output boxes are [32 rows][32 bf16] = 2 KB, SWIZZLE_64B: 16-byte piece j of row r lives at
    // r*64 + ((j ^ ((r >> 1) & 3)) * 16).  Rows >= M need no masking: the y box is zero-filled there (TMA OOB).
    constexpr int EW = EPI_WARPS_RB, BOX = 32, BOXB = 2048;
    const int quarter = warp & 3;
    const int ew = warp - 2, grp = ew >> 2;
    const int et = (int)threadIdx.x - 64;
    uint8_t* sbuf = staging + (size_t)ew * BOXB;
    uint8_t* my_y = ybuf + (size_t)ew * BOXB;
    const uint8_t* yrow = my_y + lane * 64;
    const int sw = (lane >> 1) & 3;
    const bool want_gy = ep.dlogs != nullptr;
    const bool want_gb = ep.dbias != nullptr;      // null: the caller takes the bias gradient from a ones column of its wgrad
    int acc = 0; uint32_t acc_phase = 0;
    int cached_nblk = -1;
    const bool tr = (dbg & 64) && blockIdx.x == 0 && ew == 0;
    unsigned long long w_tfull = 0, w_y = 0, w_st = 0;
    const long long t_begin = clock64();
    uint32_t yit = 0;                       // y boxes consumed so far by this warp (mbarrier parity)
    auto issue_y = [&](int st, int c0) {
      if (lane == 0) {
        const int m_blk = (st / sup_n) * CM + cm, n_blk = (st % sup_n) * CN + cn;
        mbar_arrive_expect_tx(&sh->y_bar[ew], BOXB);
        tma_load_2d(&tm_y, &sh->y_bar[ew], my_y, n_blk * block_n + c0, m_blk * BLOCK_M + quarter * 32);
      }
    };
    auto next_valid = [&](int st) {
      while (st < num_super && ((st % sup_n) * CN + cn) * block_n + grp * BOX >= N) st += num_clusters;
      return st;
    };
    {
      const int st0 = next_valid(cluster_id);
      if (st0 < num_super) issue_y(st0, grp * BOX);
    }
    for (int st = cluster_id; st < num_super; st += num_clusters) {
      const int m_blk = (st / sup_n) * CM + cm, n_blk = (st % sup_n) * CN + cn;
      const int row0 = m_blk * BLOCK_M + quarter * 32;
      if (n_blk != cached_nblk) {
        epi_bar_sync<32 * EW>();
        if (et < block_n) {
          const int n = n_blk * block_n + et;
          s_scale[et] = n < N ? expf(ep.logs[n] * ep.f) : 1.f;
        }
        epi_bar_sync<32 * EW>();
        cached_nblk = n_blk;
      }
      mbar_wait_timed(&sh->tmem_full_bar[acc], acc_phase, tr, w_tfull);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * ACC_STAGE_COLS;
      for (int c0 = grp * BOX; c0 < block_n; c0 += 4 * BOX) {
        const int ncol0 = n_blk * block_n + c0;
        if (ncol0 >= N) break;
        uint32_t raw[32];
        tmem_ld32_async(t_row + (uint32_t)c0, raw);
        mbar_wait_timed(&sh->y_bar[ew], yit & 1, tr, w_y);
        {
          const long long t0 = tr ? clock64() : 0;
          if (lane == 0) tma_store_wait_read<0>();      // the store that last read the staging box is done with it
          __syncwarp();
          if (tr) w_st += (unsigned long long)(clock64() - t0);
        }
        tmem_ld_wait();
        float v[32], ga[32];
        const float* scl = s_scale + c0;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const uint4 yraw = *reinterpret_cast<const uint4*>(yrow + ((j4 ^ sw) * 16));
          const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&yraw);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 yv = __bfloat1622float2(yp[u]);
            const int j = j4 * 8 + u * 2;
            const float g0 = yv.x > 0.f ? __uint_as_float(raw[j]) : 0.f;
            const float g1 = yv.y > 0.f ? __uint_as_float(raw[j + 1]) : 0.f;
            if (want_gy) { ga[j] = g0 * yv.x; ga[j + 1] = g1 * yv.y; }
            v[j] = g0 * scl[j]; v[j + 1] = g1 * scl[j + 1];
          }
        }
        // stage the output row (bf16) first: the reductions below consume v and ga in place
        uint8_t* srow = sbuf + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(v[8 * j + 2 * u], v[8 * j + 2 * u + 1]);
          *reinterpret_cast<uint4*>(srow + ((j ^ sw) * 16)) = *reinterpret_cast<uint4*>(h);
        }
        // y consumed: fetch this warp's next box (same tile or the next one) behind the reductions and the store
        ++yit;
        __syncwarp();
        {
          int nst = st, nc0 = c0 + 4 * BOX;
          if (nc0 >= block_n || n_blk * block_n + nc0 >= N) { nst = next_valid(st + num_clusters); nc0 = grp * BOX; }
          if (nst < num_super) issue_y(nst, nc0);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_o, sbuf, ncol0, row0);   // rows >= M and columns >= N are clipped by TMA
          tma_store_commit();
        }
        // column sums over this warp's 32 rows: dbias += sum of the scaled gradient, dlogs += f * sum g*y
        float sb = 0.f, sa = 0.f;
        if (want_gb) sb = warp_colsum32(v, lane);
        if (want_gy) sa = warp_colsum32(ga, lane);
        if ((want_gb || want_gy) && ncol0 + lane < N) {
          if (small_n) { if (want_gb) atomicAdd(&s_g[ncol0 + lane], sb); if (want_gy) atomicAdd(&s_gy[ncol0 + lane], sa); }
          else { if (want_gb) atomicAdd(ep.dbias + ncol0 + lane, sb); if (want_gy) atomicAdd(ep.dlogs + ncol0 + lane, ep.f * sa); }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(&sh->tmem_empty_bar[acc], 0));
        else mbar_arrive(&sh->tmem_empty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (small_n) {
      epi_bar_sync<32 * EW>();
      for (int n = et; n < N; n += 32 * EW) {
        const float a = s_gy[n], b = s_g[n];
        if (want_gb && b != 0.f) atomicAdd(ep.dbias + n, b);
        if (want_gy && a != 0.f) atomicAdd(ep.dlogs + n, ep.f * a);
      }
    }
    if (tr && lane == 0) {
      g_gemm_trace[5] = w_tfull; g_gemm_trace[6] = w_y; g_gemm_trace[7] = w_st;
      g_gemm_trace[8] = (unsigned long long)(clock64() - t_begin);
    }
    if (lane == 0) tma_store_wait_all();
  } else {
    // ===================== epilogue warps =====================
    // TMEM lane quarter = warp % 4 (hardware rule); the two warps of a quarter take alternate column chunks.
    const int quarter = warp & 3;
    const int ew = warp - 2, half = ew >> 2;
    const int et = (int)threadIdx.x - 64;
    constexpr int BOX = OutBox<OutT>::COLS;
    constexpr int NSUB = BOX / 32;
    uint8_t* sbuf = staging + (size_t)ew * 4096;
    uint8_t* my_y = ybuf + (size_t)ew * 4096;
    const uint8_t* yrow = my_y + lane * 128;
    int acc = 0; uint32_t acc_phase = 0;
    int cached_nblk = -1;
    const bool tr = (dbg & 64) && blockIdx.x == 0 && ew == 0;
    unsigned long long w_tfull = 0, w_y = 0, w_st = 0;
    const long long t_begin = clock64();
    uint32_t yit = 0;                       // RELU_BWD: y boxes consumed so far by this warp (mbarrier parity)
    auto issue_y = [&](int st, int c0) {
      if (lane == 0) {
        const int m_blk = (st / sup_n) * CM + cm, n_blk = (st % sup_n) * CN + cn;
        mbar_arrive_expect_tx(&sh->y_bar[ew], 4096);
        tma_load_2d(&tm_y, &sh->y_bar[ew], my_y, n_blk * block_n + c0, m_blk * BLOCK_M + quarter * 32);
      }
    };
    // first super tile at or after `st` in which this warp has a column chunk (ragged clusters / narrow N)
    auto next_valid = [&](int st) {
      while (st < num_super && ((st % sup_n) * CN + cn) * block_n + half * BOX >= N) st += num_clusters;
      return st;
    };
    if (EPI == GLOWK_EPI_RELU_BWD) {
      const int st0 = next_valid(cluster_id);
      if (st0 < num_super) issue_y(st0, half * BOX);
    }
    for (int st = cluster_id; st < num_super; st += num_clusters) {
      const int m_blk = (st / sup_n) * CM + cm, n_blk = (st % sup_n) * CN + cn;
      const int row0 = m_blk * BLOCK_M + quarter * 32;
      const int64_t m = row0 + lane;
      if (EPI != GLOWK_EPI_STORE && n_blk != cached_nblk) {
        // per-column scale = exp(f*logs), shift = bias*scale for this n-tile (ActNorm / Conv2dZeros epilogue)
        epi_bar_sync<32 * EPI_WARPS>();
        if (et < block_n) {
          const int n = n_blk * block_n + et;
          float sc = 1.f, sf = 0.f;
          if (n < N) { sc = expf(ep.logs[n] * ep.f); sf = ep.bias ? ep.bias[n] * sc : 0.f; }
          s_scale[et] = sc; s_shift[et] = sf;
        }
        epi_bar_sync<32 * EPI_WARPS>();
        cached_nblk = n_blk;
      }
      mbar_wait_timed(&sh->tmem_full_bar[acc], acc_phase, tr, w_tfull);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * ACC_STAGE_COLS;
      for (int c0 = half * BOX; c0 < block_n; c0 += 2 * BOX) {
        const int ncol0 = n_blk * block_n + c0;
        if (ncol0 >= N || (dbg & 4)) break;
        uint32_t raw[NSUB][32];
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) tmem_ld32_async(t_row + (uint32_t)(c0 + sub * 32), raw[sub]);
        if (EPI == GLOWK_EPI_RELU_BWD) mbar_wait_timed(&sh->y_bar[ew], yit & 1, tr, w_y);
        // the TMA store that last read the staging box must have finished reading it
        {
          const long long t0 = tr ? clock64() : 0;
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          if (tr) w_st += (unsigned long long)(clock64() - t0);
        }
        tmem_ld_wait();
        __syncwarp();
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[sub][j]);
          const int nc = ncol0 + sub * 32;
          const float* scl = s_scale + c0 + sub * 32;
          const float* sft = s_shift + c0 + sub * 32;
          if (EPI == GLOWK_EPI_RELU_BWD) {
            float ga[32], gb[32];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              // this lane's row of y: 16-byte chunk (sub*4 + j4) of the 128B-swizzled box
              const uint4 yraw = *reinterpret_cast<const uint4*>(yrow + (((sub * 4 + j4) ^ (lane & 7)) * 16));
              const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&yraw);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 yv = __bfloat1622float2(yp[u]);
                const int j = j4 * 8 + u * 2;
                const float g0 = yv.x > 0.f ? v[j] : 0.f, g1 = yv.y > 0.f ? v[j + 1] : 0.f;
                ga[j] = g0 * yv.x; ga[j + 1] = g1 * yv.y;
                gb[j] = g0; gb[j + 1] = g1;
                v[j] = g0 * scl[j]; v[j + 1] = g1 * scl[j + 1];
              }
            }
            if (m >= M) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { ga[j] = 0.f; gb[j] = 0.f; }
            }
            const float sa = warp_colsum32(ga, lane);
            const float sb = warp_colsum32(gb, lane);
            if (nc + lane < N) {
              if (small_n) { atomicAdd(&s_gy[nc + lane], sa); atomicAdd(&s_g[nc + lane], sb); }
              else epilogue_commit_colsums(ep, nc + lane, sa, sb);
            }
          } else if (EPI != GLOWK_EPI_STORE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float r = fmaf(v[j], scl[j], sft[j]);
              if (EPI == GLOWK_EPI_ACTNORM_RELU) r = fmaxf(r, 0.f);
              v[j] = r;
            }
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
        if (EPI == GLOWK_EPI_RELU_BWD) {
          // y box consumed: fetch this warp's next one (same tile or the next) behind the store below
          ++yit;
          __syncwarp();
          int nst = st, nc0 = c0 + 2 * BOX;
          if (nc0 >= block_n || n_blk * block_n + nc0 >= N) { nst = next_valid(st + num_clusters); nc0 = half * BOX; }
          if (nst < num_super) issue_y(nst, nc0);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_o, sbuf, ncol0, row0);   // rows >= M and columns >= N are clipped by TMA
          tma_store_commit();
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(&sh->tmem_empty_bar[acc], 0));     // the leader issues the MMAs
        else mbar_arrive(&sh->tmem_empty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (EPI == GLOWK_EPI_RELU_BWD && small_n) {
      // one global atomic per column per CTA: dlogs += f * sum g*y ; dbias += exp(f*logs) * sum g
      epi_bar_sync<32 * EPI_WARPS>();
      for (int n = et; n < N; n += 32 * EPI_WARPS) {
        const float a = s_gy[n], b = s_g[n];
        if (a != 0.f || b != 0.f) epilogue_commit_colsums(ep, n, a, b);
      }
    }
    if (tr && lane == 0) {
      g_gemm_trace[5] = w_tfull; g_gemm_trace[6] = w_y; g_gemm_trace[7] = w_st;
      g_gemm_trace[8] = (unsigned long long)(clock64() - t_begin);
    }
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();        // no CTA may exit while peers still multicast into it
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient: dW[Mo][No] += sum_p A[p][mo] * B[p][no].  Both operands are read exactly as the
// forward pass stored them ([pixels][channels]) => MN-major UMMA operands.  Work item = (output tile,
// pixel chunk); partial tiles are accumulated with red.global.add.f32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const __grid_constant__ CUtensorMap tm_d, int P, int Mo, int No, int block_n, int num_stages,
                int kb_per_item) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t a_bytes = BLOCK_M * BLOCK_K * 2;                 // two [64 k][64 m] boxes
  const uint32_t b_bytes = (uint32_t)block_n * BLOCK_K * 2;       // block_n/64 boxes
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* staging = smem + (size_t)num_stages * stage_bytes;     // one [32][32] fp32 box per epilogue warp
  Shared* sh = reinterpret_cast<Shared*>(staging + STAGING_BYTES);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int m_blocks = (Mo + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (No + block_n - 1) / block_n;
  const int total_kb = (P + BLOCK_K - 1) / BLOCK_K;
  const int k_items = (total_kb + kb_per_item - 1) / kb_per_item;
  const int num_tiles = m_blocks * n_blocks;
  const int num_items = num_tiles * k_items;
  // item = (pixel chunk ki, output tile) with the TILE fastest: the CTAs running concurrently work on the same
  // pixel chunk, so its A and B rows are fetched from HBM once and re-read from L2 by the other tiles (with the
  // chunk fastest, every wave of CTAs re-streamed both operands: 815 MB of DRAM reads for 537 MB of operands)

  pdl_trigger_entry();
  if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("glowk: dynamic smem not 1024-aligned\n"); __trap(); }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_a); prefetch_tensormap(&tm_b); prefetch_tensormap(&tm_d);
    for (int s = 0; s < num_stages; ++s) { mbar_init(&sh->full_bar[s], 1); mbar_init(&sh->empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh->tmem_full_bar[s], 1); mbar_init(&sh->tmem_empty_bar[s], EPI_WARPS); }
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
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int ki = item / num_tiles, tile = item - ki * num_tiles;
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
      pdl_trigger_drain();
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (warp-uniform control flow keeps descriptors in uniform registers: no R2UR +
    // vote loop around every UTCHMMA); one elected lane issues -- see cnet_fused_sm100.cu
    {
      const uint32_t idesc = make_idesc(block_n, 1, 1);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int ki = item / num_tiles;
        const int kb0 = ki * kb_per_item;
        const int kb1 = min(kb0 + kb_per_item, total_kb);
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
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 16 k-rows = 2048 bytes (>>4 => +128)
              tcgen05_mma_bf16(d_tmem, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc,
                               (kb > kb0) || (k != 0));
            }
            tcgen05_commit(&sh->empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) tcgen05_commit(&sh->tmem_full_bar[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // epilogue: TMEM -> registers -> 128B-swizzled staging box -> TMA reduce-add (fp32) into dW.  The
    // reduction runs in L2 on whole 32x32 boxes instead of one scattered atomic per element.
    const int quarter = warp & 3;
    const int ew = warp - 2, half = ew >> 2;
    uint8_t* sbuf = staging + (size_t)ew * 4096;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item % num_tiles;
      const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
      const int row0 = m_blk * BLOCK_M + quarter * 32;
      mbar_wait(&sh->tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * ACC_STAGE_COLS;
      for (int c0 = half * 32; c0 < block_n; c0 += 64) {
        const int nc = n_blk * block_n + c0;
        if (nc >= No) break;
        uint32_t raw[32];
        tmem_ld32_async(t_row + (uint32_t)c0, raw);
        if (lane == 0) tma_store_wait_read<0>();
        tmem_ld_wait();
        __syncwarp();
        uint8_t* srow = sbuf + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int chunk = j ^ (lane & 7);
          *reinterpret_cast<uint4*>(srow + chunk * 16) = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tm_d, sbuf, nc, row0);     // rows >= Mo / columns >= No are clipped
          tma_store_commit();
        }
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
// Host side
// ---------------------------------------------------------------------------------------------

static int pick_stages(size_t stage_bytes, size_t extra) {
  const size_t budget = 227 * 1024 - extra;
  int s = (int)(budget / stage_bytes);
  if (s > MAX_STAGES) s = MAX_STAGES;
  return s;
}

static size_t gemm_fixed_smem(int epilogue) {
  return STAGING_BYTES + 2 * 256 * sizeof(float) + sizeof(Shared) +
         (epilogue == GLOWK_EPI_RELU_BWD ? STAGING_BYTES + 2 * EPI_NMAX * sizeof(float) : 0);
}

template <int EPI, typename OutT, bool PAIR>
static int launch_gemm_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& ty,
                          int M, int N, int K, int block_n, int stages, int cm, int cn, const EpiParams& ep,
                          cudaStream_t st) {
  const size_t b_rows = PAIR ? block_n / 2 : block_n;
  const size_t smem = (size_t)stages * (BLOCK_M * BLOCK_K * 2 + b_rows * BLOCK_K * 2) + gemm_fixed_smem(EPI);
  auto kern = gemm_tc_kernel<EPI, OutT, PAIR>;
  GLOWK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int csize = cm * cn;
  const int64_t supers = ceil_div(ceil_div(M, BLOCK_M), cm) * ceil_div(ceil_div(N, block_n), cn);
  int clusters = sm_count() / csize;
  if (supers < clusters) clusters = (int)supers;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(clusters * csize), 1, 1);
  cfg.blockDim = dim3(EpiCfg<EPI>::THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (csize > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)csize; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const char* dbg_env = getenv("GLOWK_GEMM_DEBUG");      // profiling aid (1: no loads, 2: no MMA, 4: no epilogue)
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  GLOWK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, ty, M, N, K, block_n, stages, cm, cn, dbg, ep));
  GLOWK_CHECK_LAUNCH("glowk_gemm(tcgen05)");
  return GLOWK_OK;
}

}  // namespace tc

int gemm_debug_trace(unsigned long long* out16) {
  GLOWK_CUDA(cudaDeviceSynchronize());
  GLOWK_CUDA(cudaMemcpyFromSymbol(out16, tc::g_gemm_trace, 16 * sizeof(unsigned long long)));
  return GLOWK_OK;
}

bool tc_available() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
  return major == 10 && tc::encode_fn() != nullptr;
}

// Cluster shape heuristic: share the B (weight) tile among `cm` m-tiles and the A tile among `cn` n-tiles
// whenever the problem has enough tiles; measured on B200 the 128x256 tiles are L2->SM bandwidth bound
// without it.
static void default_cluster(int64_t m_blocks, int64_t n_blocks, int block_n, int* cm, int* cn) {
  *cm = 1; *cn = 1;
  const char* e = getenv("GLOWK_GEMM_CLUSTER");      // "CMxCN" override for experiments (read per call, no state)
  if (e && e[0] >= '1' && e[0] <= '8' && e[1] == 'x' && e[2] >= '1' && e[2] <= '8') { *cm = e[0] - '0'; *cn = e[2] - '0'; }
  if (*cn > n_blocks) *cn = 1;
  while (*cm > 1 && (m_blocks < 2 * *cm * 16 || block_n % (8 * *cm) != 0)) *cm /= 2;
}

int gemm_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                 int epilogue, const EpiParams& ep, void* out, int out_dtype, int64_t ldo, int cm, int cn,
                 cudaStream_t st) {
  using namespace tc;
  GLOWK_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "glowk_gemm(bf16): lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
  GLOWK_CHECK_ARG(((uintptr_t)A | (uintptr_t)B | (uintptr_t)out) % 16 == 0, "glowk_gemm(bf16): operands must be 16-byte aligned");
  GLOWK_CHECK_ARG(M < (1ll << 31) && N < (1 << 20) && K < (1 << 24), "glowk_gemm(bf16): shape out of range");
  const int elo = out_dtype == GLOWK_BF16 ? 2 : 4;
  GLOWK_CHECK_ARG((ldo * elo) % 16 == 0, "glowk_gemm(bf16): output row pitch must be a multiple of 16 bytes");
  if (epilogue == GLOWK_EPI_RELU_BWD) {
    if (out_dtype != GLOWK_BF16) return fail(GLOWK_EUNSUP, "glowk_gemm(bf16): RELU_BWD writes bf16 on the tensor-core path");
    GLOWK_CHECK_ARG(ep.ldy % 8 == 0 && ((uintptr_t)ep.y) % 16 == 0, "glowk_gemm(bf16): y must be 16-byte aligned with ldy %% 8 == 0");
  }
  const int box_cols = 128 / elo;
  const int npad = (int)ceil_div(N, 16) * 16;
  const int n_blocks = (int)ceil_div(npad, 256);
  int block_n = n_blocks == 1 ? npad : (int)ceil_div(ceil_div(npad, n_blocks), box_cols) * box_cols;
  GLOWK_CHECK_ARG(block_n <= 256 && block_n % 16 == 0, "glowk_gemm(bf16): cannot tile N=%lld", (long long)N);
  // CTA pairs (tcgen05 cta_group::2): a 256 x 256 tile per pair of SMs, each CTA holding its 128 rows of A and HALF
  // of the B tile, so the B traffic from L2 (2/3 of a 512 x 512 conv's operand stream, which sits at the L2 cap) is
  // halved and the ring is 6 x 32 KB deep instead of 4 x 48 KB.  Opt-in while it is being validated: GLOWK_GEMM_PAIR=1.
  bool pair = false;
  {
    const char* penv = getenv("GLOWK_GEMM_PAIR");
    const bool epi_ok = (epilogue == GLOWK_EPI_ACTNORM_RELU && out_dtype == GLOWK_BF16) || epilogue == GLOWK_EPI_RELU_BWD;
    if ((cm <= 0 || cn <= 0) && penv && penv[0] == '1' && epi_ok && block_n == 256 && M >= 2 * BLOCK_M) { pair = true; cm = 2; cn = 1; }
  }
  if (cm <= 0 || cn <= 0) default_cluster(ceil_div(M, BLOCK_M), n_blocks, block_n, &cm, &cn);
  GLOWK_CHECK_ARG(cm * cn <= 8 && (cm & (cm - 1)) == 0 && (cn & (cn - 1)) == 0 && block_n % (8 * cm) == 0 && cn <= 16,
                  "glowk_gemm(bf16): bad cluster shape %dx%d for block_n=%d", cm, cn, block_n);
  const size_t stage_bytes = BLOCK_M * BLOCK_K * 2 + (size_t)(pair ? block_n / 2 : block_n) * BLOCK_K * 2;
  const int stages = pick_stages(stage_bytes, gemm_fixed_smem(epilogue));
  GLOWK_CHECK_ARG(stages >= 2, "glowk_gemm(bf16): not enough shared memory for a 2-stage pipeline");

  CUtensorMap ta, tb, to, ty;
  int rc;
  if ((rc = make_map_2d(&ta, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BLOCK_K, (uint32_t)(BLOCK_M / cn)))) return rc;
  if ((rc = make_map_2d(&tb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BLOCK_K, (uint32_t)(block_n / cm)))) return rc;
  if (epilogue == GLOWK_EPI_RELU_BWD) {
    // 16 epilogue warps on [32 rows][32 bf16] boxes (64-byte rows, SWIZZLE_64B) for both the output and y
    if ((rc = make_map_2d(&to, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map_2d(&ty, ep.y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  } else {
    if ((rc = make_map_2d(&to, out, out_dtype == GLOWK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                          elo, (uint64_t)N, (uint64_t)M, (uint64_t)ldo, (uint32_t)box_cols, 32))) return rc;
    ty = to;
  }

#define GLOWK_TC_CASE(E)                                                                                                  \
  case E:                                                                                                                 \
    return out_dtype == GLOWK_BF16                                                                                        \
               ? launch_gemm_tc<E, __nv_bfloat16, false>(ta, tb, to, ty, (int)M, (int)N, (int)K, block_n, stages, cm, cn, ep, st) \
               : launch_gemm_tc<E, float, false>(ta, tb, to, ty, (int)M, (int)N, (int)K, block_n, stages, cm, cn, ep, st);
  if (pair && epilogue == GLOWK_EPI_ACTNORM_RELU && out_dtype == GLOWK_BF16)
    return launch_gemm_tc<GLOWK_EPI_ACTNORM_RELU, __nv_bfloat16, true>(ta, tb, to, ty, (int)M, (int)N, (int)K, block_n, stages, cm, cn, ep, st);
  if (pair && epilogue == GLOWK_EPI_RELU_BWD)
    return launch_gemm_tc<GLOWK_EPI_RELU_BWD, __nv_bfloat16, true>(ta, tb, to, ty, (int)M, (int)N, (int)K, block_n, stages, cm, cn, ep, st);
  switch (epilogue) {
    GLOWK_TC_CASE(GLOWK_EPI_STORE)
    GLOWK_TC_CASE(GLOWK_EPI_ACTNORM_RELU)
    GLOWK_TC_CASE(GLOWK_EPI_ACTNORM)
    GLOWK_TC_CASE(GLOWK_EPI_ZEROS)
    case GLOWK_EPI_RELU_BWD:
      return launch_gemm_tc<GLOWK_EPI_RELU_BWD, __nv_bfloat16, false>(ta, tb, to, ty, (int)M, (int)N, (int)K, block_n, stages, cm, cn, ep, st);
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
  const int stages = pick_stages(stage_bytes, STAGING_BYTES + sizeof(Shared));
  const int tiles = (int)(ceil_div(Mo, BLOCK_M) * n_blocks);
  const int total_kb = (int)ceil_div(P, BLOCK_K);
  int k_items = (int)ceil_div(2 * sm_count(), tiles);
  if (k_items > total_kb) k_items = total_kb;
  int kb_per_item = (int)ceil_div(total_kb, k_items);
  if (kb_per_item < 4) kb_per_item = total_kb < 4 ? total_kb : 4;
  k_items = (int)ceil_div(total_kb, kb_per_item);
  GLOWK_CHECK_ARG(((uintptr_t)dW) % 16 == 0 && lddw % 4 == 0, "glowk_gemm_wgrad(bf16): dW must be 16-byte aligned with lddw %% 4 == 0");
  CUtensorMap ta, tb, td;
  int rc;
  // [P][Mo] row-major: inner dimension = channels (64-wide boxes), outer = pixels (64 per stage)
  if ((rc = make_map_2d(&ta, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)Mo, (uint64_t)P, (uint64_t)lda, 64, BLOCK_K))) return rc;
  if ((rc = make_map_2d(&tb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)No, (uint64_t)P, (uint64_t)ldb, 64, BLOCK_K))) return rc;
  if ((rc = make_map_2d(&td, dW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)No, (uint64_t)Mo, (uint64_t)lddw, 32, 32))) return rc;
  const size_t smem = (size_t)stages * stage_bytes + STAGING_BYTES + sizeof(Shared);
  GLOWK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = tiles * k_items;
  const int grid = items < sm_count() ? items : sm_count();
  GLOWK_CUDA(launch_pdl(wgrad_tc_kernel, grid, GEMM_THREADS, smem, st, ta, tb, td, (int)P, (int)Mo, (int)No, block_n, stages, kb_per_item));
  GLOWK_CHECK_LAUNCH("glowk_gemm_wgrad(tcgen05)");
  return GLOWK_OK;
}

}  // namespace glowk
