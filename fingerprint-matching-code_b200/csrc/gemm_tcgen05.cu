// Dense "NT" GEMM on the 5th-generation tensor cores:  C[M,N] = act(A[M,K] * Bt[N,K]^T + bias).
//
// This is the engine behind the true dense contractions of the matching head - SplineConv's slab GEMM
// ([sum n, 768] x [768, 26*768], ~3 TFLOP per batch of 256 pairs, replacing torch_spline_conv's per-edge
// weighting used by /root/reference/src/model/spline_conv.py:17,35,38) and the AFA-U projections
// (/root/reference/src/model/afau.py:98-102,124-139,189-199).
//
// Design (sm_100a only):
//   * operand tiles [128 x 128 B] (A) and [256 x 128 B] (B) are brought in by TMA (cp.async.bulk.tensor,
//     128-byte swizzle) into a multi-stage shared-memory ring guarded by mbarriers;
//   * one elected thread issues tcgen05.mma.cta_group::1 (M=128, N=256) with the accumulators in tensor
//     memory; four epilogue warps read them back with tcgen05.ld, apply scales / bias / relu, store fp32;
//   * CTAs are rasterised in groups of 32 M-tiles so a wave's operands stay L2-resident.
// Numeric modes (all accumulate in fp32):
//   kTf32x1  operands are the raw fp32 arrays, read as tf32 (1 MMA per k-step; 2^-11 operand error);
//   kTf32x3  operands pre-split into tf32-exact hi + lo; per k-step  lo*hi + hi*lo  go to a correction
//            accumulator (TMEM columns 256..511) and  hi*hi  to the main one: fp32-faithful products.
//            (One accumulator chain of 288 MMAs measured 3e-5 relative error - the fp32 accumulate of the
//            tensor core truncates - hence the separate correction tile, added in the epilogue.)
//   kF16x3   the same error-compensated scheme on fp16 operands: rows are scaled by a power of two into
//            fp16 range, a*s = hi + 2^-11 lo' with hi, lo' fp16 (11-bit mantissas -> 22 bits, exact products
//            in the fp32 accumulator); the epilogue computes (main + 2^-11 corr) / (s_a[m] s_b[n]).  Same
//            accuracy as kTf32x3 at twice the MMA rate and half the operand bytes.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace fpm {

constexpr int TBM = 128, TBN = 256;
constexpr int kRowBytes = 128;                         // one swizzle row = one k-block of a tile
constexpr uint32_t kABytes = TBM * kRowBytes;          // 16 KB
constexpr uint32_t kBBytes = TBN * kRowBytes;          // 32 KB
enum GemmMode { kTf32x1 = 1, kTf32x3 = 3, kF16x3 = 6 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  unsigned long long spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1ull << 26)) __trap();      // a lost arrival becomes an error, not a hang
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Same load, delivered to the same shared-memory offset (and mbarrier offset) of every CTA in cta_mask.
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
  return d;
}
// fp32 accumulate (bits 4-5 = 1), operand format fmt (0 = f16, 2 = tf32) for A (bits 7-9) and B (10-12),
// both K-major, N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Optional per-CTA phase trace (fpm_gemm_set_trace): {smid, t_entry, t_setup_done, t_mainloop_done, t_end} in ns.
__device__ unsigned long long* g_gemm_trace = nullptr;
__device__ int g_gemm_trace_cap = 0;
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

template <bool kF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (kF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// kCluster = 2: two CTAs with adjacent M-tiles and the same N-tile form a cluster; each loads one half of the
// B tile and multicasts it into both CTAs' shared memory, cutting L2->SM operand traffic per CTA from
// 96 KB to 64 KB per k-block (the r1 captures show the kernel bound by the ~7 TB/s L2->SM feed, not by MMA).
template <int kMode, int kStages, int kCluster>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const float* __restrict__ inv_a, const float* __restrict__ inv_b,
               const float* __restrict__ bias, float* __restrict__ Cm, int M, int N, int K, int ldc, int act) {
  constexpr bool kF16 = kMode == kF16x3;
  constexpr bool kTwoAcc = kMode != kTf32x1;
  constexpr int kElemBytes = kF16 ? 2 : 4;
  constexpr int TBK = kRowBytes / kElemBytes;            // 32 tf32 or 64 fp16 elements per k-block
  constexpr int kUmmaKBytes = 32;                        // K = 8 tf32 / 16 fp16 per instruction
  constexpr int kTmemCols = kTwoAcc ? 512 : 256;
  constexpr uint32_t kStageBytes = (kTwoAcc ? 2 : 1) * (kABytes + kBBytes);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = (uint32_t*)(tmem_full_bar + 1);
  float* epi_tiles = (float*)(smem + (size_t)kStages * kStageBytes + 256);   // 8 warps x [32][33] floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tracing = g_gemm_trace != nullptr && (int)blockIdx.x < g_gemm_trace_cap && threadIdx.x == 64;
  unsigned long long tr0 = 0, tr1 = 0, tr2 = 0;
  if (tracing) tr0 = gtime_ns();
  // Grouped rasterisation for L2 reuse: kRasterGroup consecutive M-tiles walk the N-tiles together.  With
  // the plain (n fastest) order the r1 ncu capture showed 13.0 GB of DRAM reads per launch for 0.28 GB of
  // operands: every row of M-tiles swept all of B (123 MB ~ the whole L2).
  constexpr int kRasterGroup = 32;
  const int tiles_m = ((M + TBM - 1) / TBM + kCluster - 1) / kCluster * kCluster;   // padded to the cluster size
  const int tiles_n = (N + TBN - 1) / TBN;
  const int per_group = kRasterGroup * tiles_n;
  const int grp = (int)blockIdx.x / per_group, rem = (int)blockIdx.x - grp * per_group;
  const int gsize = min(kRasterGroup, tiles_m - grp * kRasterGroup);
  const int m0 = (grp * kRasterGroup + rem % gsize) * TBM, n0 = (rem / gsize) * TBN;
  const int nk = (K + TBK - 1) / TBK;

  const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);
  if (threadIdx.x == 0) {
    // a slot is free again when the MMA warps of ALL CTAs that receive multicast data into it released it
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kCluster); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();      // peers must see initialised barriers before signalling them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* st = smem + (size_t)s * kStageBytes;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        const int kc = kb * TBK;
        tma_load_2d(&tmA_hi, &full_bar[s], st, kc, m0);
        if (kTwoAcc) tma_load_2d(&tmA_lo, &full_bar[s], st + kABytes + kBBytes, kc, m0);
        if (kCluster == 1) {
          tma_load_2d(&tmB_hi, &full_bar[s], st + kABytes, kc, n0);
          if (kTwoAcc) tma_load_2d(&tmB_lo, &full_bar[s], st + 2 * kABytes + kBBytes, kc, n0);
        } else {
          // this CTA's share of the B tile (TBN / kCluster rows), delivered to every CTA of the cluster
          constexpr uint32_t kShare = kBBytes / kCluster;
          const int nrow = n0 + (int)cta_rank * (TBN / kCluster);
          tma_load_2d_mc(&tmB_hi, &full_bar[s], st + kABytes + cta_rank * kShare, kc, nrow, kMask);
          if (kTwoAcc)
            tma_load_2d_mc(&tmB_lo, &full_bar[s], st + 2 * kABytes + kBBytes + cta_rank * kShare, kc, nrow, kMask);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TBM, TBN, kF16 ? 0u : 2u);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
        const uint64_t a_hi = make_smem_desc(st), b_hi = make_smem_desc(st + kABytes);
        if (kTwoAcc) {
          const uint64_t a_lo = make_smem_desc(st + kABytes + kBBytes);
          const uint64_t b_lo = make_smem_desc(st + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < kRowBytes / kUmmaKBytes; ++k) {
            const uint64_t adv = (uint64_t)((k * kUmmaKBytes) >> 4);
            umma<kF16>(tmem_base + TBN, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma<kF16>(tmem_base + TBN, a_hi + adv, b_lo + adv, idesc, 1u);
            umma<kF16>(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
          }
        } else {
#pragma unroll
          for (int k = 0; k < kRowBytes / kUmmaKBytes; ++k) {
            const uint64_t adv = (uint64_t)((k * kUmmaKBytes) >> 4);
            umma<kF16>(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
          }
        }
        // frees the smem slot (in every CTA that multicasts into it) once these MMAs have read it
        if (kCluster == 1) umma_commit(&empty_bar[s]); else umma_commit_mc(&empty_bar[s], kMask);
      }
      umma_commit(tmem_full_bar);             // accumulators complete
    }
  } else {
    // ===== epilogue: warps 2..9.  TMEM lane quarter = warp % 4; the two warps of a quarter split the columns.
    // Per 32-column chunk: tcgen05.ld (main + correction) -> scales / bias / relu in registers -> transpose
    // through a padded shared tile -> 128-byte coalesced row stores.  (The first version stored 16 bytes per
    // thread per row: 32 cache lines per store instruction, and it re-loaded the column scale per element; with
    // one warp per scheduler nothing hid those latencies and the epilogue cost as much as the main loop.)
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* tile = epi_tiles + (size_t)(warp - 2) * (32 * 33);
    if (tracing) tr1 = gtime_ns();
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (tracing) tr2 = gtime_ns();
    const int m = m0 + q * 32 + lane;
    const float row_scale = (kF16 && m < M) ? inv_a[m] : 1.f;
    constexpr float kCorrScale = kF16 ? (1.0f / 2048.0f) : 1.0f;
    constexpr int kChunksPerWarp = TBN / 32 / 2;
#pragma unroll 1
    for (int ch = half * kChunksPerWarp; ch < (half + 1) * kChunksPerWarp; ++ch) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
      const int nb = n0 + ch * 32;
      // this lane's column constants, broadcast below with shuffles
      const int ncol = nb + lane;
      const float col_scale = (kF16 && ncol < N) ? inv_b[ncol] : 1.f;
      const float col_bias = (bias && ncol < N) ? bias[ncol] : 0.f;
      if (kTwoAcc) {
        uint32_t r2[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(TBN + ch * 32), r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(fmaf(__uint_as_float(r2[j]), kCorrScale, __uint_as_float(r[j])));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]);
        if (kF16) x = x * row_scale * __shfl_sync(0xffffffffu, col_scale, j);     // exact: powers of two
        x += __shfl_sync(0xffffffffu, col_bias, j);
        if (act == 1) x = fmaxf(x, 0.f);
        tile[lane * 33 + j] = x;                     // row = this thread's accumulator row, conflict-free
      }
      __syncwarp();
      if (ncol < N) {
        const int mrow0 = m0 + q * 32;
        const int nrows = min(32, M - mrow0);
        float* cbase = Cm + (size_t)mrow0 * ldc + ncol;
        if (nrows == 32) {
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += 8) {           // 8 independent LDS in flight, then 8 row stores
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = tile[(r0 + u) * 33 + lane];
#pragma unroll
            for (int u = 0; u < 8; ++u) cbase[(size_t)(r0 + u) * ldc] = v[u];
          }
        } else {
          for (int rr = 0; rr < nrows; ++rr) cbase[(size_t)rr * ldc] = tile[rr * 33 + lane];
        }
      }
      __syncwarp();
    }
    if (tracing) {
      unsigned long long* t = g_gemm_trace + (size_t)blockIdx.x * 5;
      t[0] = smid(); t[1] = tr0; t[2] = tr1; t[3] = tr2; t[4] = gtime_ns();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();      // no CTA may retire while a peer can still write into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// Persistent CTA-pair variant (the default for the error-compensated modes).
//
// Why: the r1b captures show the one-tile-per-CTA kernel above keeps the tensor pipe busy 48.6 % of the
// time - its main loop runs at ~86 % (shared-memory bandwidth: the three products of a k-step re-read their
// operands), but the epilogue (32 % of a CTA's life) and the CTA turn-over (7 %) are serial, because the two
// accumulators of a 128x256 tile fill all 512 TMEM columns and nothing can be computed while they drain.
// Here two CTAs of a cluster issue tcgen05.mma.cta_group::2 on a 256 x 128 tile: each CTA holds 128 rows x 128
// columns (main + correction = 256 TMEM columns), so TMEM has room for TWO tiles and the epilogue of tile i
// overlaps the main loop of tile i+1.  Each CTA loads its own A rows and one half of the B tile (the pair's
// tensor cores share both halves), every TMA load signals the LEADER's full barrier, the leader alone issues
// the MMAs, and tcgen05.commit multicasts slot-release / accumulator-ready to both CTAs.  The kernel is
// persistent: one cluster per SM pair walks the tile list with the same L2-friendly rasterisation as above.
// ---------------------------------------------------------------------------------------------------
constexpr int P_TBM = 128;                    // rows per CTA (256 per pair)
constexpr int P_TBN = 128;                    // columns per pair tile
constexpr uint32_t kPABytes = P_TBM * kRowBytes;              // 16 KB
constexpr uint32_t kPBHalfBytes = (P_TBN / 2) * kRowBytes;    //  8 KB: this CTA's half of the B tile
constexpr uint32_t kPStageBytes = 2 * (kPABytes + kPBHalfBytes);   // hi + lo: 48 KB
constexpr int kPStagesMax = 4;            // pipeline depth: template parameter kPStages of the pair kernel (3 or 4)
constexpr int kEpiWarps = 8;

__device__ __forceinline__ uint32_t mapa_rank0(uint32_t saddr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA-pair kernel: data lands in THIS CTA's shared memory, the bytes are counted on the mbarrier at
// `bar_cluster_addr` (the leader CTA's barrier, a shared::cluster address).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <bool kF16>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (kF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// Optional tile table (device memory, built without a host round trip by the SplineConv planner in spline.cu):
// entry t = {first A row of the 256-row pair tile, first B row of the 128-row tile, first C column, rowmap offset};
// rowmap offset < 0: C rows = A rows; otherwise C row of A row a0 + r is rowmap[offset + r] (-1 = padding, skipped).
// *tab_count entries are valid.  Lets one launch compute only the (row block, weight slab) products that are used.
struct PairTile { int a_row0, b_row0, c_col0, rowmap_off; };

// kPStages = 4 fills the SM's shared memory (226 KB); 3 leaves 46 KB so that light CTAs of ANOTHER kernel (the
// SplineConv gather of the other image, launched on a second stream) can run beside the GEMM CTA.
// __maxnreg__(128): the register file is split over the SM's four sub-partitions (16 K registers each) and a CTA's
// warps are dealt to them round-robin; at the 160 registers ptxas takes when left alone the fullest sub-partition
// (3 of the 10 warps) keeps 1 K registers free and no warp of any other kernel fits beside the GEMM CTA.  At 128
// (no spills) it keeps 4 K: two warps of the 56-register gather kernel.
template <int kMode, int kPStages, int kRegs>
__global__ void __maxnreg__(kRegs)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const float* __restrict__ inv_a, const float* __restrict__ inv_b,
                    const float* __restrict__ bias, float* __restrict__ Cm, int M, int N, int K, int ldc, int act,
                    const PairTile* __restrict__ tab, const int* __restrict__ tab_count,
                    const int* __restrict__ rowmap, int m_ident) {
  static_assert(kMode == kTf32x3 || kMode == kF16x3, "the pair kernel serves the two-accumulator modes");
  constexpr bool kF16 = kMode == kF16x3;
  constexpr int kElemBytes = kF16 ? 2 : 4;
  constexpr int TBK = kRowBytes / kElemBytes;
  constexpr int kUmmaKBytes = 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)kPStages * kPStageBytes);
  uint64_t* empty_bar = full_bar + kPStages;
  uint64_t* tmem_full_bar = empty_bar + kPStages;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2] (the leader's copy is the one in use)
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty_bar + 2);
  // kEpiWarps x [32][32] floats, element (row, col) at row * 32 + (col ^ row): conflict-free for the row-wise writes and
  // the column-wise reads below without the 33-float padding (1 KB less shared memory per CTA: together with the
  // 16-edge chunks of spline_gather_max_kernel that lets one gather CTA run beside the GEMM CTA of the OTHER graph)
  float* epi_tiles = (float*)(smem + (size_t)kPStages * kPStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster_id = (int)blockIdx.x >> 1, num_clusters = (int)gridDim.x >> 1;

  constexpr int kRasterGroup = 16;                       // 16 pair tiles = 4096 rows walk the N tiles together
  const int tiles_m = (M + 2 * P_TBM - 1) / (2 * P_TBM);
  const int tiles_n = (N + P_TBN - 1) / P_TBN;
  const long long total_tiles = tab ? (long long)*tab_count : (long long)tiles_m * tiles_n;
  const int per_group = kRasterGroup * tiles_n;
  const int nk = (K + TBK - 1) / TBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // m0: first A row, n0: first B row, c0: first C column, rm: rowmap offset (-1 = identity rows)
  auto tile_origin = [&](long long tile, int& m0, int& n0, int& c0, int& rm) {
    if (tab) {
      const PairTile t = tab[tile];
      m0 = t.a_row0; n0 = t.b_row0; c0 = t.c_col0; rm = t.rowmap_off;
      return;
    }
    const int grp = (int)(tile / per_group), rem = (int)(tile - (long long)grp * per_group);
    const int gsize = min(kRasterGroup, tiles_m - grp * kRasterGroup);
    m0 = (grp * kRasterGroup + rem % gsize) * (2 * P_TBM);
    n0 = (rem / gsize) * P_TBN;
    c0 = n0; rm = -1;
  };

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows + own half of the B tile, counted on the leader's barrier =====
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        int m0, n0, c0, rm;
        tile_origin(tile, m0, n0, c0, rm);
        const int am = m0 + (int)cta_rank * P_TBM, bn = n0 + (int)cta_rank * (P_TBN / 2);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % kPStages;
          const uint32_t ph = (it / kPStages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (cta_rank == 0) mbar_expect_tx(&full_bar[s], 2 * kPStageBytes);
          const uint32_t bar = mapa_rank0(smem_u32(&full_bar[s]));
          uint8_t* st = smem + (size_t)s * kPStageBytes;
          const int kc = kb * TBK;
          tma_load_2d_pair(&tmA_hi, bar, st, kc, am);
          tma_load_2d_pair(&tmA_lo, bar, st + kPABytes, kc, am);
          tma_load_2d_pair(&tmB_hi, bar, st + 2 * kPABytes, kc, bn);
          tma_load_2d_pair(&tmB_lo, bar, st + 2 * kPABytes + kPBHalfBytes, kc, bn);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (cta_rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * P_TBM, P_TBN, kF16 ? 0u : 2u);
      uint32_t it = 0, t = 0;
      for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters, ++t) {
        const uint32_t buf = t & 1u, bph = (t >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[buf], bph ^ 1u);        // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        const uint32_t acc_main = tmem_base + buf * (2 * P_TBN), acc_corr = acc_main + P_TBN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % kPStages;
          const uint32_t ph = (it / kPStages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)s * kPStageBytes);
          const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + kPABytes);
          const uint64_t b_hi = make_smem_desc(st + 2 * kPABytes);
          const uint64_t b_lo = make_smem_desc(st + 2 * kPABytes + kPBHalfBytes);
#pragma unroll
          for (int k = 0; k < kRowBytes / kUmmaKBytes; ++k) {
            const uint64_t adv = (uint64_t)((k * kUmmaKBytes) >> 4);
            umma_pair<kF16>(acc_corr, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma_pair<kF16>(acc_corr, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_pair<kF16>(acc_main, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
          }
          umma_commit_pair(&empty_bar[s]);                // frees slot s in both CTAs
        }
        umma_commit_pair(&tmem_full_bar[buf]);            // accumulators of this tile complete, in both CTAs
      }
    }
  } else {
    // ===== epilogue (both CTAs): warps 2..9, TMEM lane quarter = warp % 4, two warps per quarter split the columns
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* tile_s = epi_tiles + (size_t)(warp - 2) * (32 * 32);
    const uint32_t empty_addr0 = mapa_rank0(smem_u32(&tmem_empty_bar[0]));
    const uint32_t empty_addr1 = mapa_rank0(smem_u32(&tmem_empty_bar[1]));
    constexpr float kCorrScale = kF16 ? (1.0f / 2048.0f) : 1.0f;
    constexpr int kChunksPerWarp = P_TBN / 32 / 2;          // 2
    uint32_t t = 0;
    for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters, ++t) {
      int m0, n0, c0, rm;
      tile_origin(tile, m0, n0, c0, rm);
      m0 += (int)cta_rank * P_TBM;
      if (rm >= 0) rm += (int)cta_rank * P_TBM;
      const uint32_t buf = t & 1u, bph = (t >> 1) & 1u;
      mbar_wait(&tmem_full_bar[buf], bph);
      tc_fence_after();
      const uint32_t acc_main = tmem_base + buf * (2 * P_TBN) + ((uint32_t)(q * 32) << 16);
      // all of this warp's accumulator columns -> registers, then hand the TMEM buffer back before storing
      uint32_t r[kChunksPerWarp][32];
#pragma unroll
      for (int c = 0; c < kChunksPerWarp; ++c) {
        const int ch = half * kChunksPerWarp + c;
        uint32_t r2[32];
        tmem_ld32(acc_main + (uint32_t)(ch * 32), r[c]);
        tmem_ld32(acc_main + (uint32_t)(P_TBN + ch * 32), r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[c][j] = __float_as_uint(fmaf(__uint_as_float(r2[j]), kCorrScale, __uint_as_float(r[c][j])));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0);

      const int m = m0 + q * 32 + lane;
      const float row_scale = (kF16 && m < M) ? inv_a[m] : 1.f;
      // destination row of this lane's accumulator row (identity, or through the planner's row map)
      // (m_ident: rows of C that identity tiles may write - the A buffer of a planned launch is longer than C)
      const int crow_lane = rm < 0 ? (m < m_ident ? m : -1) : (m < M ? rowmap[rm + q * 32 + lane] : -1);
#pragma unroll
      for (int c = 0; c < kChunksPerWarp; ++c) {
        const int ch = half * kChunksPerWarp + c;
        const int brow = n0 + ch * 32 + lane;                 // row of Bt = output feature
        const int ncol = c0 + ch * 32 + lane;                 // column of C
        const float col_scale = (kF16 && brow < N) ? inv_b[brow] : 1.f;
        const float col_bias = (bias && brow < N) ? bias[brow] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[c][j]);
          if (kF16) x = x * row_scale * __shfl_sync(0xffffffffu, col_scale, j);
          x += __shfl_sync(0xffffffffu, col_bias, j);
          if (act == 1) x = fmaxf(x, 0.f);
          tile_s[lane * 32 + (j ^ lane)] = x;
        }
        __syncwarp();
        if (rm >= 0) {
          // mapped rows (gathered A): every row has its own destination
          for (int rr = 0; rr < 32; ++rr) {
            const int crow = __shfl_sync(0xffffffffu, crow_lane, rr);
            if (crow >= 0 && brow < N) Cm[(size_t)crow * ldc + ncol] = tile_s[rr * 32 + (lane ^ rr)];
          }
        } else if (brow < N) {
          const int mrow0 = m0 + q * 32;
          const int nrows = min(32, m_ident - mrow0);
          float* cbase = Cm + (size_t)mrow0 * ldc + ncol;
          if (nrows == 32) {
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
              float v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = tile_s[(r0 + u) * 32 + (lane ^ (r0 + u))];
#pragma unroll
              for (int u = 0; u < 8; ++u) cbase[(size_t)(r0 + u) * ldc] = v[u];
            }
          } else {
            for (int rr = 0; rr < nrows; ++rr) cbase[(size_t)rr * ldc] = tile_s[rr * 32 + (lane ^ rr)];
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// a = hi + lo with hi, lo exactly representable in tf32 (round-to-nearest split).
__global__ void tf32_split_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                  size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = ((const float4*)src)[i];
  float4 h, l;
  auto split = [](float x, float& hh, float& ll) {
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
    hh = __uint_as_float(hb);
    const float rem = x - hh;                 // exact
    uint32_t lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
    ll = __uint_as_float(lb);
  };
  split(a.x, h.x, l.x); split(a.y, h.y, l.y); split(a.z, h.z, l.z); split(a.w, h.w, l.w);
  ((float4*)hi)[i] = h;
  ((float4*)lo)[i] = l;
}

// One warp per row: s = 2^-e with amax * s in [0.5, 1);  a*s = hi + 2^-11 * lo,  hi/lo fp16;  inv[row] = 2^e.
__global__ void __launch_bounds__(256)
f16_split_rows_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                      float* __restrict__ inv_scale, int rows, int K) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* s4 = (const float4*)(src + (size_t)row * K);
  const int n4 = K >> 2;
  float amax = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = s4[i];
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  amax = warp_max(amax);
  int e = 0;
  if (amax > 0.f && amax < INFINITY) frexpf(amax, &e);
  e = max(-100, min(100, e));
  const float s = ldexpf(1.f, -e);
  if (lane == 0) inv_scale[row] = ldexpf(1.f, e);
  __half2* h2 = (__half2*)(hi + (size_t)row * K);
  __half2* l2 = (__half2*)(lo + (size_t)row * K);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = s4[i];
    const float x[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
    __half hh[4], ll[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hh[j] = __float2half_rn(x[j]);
      ll[j] = __float2half_rn((x[j] - __half2float(hh[j])) * 2048.f);
    }
    h2[2 * i] = __halves2half2(hh[0], hh[1]); h2[2 * i + 1] = __halves2half2(hh[2], hh[3]);
    l2[2 * i] = __halves2half2(ll[0], ll[1]); l2[2 * i + 1] = __halves2half2(ll[2], ll[3]);
  }
}

}  // namespace fpm

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-d map over a row-major [rows, K] matrix with leading dimension ld (elements); box = [box_rows x 128 bytes].
static int make_map(CUtensorMap* map, const void* base, int rows, int K, int ld, int box_rows, bool f16) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { fpm_set_error("cuTensorMapEncodeTiled unavailable"); return FPM_ERR_UNSUPPORTED; }
  const int eb = f16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * eb};
  cuuint32_t box[2] = {(cuuint32_t)(fpm::kRowBytes / eb), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base,
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fpm_set_error("cuTensorMapEncodeTiled failed"); return FPM_ERR_ARG; }
  return FPM_OK;
}

static int g_tc_cluster = -1;     // FPMATCH_GEMM_CLUSTER: 1 = no cluster, 2 = B-tile multicast pairs (default)

template <int kMode, int kStages, int kCluster>
static int launch_tc_impl(const void* A_hi, const void* A_lo, const void* B_hi, const void* B_lo, const float* inv_a,
                          const float* inv_b, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
                          int ldc, int act, cudaStream_t st) {
  constexpr bool f16 = kMode == fpm::kF16x3;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, A_hi, M, K, lda, fpm::TBM, f16)) != FPM_OK) return rc;
  if ((rc = make_map(&mAl, A_lo, M, K, lda, fpm::TBM, f16)) != FPM_OK) return rc;
  if ((rc = make_map(&mBh, B_hi, N, K, ldb, fpm::TBN / kCluster, f16)) != FPM_OK) return rc;
  if ((rc = make_map(&mBl, B_lo, N, K, ldb, fpm::TBN / kCluster, f16)) != FPM_OK) return rc;
  const long long tiles_m = (fpm_cdiv(M, fpm::TBM) + kCluster - 1) / kCluster * kCluster;
  const long long tiles = (long long)fpm_cdiv(N, fpm::TBN) * tiles_m;
  FPM_CHECK_ARG(tiles <= 0x7fffffffLL, "gemm_tc: too many tiles");
  const size_t smem = (size_t)kStages * (kMode == fpm::kTf32x1 ? 1 : 2) * (fpm::kABytes + fpm::kBBytes) + 1024 + 256 + 8 * 32 * 33 * 4;
  auto kern = fpm::gemm_tc_kernel<kMode, kStages, kCluster>;
  FPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kCluster > 1 ? 1 : 0;
  FPM_CUDA(cudaLaunchKernelEx(&cfg, kern, mAh, mAl, mBh, mBl, inv_a, inv_b, bias, C, M, N, K, ldc, act));
  return FPM_OK;
}

static int g_tc_pair = -1;        // FPMATCH_GEMM_PAIR / fpm_gemm_set_pair: 1 = persistent CTA-pair kernel (default), 0 = off
static int g_pair_cluster_cap = 0;        // fpm_gemm_set_max_clusters: 0 = every SM pair
static int g_pair_clusters[2] = {0, 0};   // co-resident clusters of the pair kernel per mode (occupancy query, cached)

template <int kMode>
static int launch_tc_pair(const void* A_hi, const void* A_lo, const void* B_hi, const void* B_lo, const float* inv_a,
                          const float* inv_b, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
                          int ldc, int act, cudaStream_t st, const fpm::PairTile* tab = nullptr,
                          const int* tab_count = nullptr, const int* rowmap = nullptr, long long max_tiles = 0,
                          int m_ident = -1) {
  constexpr bool f16 = kMode == fpm::kF16x3;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, A_hi, M, K, lda, fpm::P_TBM, f16)) != FPM_OK) return rc;
  if ((rc = make_map(&mAl, A_lo, M, K, lda, fpm::P_TBM, f16)) != FPM_OK) return rc;
  if ((rc = make_map(&mBh, B_hi, N, K, ldb, fpm::P_TBN / 2, f16)) != FPM_OK) return rc;
  if ((rc = make_map(&mBl, B_lo, N, K, ldb, fpm::P_TBN / 2, f16)) != FPM_OK) return rc;
  // FPMATCH_GEMM_STAGES=2|3|4.  Default 3: measured as fast as 4 on the slab GEMM (0.751 vs 0.757 ms per launch, r2d) and
  // it leaves 46 KB of shared memory to CTAs of other streams' kernels.
  static int stages = 0;
  if (stages == 0) {
    const char* e = getenv("FPMATCH_GEMM_STAGES");
    stages = (e && e[0] == '4') ? 4 : (e && e[0] == '2') ? 2 : 3;
  }
  const size_t smem = (size_t)stages * fpm::kPStageBytes + 1024 + 256 + (size_t)fpm::kEpiWarps * 32 * 32 * 4;
  // FPMATCH_GEMM_REGS=96: the 3-stage kernel capped at 96 registers per thread (104 bytes of spill in the epilogue)
  // so that more of the other streams' CTAs fit beside it.
  static int regs = 0;
  if (regs == 0) {
    const char* e = getenv("FPMATCH_GEMM_REGS");
    regs = (e && atoi(e) == 96) ? 96 : 128;
  }
  auto kern = stages == 3 ? (regs == 96 ? fpm::gemm_tc_pair_kernel<kMode, 3, 96> : fpm::gemm_tc_pair_kernel<kMode, 3, 128>)
                          : stages == 2 ? fpm::gemm_tc_pair_kernel<kMode, 2, 128>
                                        : fpm::gemm_tc_pair_kernel<kMode, 4, 128>;
  FPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(64 + 32 * fpm::kEpiWarps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int& ncl = g_pair_clusters[f16 ? 1 : 0];
  if (ncl == 0) {
    // a persistent kernel must not launch more clusters than can be resident at once: a late cluster would run
    // its whole tile list after the others have finished
    int dev = 0, sms = 0, maxc = 0;
    FPM_CUDA(cudaGetDevice(&dev));
    FPM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cfg.gridDim = dim3((unsigned)(sms / 2 * 2));
    if (cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg) != cudaSuccess || maxc <= 0) {
      (void)cudaGetLastError();
      maxc = sms / 2;
    }
    ncl = maxc < sms / 2 ? maxc : sms / 2;
    const char* e = getenv("FPMATCH_GEMM_CLUSTERS");   // experiment: leave SMs free for other streams' kernels
    if (e && atoi(e) > 0 && atoi(e) < ncl) ncl = atoi(e);
  }
  const long long tiles = tab ? max_tiles : (long long)fpm_cdiv(M, 2 * fpm::P_TBM) * fpm_cdiv(N, fpm::P_TBN);
  const int ncl_eff = (g_pair_cluster_cap > 0 && g_pair_cluster_cap < ncl) ? g_pair_cluster_cap : ncl;
  const long long clusters = tiles < ncl_eff ? (tiles > 0 ? tiles : 1) : ncl_eff;
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  FPM_CUDA(cudaLaunchKernelEx(&cfg, kern, mAh, mAl, mBh, mBl, inv_a, inv_b, bias, C, M, N, K, ldc, act, tab, tab_count,
                              rowmap, m_ident < 0 ? M : m_ident));
  return FPM_OK;
}

template <int kMode, int kStages>
static int launch_tc(const void* A_hi, const void* A_lo, const void* B_hi, const void* B_lo, const float* inv_a,
                     const float* inv_b, const float* bias, float* C, int M, int N, int K, int lda, int ldb,
                     int ldc, int act, cudaStream_t st) {
  if (g_tc_pair < 0) {
    const char* e = getenv("FPMATCH_GEMM_PAIR");
    g_tc_pair = (e && e[0] == '0') ? 0 : 1;
  }
  if constexpr (kMode != fpm::kTf32x1) {
    if (g_tc_pair == 1)
      return launch_tc_pair<kMode>(A_hi, A_lo, B_hi, B_lo, inv_a, inv_b, bias, C, M, N, K, lda, ldb, ldc, act, st);
  }
  if (g_tc_cluster < 0) {
    const char* e = getenv("FPMATCH_GEMM_CLUSTER");
    g_tc_cluster = (e && e[0] == '1') ? 1 : 2;
  }
  if (g_tc_cluster == 2)
    return launch_tc_impl<kMode, kStages, 2>(A_hi, A_lo, B_hi, B_lo, inv_a, inv_b, bias, C, M, N, K, lda, ldb, ldc, act, st);
  return launch_tc_impl<kMode, kStages, 1>(A_hi, A_lo, B_hi, B_lo, inv_a, inv_b, bias, C, M, N, K, lda, ldb, ldc, act, st);
}

// Debug aid: point the kernels at a device buffer of 5 * cap uint64 to record per-CTA phase timestamps
// (NULL switches tracing off).  Used by tools/gemm_phase_trace.py; not part of the product path.
extern "C" int fpm_gemm_set_trace(void* buf, int cap) {
  unsigned long long* p = (unsigned long long*)buf;
  FPM_CUDA(cudaMemcpyToSymbol(fpm::g_gemm_trace, &p, sizeof(p)));
  FPM_CUDA(cudaMemcpyToSymbol(fpm::g_gemm_trace_cap, &cap, sizeof(cap)));
  return FPM_OK;
}

// Upper bound on the clusters (SM pairs) of the persistent pair kernel; 0 = all.  With several batches in flight on
// different streams a few SMs left free of GEMM CTAs let the other batch's shared-memory-heavy tail kernels (Sinkhorn,
// the association-graph layers) run during the GEMM phases: 74 -> 70 clusters took 7.77 -> 7.59 ms per step (r2n).
extern "C" int fpm_gemm_set_max_clusters(int clusters) {
  FPM_CHECK_ARG(clusters >= 0, "fpm_gemm_set_max_clusters: negative");
  g_pair_cluster_cap = clusters;
  return FPM_OK;
}

// 1: the error-compensated modes run the persistent CTA-pair kernel (default); 0: the one-tile-per-CTA kernel.
extern "C" int fpm_gemm_set_pair(int on) {
  g_tc_pair = on ? 1 : 0;
  return FPM_OK;
}

extern "C" int fpm_tf32_split(const float* src, float* hi, float* lo, long long n, void* stream) {
  FPM_CHECK_ARG(src && hi && lo, "fpm_tf32_split: null tensor");
  FPM_CHECK_ARG(n >= 0 && (n & 3) == 0, "fpm_tf32_split: element count must be a multiple of 4");
  FPM_CHECK_ARG(((((size_t)src) | ((size_t)hi) | ((size_t)lo)) & 15) == 0, "fpm_tf32_split: 16-byte alignment required");
  if (n == 0) return FPM_OK;
  const size_t n4 = (size_t)n / 4;
  fpm::tf32_split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, hi, lo, n4);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_f16_split_rows(const float* src, void* hi, void* lo, float* inv_scale, int rows, int K,
                                  void* stream) {
  FPM_CHECK_ARG(src && hi && lo && inv_scale, "fpm_f16_split_rows: null tensor");
  FPM_CHECK_ARG(rows >= 0 && K > 0 && (K & 7) == 0, "fpm_f16_split_rows: K must be a multiple of 8");
  FPM_CHECK_ARG(((((size_t)src) | ((size_t)hi) | ((size_t)lo)) & 15) == 0, "fpm_f16_split_rows: 16-byte alignment required");
  if (rows == 0) return FPM_OK;
  fpm::f16_split_rows_kernel<<<fpm_cdiv((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      src, (__half*)hi, (__half*)lo, inv_scale, rows, K);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// passes = 1: A_hi / Bt_hi are the raw fp32 operands (the tensor core reads them as tf32), *_lo unused.
// passes = 3: (A_hi, A_lo), (Bt_hi, Bt_lo) are tf32-exact splits from fpm_tf32_split.
extern "C" int fpm_gemm_nt_tc(const float* A_hi, const float* A_lo, const float* Bt_hi, const float* Bt_lo,
                              const float* bias, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                              int act, int passes, void* stream) {
  FPM_CHECK_ARG(A_hi && Bt_hi && C, "fpm_gemm_nt_tc: null tensor");
  FPM_CHECK_ARG(M >= 0 && N > 0 && K > 0, "fpm_gemm_nt_tc: bad sizes");
  FPM_CHECK_ARG(passes == 1 || passes == 3, "fpm_gemm_nt_tc: passes must be 1 or 3");
  FPM_CHECK_ARG(passes == 1 || (A_lo && Bt_lo), "fpm_gemm_nt_tc: 3-pass mode needs the lo parts");
  FPM_CHECK_ARG(act == 0 || act == 1, "fpm_gemm_nt_tc: unknown activation");
  FPM_CHECK_ARG((K & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0, "fpm_gemm_nt_tc: K, lda, ldb must be multiples of 4");
  FPM_CHECK_ARG((((size_t)A_hi) & 15) == 0 && (((size_t)Bt_hi) & 15) == 0, "fpm_gemm_nt_tc: operands must be 16-byte aligned");
  FPM_CHECK_ARG(passes == 1 || (((((size_t)A_lo) | ((size_t)Bt_lo)) & 15) == 0), "fpm_gemm_nt_tc: operands must be 16-byte aligned");
  if (M == 0) return FPM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (passes == 1)
    return launch_tc<fpm::kTf32x1, 4>(A_hi, A_hi, Bt_hi, Bt_hi, nullptr, nullptr, bias, C, M, N, K, lda, ldb, ldc, act, st);
  return launch_tc<fpm::kTf32x3, 2>(A_hi, A_lo, Bt_hi, Bt_lo, nullptr, nullptr, bias, C, M, N, K, lda, ldb, ldc, act, st);
}

// Error-compensated fp16 mode: operands from fpm_f16_split_rows (fp16 hi / lo + per-row inverse scales).
extern "C" int fpm_gemm_nt_f16x3(const void* A_hi, const void* A_lo, const float* inv_a, const void* Bt_hi,
                                 const void* Bt_lo, const float* inv_b, const float* bias, float* C, int M, int N,
                                 int K, int lda, int ldb, int ldc, int act, void* stream) {
  FPM_CHECK_ARG(A_hi && A_lo && inv_a && Bt_hi && Bt_lo && inv_b && C, "fpm_gemm_nt_f16x3: null tensor");
  FPM_CHECK_ARG(M >= 0 && N > 0 && K > 0, "fpm_gemm_nt_f16x3: bad sizes");
  FPM_CHECK_ARG(act == 0 || act == 1, "fpm_gemm_nt_f16x3: unknown activation");
  FPM_CHECK_ARG((K & 7) == 0 && (lda & 7) == 0 && (ldb & 7) == 0, "fpm_gemm_nt_f16x3: K, lda, ldb must be multiples of 8");
  FPM_CHECK_ARG(((((size_t)A_hi) | ((size_t)A_lo) | ((size_t)Bt_hi) | ((size_t)Bt_lo)) & 15) == 0,
                "fpm_gemm_nt_f16x3: operands must be 16-byte aligned");
  if (M == 0) return FPM_OK;
  return launch_tc<fpm::kF16x3, 2>(A_hi, A_lo, Bt_hi, Bt_lo, inv_a, inv_b, bias, C, M, N, K, lda, ldb, ldc, act,
                                   (cudaStream_t)stream);
}

// Tile-table form of fpm_gemm_nt_f16x3 (persistent CTA-pair kernel only): computes C blocks for the *tab_count tiles
// listed in `tab` (device memory, 4 ints per tile: first A row of a 256-row block, first Bt row of a 128-row block,
// first C column, rowmap offset or -1); see PairTile above.  max_tiles bounds the grid (host-side upper bound of
// *tab_count); M = rows of the A buffer, m_ident = rows of C that identity-mapped tiles may write.  Used by the SplineConv forward to skip the (node block, weight slab) products no edge refers to.
extern "C" int fpm_gemm_nt_f16x3_tiles(const void* A_hi, const void* A_lo, const float* inv_a, const void* Bt_hi,
                                       const void* Bt_lo, const float* inv_b, float* C, int M, int N, int K, int lda,
                                       int ldb, int ldc, const int* tab, const int* tab_count, const int* rowmap,
                                       long long max_tiles, int m_ident, void* stream) {
  FPM_CHECK_ARG(A_hi && A_lo && inv_a && Bt_hi && Bt_lo && inv_b && C && tab && tab_count,
                "fpm_gemm_nt_f16x3_tiles: null tensor");
  FPM_CHECK_ARG(M >= 0 && N > 0 && K > 0 && max_tiles >= 0 && m_ident >= 0 && m_ident <= M,
                "fpm_gemm_nt_f16x3_tiles: bad sizes");
  FPM_CHECK_ARG((K & 7) == 0 && (lda & 7) == 0 && (ldb & 7) == 0, "fpm_gemm_nt_f16x3_tiles: K, lda, ldb must be multiples of 8");
  FPM_CHECK_ARG(((((size_t)A_hi) | ((size_t)A_lo) | ((size_t)Bt_hi) | ((size_t)Bt_lo) | ((size_t)tab)) & 15) == 0,
                "fpm_gemm_nt_f16x3_tiles: operands must be 16-byte aligned");
  if (M == 0 || max_tiles == 0) return FPM_OK;
  return launch_tc_pair<fpm::kF16x3>(A_hi, A_lo, Bt_hi, Bt_lo, inv_a, inv_b, nullptr, C, M, N, K, lda, ldb, ldc, 0,
                                     (cudaStream_t)stream, (const fpm::PairTile*)tab, tab_count, rowmap, max_tiles,
                                     m_ident);
}
