// Dense "NT" GEMM on the 5th-generation tensor cores:  C[M,N] = act(A[M,K] * Bt[N,K]^T + bias).
//
// This is the engine behind the true dense contractions of the matching head - SplineConv's slab GEMM
// ([sum n, 768] x [768, 26*768], ~3 TFLOP per batch of 256 pairs, replacing torch_spline_conv's per-edge
// weighting used by /root/reference/src/model/spline_conv.py:17,35,38) and the AFA-U projections
// (/root/reference/src/model/afau.py:98-102,124-139,189-199).
//
// Design (sm_100a only):
//   * operands are fp32 in HBM; tiles [128 x 32] (A) and [256 x 32] (B) are brought in by TMA
//     (cp.async.bulk.tensor, 128-byte swizzle) into a multi-stage shared-memory ring guarded by mbarriers;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=256, K=8) with the
//     accumulator (128 lanes x 256 fp32 columns) in tensor memory;
//   * four epilogue warps read the accumulator back with tcgen05.ld, add bias / relu and store fp32.
//   * passes = 3 ("3xTF32"): A and B are pre-split into tf32-exact hi and lo parts (a = hi + lo) and the
//     kernel accumulates  lo*hi + hi*lo + hi*hi  into the same TMEM tile, which restores fp32-level
//     accuracy (dropped term ~2^-22 relative) - needed because the head's outputs are compared to an fp32
//     reference at 1e-4 after three tau = 0.01 Sinkhorn amplifications.  passes = 1 is plain TF32.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "common.cuh"
#include <cuda.h>

namespace fpm {

constexpr int TBM = 128, TBN = 256, TBK = 32;          // tile; TBK floats = one 128-byte swizzle row
constexpr int UMMA_K = 8;                              // tf32
constexpr uint32_t kABytes = TBM * TBK * 4;            // 16 KB
constexpr uint32_t kBBytes = TBN * TBK * 4;            // 32 KB
constexpr int kTmemCols1 = 256, kTmemCols3 = 512;   // 1-pass: one accumulator; 3-pass: main + correction

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
// kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = 256.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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

template <int kPasses, int kStages>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const float* __restrict__ bias, float* __restrict__ Cm, int M, int N, int K, int ldc, int act) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kStageBytes = (kPasses == 3 ? 2 : 1) * (kABytes + kBBytes);
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = (uint32_t*)(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  const int nk = (K + TBK - 1) / TBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "n"(kPasses == 3 ? kTmemCols3 : kTmemCols1)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
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
        tma_load_2d(&tmB_hi, &full_bar[s], st + kABytes, kc, n0);
        if (kPasses == 3) {
          tma_load_2d(&tmA_lo, &full_bar[s], st + kABytes + kBBytes, kc, m0);
          tma_load_2d(&tmB_lo, &full_bar[s], st + 2 * kABytes + kBBytes, kc, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TBM, TBN);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
        const uint64_t a_hi = make_smem_desc(st), b_hi = make_smem_desc(st + kABytes);
        if (kPasses == 3) {
          const uint64_t a_lo = make_smem_desc(st + kABytes + kBBytes);
          const uint64_t b_lo = make_smem_desc(st + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
            // The tensor core's fp32 accumulate truncates, so the error grows with the number of MMAs
            // chained into one accumulator (measured: 288 chained MMAs -> 3e-5 relative).  The correction
            // terms therefore get their own accumulator (columns 256..511): their truncation error is
            // 2^-11 smaller, and the main chain shrinks 3x.  The epilogue adds the two in fp32 (RN).
            umma_tf32(tmem_base + TBN, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma_tf32(tmem_base + TBN, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
          }
        } else {
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
            umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
          }
        }
        umma_commit(&empty_bar[s]);           // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);             // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int m = m0 + q * 32 + lane;
    float* crow = Cm + (size_t)m * ldc;
    const bool vec_ok = ((ldc & 3) == 0) && ((((uintptr_t)Cm) & 15) == 0);
#pragma unroll 1
    for (int ch = 0; ch < TBN / 32; ++ch) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
      if (kPasses == 3) {
        uint32_t r2[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(TBN + ch * 32), r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int nb = n0 + ch * 32;
      if (m < M && nb < N) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int n = nb + g * 4;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x = __uint_as_float(r[g * 4 + j]);
            if (bias && n + j < N) x += bias[n + j];
            if (act == 1) x = fmaxf(x, 0.f);
            v[j] = x;
          }
          if (n + 3 < N && vec_ok) {
            *(float4*)(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < N) crow[n + j] = v[j];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kPasses == 3 ? kTmemCols3 : kTmemCols1)
                 : "memory");
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

static int make_map(CUtensorMap* map, const float* base, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { fpm_set_error("cuTensorMapEncodeTiled unavailable"); return FPM_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)fpm::TBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fpm_set_error("cuTensorMapEncodeTiled failed"); return FPM_ERR_ARG; }
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
  if (passes == 1) { A_lo = A_hi; Bt_lo = Bt_hi; }
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, A_hi, M, K, lda, fpm::TBM)) != FPM_OK) return rc;
  if ((rc = make_map(&mAl, A_lo, M, K, lda, fpm::TBM)) != FPM_OK) return rc;
  if ((rc = make_map(&mBh, Bt_hi, N, K, ldb, fpm::TBN)) != FPM_OK) return rc;
  if ((rc = make_map(&mBl, Bt_lo, N, K, ldb, fpm::TBN)) != FPM_OK) return rc;
  dim3 grid(fpm_cdiv(N, fpm::TBN), fpm_cdiv(M, fpm::TBM));
  FPM_CHECK_ARG(grid.y <= 65535, "fpm_gemm_nt_tc: M too large");
  if (passes == 3) {
    constexpr int kStages = 2;
    const size_t smem = (size_t)kStages * 2 * (fpm::kABytes + fpm::kBBytes) + 1024 + 256;
    FPM_CUDA(cudaFuncSetAttribute(fpm::gemm_tc_kernel<3, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fpm::gemm_tc_kernel<3, kStages><<<grid, 192, smem, st>>>(mAh, mAl, mBh, mBl, bias, C, M, N, K, ldc, act);
  } else {
    constexpr int kStages = 4;
    const size_t smem = (size_t)kStages * (fpm::kABytes + fpm::kBBytes) + 1024 + 256;
    FPM_CUDA(cudaFuncSetAttribute(fpm::gemm_tc_kernel<1, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fpm::gemm_tc_kernel<1, kStages><<<grid, 192, smem, st>>>(mAh, mAl, mBh, mBl, bias, C, M, N, K, ldc, act);
  }
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
