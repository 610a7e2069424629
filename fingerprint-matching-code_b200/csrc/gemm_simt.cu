// fp32 CUDA-core GEMMs: the reference-precision path.
//
//   fpm_gemm_nt_f32     C[M,N] = act(A[M,K] * Bt[N,K]^T + bias)   (nn.Linear convention)
//                       used for every dense contraction of the head when the tensor-core kernel
//                       (gemm_tcgen05.cu) is switched off, and as the on-device checker for it.
//   fpm_affinity        ragged batched  softplus((X1_b (.) c_b) X2_b^T) - 0.5  with optional edge
//                       rows  x[src] - x[dst]  formed on the fly; replaces
//                       /root/reference/src/model/affinity_layer.py:11-22 as used for Kp / Ke at
//                       /root/reference/src/model/ngm.py:277-287 (+ the zero padding of :317-318).
#include "common.cuh"

namespace fpm {

constexpr int GM = 128, GN = 128, GK = 16;

template <int ACT>   // 0 none, 1 relu
__global__ void __launch_bounds__(256)
gemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ Bt, const float* __restrict__ bias,
               float* __restrict__ Cm, int M, int N, int K, int lda, int ldb, int ldc) {
  __shared__ __align__(16) float As[2][GK][GM + 4];
  __shared__ __align__(16) float Bs[2][GK][GN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int lrow = tid >> 2, lk = (tid & 3) << 2;        // loader: row 0..63 (+64), k quad
  const int ty = tid >> 4, tx = tid & 15;                // compute: 16 x 16 threads, 8 x 8 each

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
      ra[h] = (gm < M && gk < K) ? *(const float4*)(A + (size_t)gm * lda + gk) : make_float4(0, 0, 0, 0);
      rb[h] = (gn < N && gk < K) ? *(const float4*)(Bt + (size_t)gn * ldb + gk) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  const int nk = (K + GK - 1) / GK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *(const float4*)&As[buf][k][ty * 4];
      const float4 a1 = *(const float4*)&As[buf][k][64 + ty * 4];
      const float4 b0 = *(const float4*)&Bs[buf][k][tx * 4];
      const float4 b1 = *(const float4*)&Bs[buf][k][64 + tx * 4];
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int gn = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (bias && gn + j < N) x += bias[gn + j];
        if (ACT == 1) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      float* dst = Cm + (size_t)gm * ldc + gn;
      if (gn + 3 < N && ((ldc & 3) == 0) && ((((size_t)Cm) & 15) == 0)) {
        *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < N) dst[j] = v[j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Ragged affinity.  Row r of pair b's left operand is
//     X[ptrA[b] + r]                                   (idxA0 == nullptr: node features), or
//     X[ptrA[b] + idxA0[eA[b] + r]] - X[ptrA[b] + idxA1[eA[b] + r]]   (edge features, spline_conv.py:73-81)
// scaled by coeff[b, :]; same for the right operand without the scaling.
// out[b, i, j] = scale * (softplus(dot) - 0.5) inside [nA_b, nB_b], 0 in the padding.
// ------------------------------------------------------------------------------------------
constexpr int AM = 64, AN = 64, AK = 16;

__global__ void __launch_bounds__(256)
affinity_kernel(const float* __restrict__ XA, const float* __restrict__ XB, const float* __restrict__ coeff,
                const int64_t* __restrict__ ptrA, const int64_t* __restrict__ ptrB,
                const int64_t* __restrict__ eptrA, const int64_t* __restrict__ eptrB,
                const int64_t* __restrict__ eidxA, const int64_t* __restrict__ eidxB, int EA, int EB,
                float* __restrict__ out, float* __restrict__ out_t, int Rmax, int Cmax, int Kdim,
                float scale, int raw) {
  __shared__ __align__(16) float As[AK][AM + 4];
  __shared__ __align__(16) float Bs[AK][AN + 4];
  const int b = blockIdx.z, tid = threadIdx.x;
  const int i0 = blockIdx.y * AM, j0 = blockIdx.x * AN;
  const bool edge_mode = eidxA != nullptr;
  const int nA = edge_mode ? (int)(eptrA[b + 1] - eptrA[b]) : (int)(ptrA[b + 1] - ptrA[b]);
  const int nB = edge_mode ? (int)(eptrB[b + 1] - eptrB[b]) : (int)(ptrB[b + 1] - ptrB[b]);
  const int lrow = tid >> 2, lk = (tid & 3) << 2;
  const int ty = tid >> 4, tx = tid & 15;
  const float* cb = coeff + (size_t)b * Kdim;

  // resolve this thread's operand rows once
  const float *pa0 = nullptr, *pa1 = nullptr, *pb0 = nullptr, *pb1 = nullptr;
  if (i0 + lrow < nA) {
    if (edge_mode) {
      // edge_index holds GLOBAL node ids (PyG batch), so no ptr offset is added
      pa0 = XA + (size_t)eidxA[eptrA[b] + i0 + lrow] * Kdim;
      pa1 = XA + (size_t)eidxA[(size_t)EA + eptrA[b] + i0 + lrow] * Kdim;
    } else {
      pa0 = XA + (size_t)(ptrA[b] + i0 + lrow) * Kdim;
    }
  }
  if (j0 + lrow < nB) {
    if (edge_mode) {
      pb0 = XB + (size_t)eidxB[eptrB[b] + j0 + lrow] * Kdim;
      pb1 = XB + (size_t)eidxB[(size_t)EB + eptrB[b] + j0 + lrow] * Kdim;
    } else {
      pb0 = XB + (size_t)(ptrB[b] + j0 + lrow) * Kdim;
    }
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (i0 < nA && j0 < nB) {
    for (int k0 = 0; k0 < Kdim; k0 += AK) {
      const int gk = k0 + lk;
      float4 a = make_float4(0, 0, 0, 0), bb = make_float4(0, 0, 0, 0);
      if (gk < Kdim) {
        if (pa0) {
          a = *(const float4*)(pa0 + gk);
          if (pa1) {
            const float4 a2 = *(const float4*)(pa1 + gk);
            a.x -= a2.x; a.y -= a2.y; a.z -= a2.z; a.w -= a2.w;
          }
          const float4 c4 = *(const float4*)(cb + gk);
          a.x = __fmul_rn(a.x, c4.x); a.y = __fmul_rn(a.y, c4.y);
          a.z = __fmul_rn(a.z, c4.z); a.w = __fmul_rn(a.w, c4.w);
        }
        if (pb0) {
          bb = *(const float4*)(pb0 + gk);
          if (pb1) {
            const float4 b2 = *(const float4*)(pb1 + gk);
            bb.x -= b2.x; bb.y -= b2.y; bb.z -= b2.z; bb.w -= b2.w;
          }
        }
      }
      __syncthreads();
      As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
      Bs[lk + 0][lrow] = bb.x; Bs[lk + 1][lrow] = bb.y; Bs[lk + 2][lrow] = bb.z; Bs[lk + 3][lrow] = bb.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < AK; ++k) {
        const float4 av = *(const float4*)&As[k][ty * 4];
        const float4 bv = *(const float4*)&Bs[k][tx * 4];
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    if (gi >= Rmax) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = j0 + tx * 4 + j;
      if (gj >= Cmax) continue;
      float v = 0.f;
      if (gi < nA && gj < nB) v = raw ? acc[i][j] : scale * (softplus_torch(acc[i][j]) - 0.5f);
      out[((size_t)b * Rmax + gi) * Cmax + gj] = v;
      if (out_t) out_t[((size_t)b * Cmax + gj) * Rmax + gi] = v;
    }
  }
}

// Edge affinity through linearity.  Edge features are differences of node features
// (/root/reference/src/model/spline_conv.py:73-81), so with P = (X1 (.) c) X2^T  [n1 x n2]:
//     (E1 (.) c) E2^T [k1, k2] = P[s1,s2] - P[s1,d2] - P[d1,s2] + P[d1,d2],   edge k = (s -> d)
// which replaces the reference's [e1 x 768] x [768 x e2] product per pair (0.50 GFLOP at n = 100) by a
// [n1 x 768] x [768 x n2] product (15 MFLOP) plus a 4-term gather: the kernel is bound by writing Ke.
// One CTA per (pair, kKeRows edge rows k1) and ALL columns k2.  The block first forms, for each of its rows, the
// difference of the two P rows it needs, D_r[j] = P[s1_r, j] - P[d1_r, j], transposed into shared memory as
// Dt[j][r] (row pitch kKeRows + 4 floats: 16-byte aligned and 8 distinct bank offsets).  A thread owns a column
// k2 = (s2 -> d2): two runs of 128-bit loads fetch D_.[s2] and D_.[d2] for all 32 rows, so an output costs
// 1/8 + 1/8 shared-memory load instead of four scalar loads at random banks.  Stores walk k2 across the warp
// (coalesced).  Before: one CTA per (128 columns, 16 rows) with the 32 P rows staged by a serial, dependent
// index -> row loop in every one of the five column blocks: 0.36 ms at B = 256, n = 100, seven times the cost of
// writing Ke.
constexpr int kKeRows = 32;
constexpr int kKeLd = kKeRows + 4;
__global__ void __launch_bounds__(192)
ke_factored_kernel(const float* __restrict__ P, const int64_t* __restrict__ eidxA,
                   const int64_t* __restrict__ eptrA, const int64_t* __restrict__ ptrA,
                   const int64_t* __restrict__ eidxB, const int64_t* __restrict__ eptrB,
                   const int64_t* __restrict__ ptrB, int EA, int EB, float* __restrict__ out, int Rn, int Cn,
                   int e1max, int e2max, float scale) {
  extern __shared__ __align__(16) float Dt[];   // [Cn][kKeLd]
  __shared__ int ends[2 * kKeRows];
  const int b = blockIdx.y, k1_0 = blockIdx.x * kKeRows;
  const int e1 = (int)(eptrA[b + 1] - eptrA[b]), e2 = (int)(eptrB[b + 1] - eptrB[b]);
  const float* Pb = P + (size_t)b * Rn * Cn;
  const int64_t pa = ptrA[b], ea0 = eptrA[b];
  const int nrow = min(kKeRows, e1 - k1_0);     // valid edge rows of this block (<= 0: padding only)
  if (threadIdx.x < 2 * kKeRows) {
    const int r = threadIdx.x >> 1, which = threadIdx.x & 1;
    ends[threadIdx.x] = r < nrow ? (int)(eidxA[(size_t)which * EA + ea0 + k1_0 + r] - pa) : 0;
  }
  __syncthreads();
  // warp per edge row, lanes along the node columns: coalesced row reads, several loads in flight, no division
  for (int r = threadIdx.x >> 5; r < kKeRows; r += blockDim.x >> 5) {
    const float* ps = Pb + (size_t)ends[2 * r] * Cn;
    const float* pd = Pb + (size_t)ends[2 * r + 1] * Cn;
    const bool on = r < nrow;
#pragma unroll 4
    for (int j = threadIdx.x & 31; j < Cn; j += 32) Dt[j * kKeLd + r] = on ? ps[j] - pd[j] : 0.f;
  }
  __syncthreads();
  const int64_t pb = ptrB[b], eb0 = eptrB[b];
  for (int k2 = threadIdx.x; k2 < e2max; k2 += blockDim.x) {
    int s2 = 0, d2 = 0;
    const bool col_ok = k2 < e2;
    if (col_ok) { s2 = (int)(eidxB[eb0 + k2] - pb); d2 = (int)(eidxB[(size_t)EB + eb0 + k2] - pb); }
    const float4* as = (const float4*)(Dt + s2 * kKeLd);
    const float4* ad = (const float4*)(Dt + d2 * kKeLd);
    float* ob = out + ((size_t)b * e1max + k1_0) * e2max + k2;
#pragma unroll
    for (int q = 0; q < kKeRows / 4; ++q) {
      const float4 x = as[q], y = ad[q];
      const float dot[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int r = 4 * q + t;
        if (k1_0 + r < e1max) {
          // Ke feeds nothing: branch-free softplus = max(x, 0) + log1p(e^-|x|) on the two
          // SFU approximations, |error| < 2e-6 (the log1p argument is formed as 1 + t: absolute, not relative, accuracy)
          const float v = (col_ok && r < nrow) ? scale * (softplus_sfu(dot[t]) - 0.5f) : 0.f;
          ob[(size_t)r * e2max] = v;
        }
      }
    }
  }
}

}  // namespace fpm

extern "C" int fpm_gemm_nt_f32(const float* A, const float* Bt, const float* bias, float* C, int M, int N,
                               int K, int lda, int ldb, int ldc, int act, void* stream) {
  FPM_CHECK_ARG(A && Bt && C, "fpm_gemm_nt_f32: null tensor");
  FPM_CHECK_ARG(M >= 0 && N > 0 && K > 0, "fpm_gemm_nt_f32: bad sizes");
  FPM_CHECK_ARG((K & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0, "fpm_gemm_nt_f32: K, lda, ldb must be multiples of 4");
  FPM_CHECK_ARG((((size_t)A) & 15) == 0 && (((size_t)Bt) & 15) == 0, "fpm_gemm_nt_f32: operands must be 16-byte aligned");
  FPM_CHECK_ARG(act == 0 || act == 1, "fpm_gemm_nt_f32: unknown activation");
  if (M == 0) return FPM_OK;
  dim3 grid(fpm_cdiv(N, fpm::GN), fpm_cdiv(M, fpm::GM));
  FPM_CHECK_ARG(grid.y <= 65535, "fpm_gemm_nt_f32: M too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (act == 0) fpm::gemm_nt_kernel<0><<<grid, 256, 0, st>>>(A, Bt, bias, C, M, N, K, lda, ldb, ldc);
  else fpm::gemm_nt_kernel<1><<<grid, 256, 0, st>>>(A, Bt, bias, C, M, N, K, lda, ldb, ldc);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_affinity(const float* XA, const float* XB, const float* coeff, const long long* ptrA,
                            const long long* ptrB, const long long* eptrA, const long long* eptrB,
                            const long long* eidxA, const long long* eidxB, int EA, int EB, float* out,
                            float* out_t, int B, int Rmax, int Cmax, int Kdim, float scale, int raw, void* stream) {
  FPM_CHECK_ARG(XA && XB && coeff && out, "fpm_affinity: null tensor");
  FPM_CHECK_ARG((eidxA == nullptr) == (eidxB == nullptr), "fpm_affinity: edge indices must be given for both sides");
  FPM_CHECK_ARG(eidxA ? (eptrA && eptrB) : (ptrA && ptrB), "fpm_affinity: missing offsets");
  FPM_CHECK_ARG((Kdim & 3) == 0, "fpm_affinity: feature dim must be a multiple of 4");
  FPM_CHECK_ARG(B >= 0 && Rmax > 0 && Cmax > 0, "fpm_affinity: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_affinity: batch too large");
  dim3 grid(fpm_cdiv(Cmax, fpm::AN), fpm_cdiv(Rmax, fpm::AM), B);
  fpm::affinity_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      XA, XB, coeff, (const int64_t*)ptrA, (const int64_t*)ptrB, (const int64_t*)eptrA,
      (const int64_t*)eptrB, (const int64_t*)eidxA, (const int64_t*)eidxB, EA, EB, out, out_t, Rmax, Cmax,
      Kdim, scale, raw);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_affinity_edges_factored(const float* P, const long long* eidxA, const long long* eptrA,
                                           const long long* ptrA, const long long* eidxB, const long long* eptrB,
                                           const long long* ptrB, int EA, int EB, float* out, int B, int Rn, int Cn,
                                           int e1max, int e2max, float scale, void* stream) {
  FPM_CHECK_ARG(P && eidxA && eptrA && ptrA && eidxB && eptrB && ptrB && out, "fpm_affinity_edges_factored: null tensor");
  FPM_CHECK_ARG(B >= 0 && Rn > 0 && Cn > 0 && e1max > 0 && e2max > 0, "fpm_affinity_edges_factored: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535 && e1max <= 65535, "fpm_affinity_edges_factored: batch or edge count too large");
  const size_t smem = (size_t)Cn * fpm::kKeLd * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_affinity_edges_factored: too many columns for shared memory");
  FPM_CUDA(cudaFuncSetAttribute(fpm::ke_factored_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(fpm_cdiv(e1max, fpm::kKeRows), B);
  fpm::ke_factored_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(
      P, (const int64_t*)eidxA, (const int64_t*)eptrA, (const int64_t*)ptrA, (const int64_t*)eidxB,
      (const int64_t*)eptrB, (const int64_t*)ptrB, EA, EB, out, Rn, Cn, e1max, e2max, scale);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
