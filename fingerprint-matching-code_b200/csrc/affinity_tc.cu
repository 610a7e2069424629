// Node / edge affinities on the tensor cores.
//
// InnerProductWithWeightsAffinity (/root/reference/src/model/affinity_layer.py:11-19) evaluates, per pair,
//     softplus((X1_b (.) c_b) X2_b^T) - 0.5,       X1_b [n1_b, 768], X2_b [n2_b, 768], c_b = tanh(A w_b + a)
// a true dense contraction per pair.  gemm_simt.cu::affinity_kernel runs it on the CUDA cores (64 x 64 x 16 tiles,
// FMA pipe 44 % busy, 0.23 ms per launch at 256 pairs x 100 keypoints, twice per forward: Kp and the raw products the
// factored Ke is built from).  Here the products go through the persistent CTA-pair tcgen05 GEMM of gemm_tcgen05.cu
// in its tile-table form, with the same error-compensated fp16 operands as the SplineConv slabs (fp32-faithful):
//   1. f16_split_rows_scaled_kernel: rows of X1 times their pair's coefficient vector -> fp16 hi / lo + row scale
//      (X2 goes through the plain fpm_f16_split_rows),
//   2. affinity_tiles_kernel: one tile-table entry per (pair, 256-row block of X1_b, 128-row block of X2_b): A rows
//      start at ptr1[b], Bt rows at ptr2[b]; rows past the pair's end belong to the next pair (or are TMA zero fill)
//      and are dropped by the row map / ignored by step 4,
//   3. fpm_gemm_nt_f16x3_tiles: raw products into a scratch [B * n1max, 128 * tilesB],
//   4. affinity_finish_kernel: softplus - 0.5 (or the raw value), zero padding, the padded and the transposed copy.
#include "common.cuh"
#include <cuda_fp16.h>

namespace fpm {

// One warp per row; the row's pair is found by bisection of ptr (all lanes take the same path).
__global__ void __launch_bounds__(256)
f16_split_rows_scaled_kernel(const float* __restrict__ src, const float* __restrict__ coeff,
                             const int64_t* __restrict__ ptr, int B, __half* __restrict__ hi, __half* __restrict__ lo,
                             float* __restrict__ inv_scale, int rows, int K) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  int lo_b = 0, hi_b = B;                      // ptr[lo_b] <= row < ptr[hi_b]
  while (hi_b - lo_b > 1) {
    const int mid = (lo_b + hi_b) >> 1;
    if ((int64_t)row >= ptr[mid]) lo_b = mid; else hi_b = mid;
  }
  const float4* s4 = (const float4*)(src + (size_t)row * K);
  const float4* c4 = (const float4*)(coeff + (size_t)lo_b * K);
  const int n4 = K >> 2;
  float amax = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = s4[i], c = c4[i];
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(__fmul_rn(v.x, c.x)), fabsf(__fmul_rn(v.y, c.y))),
                             fmaxf(fabsf(__fmul_rn(v.z, c.z)), fabsf(__fmul_rn(v.w, c.w)))));
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
    const float4 v = s4[i], c = c4[i];
    // (x * c) first, rounded to fp32 as the reference's elementwise product is, then the power-of-two scale (exact)
    const float x[4] = {__fmul_rn(v.x, c.x) * s, __fmul_rn(v.y, c.y) * s, __fmul_rn(v.z, c.z) * s,
                        __fmul_rn(v.w, c.w) * s};
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

// grid B, block 256.  tab [B * tA * tB][4] = {first A row, first Bt row, first scratch column, rowmap offset};
// rowmap [B * tA * 256]: scratch row b * Rmax + i for row i of the pair, -1 beyond its n1_b rows.
__global__ void __launch_bounds__(256)
affinity_tiles_kernel(const int64_t* __restrict__ ptrA, const int64_t* __restrict__ ptrB, int tA, int tB, int Rmax,
                      int* __restrict__ tab, int* __restrict__ tab_count, int* __restrict__ rowmap) {
  const int b = blockIdx.x;
  const int a0 = (int)ptrA[b], nA = (int)(ptrA[b + 1] - ptrA[b]);
  const int b0 = (int)ptrB[b];
  for (int t = threadIdx.x; t < tA * tB; t += blockDim.x) {
    const int ta = t / tB, tb = t - ta * tB;
    int* e = tab + ((size_t)b * tA * tB + t) * 4;
    e[0] = a0 + 256 * ta; e[1] = b0 + 128 * tb; e[2] = 128 * tb; e[3] = (b * tA + ta) * 256;
  }
  for (int r = threadIdx.x; r < tA * 256; r += blockDim.x)
    rowmap[(size_t)b * tA * 256 + r] = r < nA ? b * Rmax + r : -1;
  if (b == 0 && threadIdx.x == 0) *tab_count = (int)gridDim.x * tA * tB;
}

// P [B * Rmax, ldp] raw products -> out [B, Rmax, Cmax] (and out_t [B, Cmax, Rmax]) with the activation and the
// zero padding of the reference's pad_tensor.  32 x 32 tiles through shared memory so both copies store coalesced.
__global__ void __launch_bounds__(256)
affinity_finish_kernel(const float* __restrict__ P, const int64_t* __restrict__ ptrA,
                       const int64_t* __restrict__ ptrB, float* __restrict__ out, float* __restrict__ out_t, int Rmax,
                       int Cmax, int ldp, float scale, int raw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int nA = (int)(ptrA[b + 1] - ptrA[b]), nB = (int)(ptrB[b + 1] - ptrB[b]);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    float v = 0.f;
    if (i < nA && j < nB) {
      const float p = P[((size_t)b * Rmax + i) * ldp + j];
      v = raw ? p : scale * (softplus_torch(p) - 0.5f);
    }
    if (i < Rmax && j < Cmax) out[((size_t)b * Rmax + i) * Cmax + j] = v;
    tile[r][tx] = v;
  }
  if (out_t == nullptr) return;
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + tx;
    if (j < Cmax && i < Rmax) out_t[((size_t)b * Cmax + j) * Rmax + i] = tile[tx][r];
  }
}

}  // namespace fpm

extern "C" int fpm_f16_split_rows_scaled(const float* src, const float* coeff, const long long* ptr, int B, void* hi,
                                         void* lo, float* inv_scale, int rows, int K, void* stream) {
  FPM_CHECK_ARG(src && coeff && ptr && hi && lo && inv_scale, "fpm_f16_split_rows_scaled: null tensor");
  FPM_CHECK_ARG(rows >= 0 && B > 0 && K > 0 && (K & 7) == 0, "fpm_f16_split_rows_scaled: K must be a multiple of 8");
  FPM_CHECK_ARG(((((size_t)src) | ((size_t)coeff) | ((size_t)hi) | ((size_t)lo)) & 15) == 0,
                "fpm_f16_split_rows_scaled: 16-byte alignment required");
  if (rows == 0) return FPM_OK;
  fpm::f16_split_rows_scaled_kernel<<<fpm_cdiv((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      src, coeff, (const int64_t*)ptr, B, (__half*)hi, (__half*)lo, inv_scale, rows, K);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// tab: [B * tA * tB, 4] int32 with tA = ceil(Rmax / 256), tB = ceil(Cmax / 128); rowmap: [B * tA * 256] int32.
extern "C" int fpm_affinity_tiles(const long long* ptrA, const long long* ptrB, int B, int Rmax, int Cmax, int* tab,
                                  int* tab_count, int* rowmap, void* stream) {
  FPM_CHECK_ARG(ptrA && ptrB && tab && tab_count && rowmap, "fpm_affinity_tiles: null tensor");
  FPM_CHECK_ARG(B >= 0 && Rmax > 0 && Cmax > 0, "fpm_affinity_tiles: bad sizes");
  if (B == 0) return FPM_OK;
  const int tA = fpm_cdiv(Rmax, 256), tB = fpm_cdiv(Cmax, 128);
  fpm::affinity_tiles_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const int64_t*)ptrA, (const int64_t*)ptrB, tA, tB,
                                                                  Rmax, tab, tab_count, rowmap);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_affinity_finish(const float* P, const long long* ptrA, const long long* ptrB, float* out,
                                   float* out_t, int B, int Rmax, int Cmax, int ldp, float scale, int raw,
                                   void* stream) {
  FPM_CHECK_ARG(P && ptrA && ptrB && out, "fpm_affinity_finish: null tensor");
  FPM_CHECK_ARG(B >= 0 && Rmax > 0 && Cmax > 0 && ldp >= Cmax, "fpm_affinity_finish: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_affinity_finish: batch too large");
  dim3 grid(fpm_cdiv(Cmax, 32), fpm_cdiv(Rmax, 32), B);
  fpm::affinity_finish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, (const int64_t*)ptrA, (const int64_t*)ptrB,
                                                                     out, out_t, Rmax, Cmax, ldp, scale, raw);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
