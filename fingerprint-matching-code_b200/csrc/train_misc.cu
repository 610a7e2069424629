// Small dense helpers of the training path (BASELINE.json config 3: stage-1 step on genuine pairs).
//
// The reference gets every gradient from torch autograd; the kernels here are the pieces of the hand-written
// backward that are not tied to one forward kernel:
//   fpm_transpose_f32      [R, C] -> [C, ldo] (zero padded to ldo): the K-major operands of the SplineConv
//                          weight-gradient GEMM  dW = dY^T X  (src/model/spline_conv.py:17 via autograd)
//   fpm_bmm_ragged         per pair  Out[ptrO[b]+i, :] = sum_j Mat[b,i,j] * (X[ptrX[b]+j, :] (.) coeff[b, :])
//                          (or Mat^T): the two products of the affinity backward
//                          dX1 = c (.) (dP X2),  dX2 = dP^T (c (.) X1)   (src/model/affinity_layer.py:11-19)
//   fpm_segment_rowdot     out[b, :] = sum_{i in pair b} X[i, :] (.) Y[i, :]  (gradient of the per-pair coefficient)
#include "common.cuh"

namespace fpm {

__global__ void __launch_bounds__(256)
transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C, int ldo) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, c = c0 + tx;
    tile[ty + k][tx] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, r = r0 + tx;
    if (c < C && r < ldo) dst[(size_t)c * ldo + r] = tile[tx][ty + k];   // r >= R: zero padding
  }
}

// One CTA per (pair, 128-channel chunk): the pair's matrix sits in shared memory; every thread owns one channel
// and walks the output rows in blocks of 8 (the matrix entries are shared-memory broadcasts).
constexpr int BR = 8;
constexpr int kRowsPerCta = 32;
__global__ void __launch_bounds__(128)
bmm_ragged_kernel(const float* __restrict__ Mat, int Rmax, int Cmax, int trans, const float* __restrict__ X,
                  const int64_t* __restrict__ ptrX, const int64_t* __restrict__ ptrO,
                  const float* __restrict__ coeff_in, const float* __restrict__ coeff_out,
                  float* __restrict__ Out, int D) {
  extern __shared__ float sm[];                    // [Rmax * Cmax]
  const int b = blockIdx.y;
  const int d = blockIdx.x * 128 + threadIdx.x;
  const float* mb = Mat + (size_t)b * Rmax * Cmax;
  for (int i = threadIdx.x; i < Rmax * Cmax; i += blockDim.x) sm[i] = mb[i];
  __syncthreads();
  if (d >= D) return;
  const int nX = (int)(ptrX[b + 1] - ptrX[b]), nO = (int)(ptrO[b + 1] - ptrO[b]);
  const float* xb = X + (size_t)ptrX[b] * D + d;
  float* ob = Out + (size_t)ptrO[b] * D + d;
  const float ci = coeff_in ? coeff_in[(size_t)b * D + d] : 1.f;
  const float co = coeff_out ? coeff_out[(size_t)b * D + d] : 1.f;
  // blockIdx.z splits the output rows in chunks of kRowsPerCta: a small batch still fills the machine
  const int i_end = min(nO, (int)(blockIdx.z + 1) * kRowsPerCta);
  for (int i0 = blockIdx.z * kRowsPerCta; i0 < i_end; i0 += BR) {
    float acc[BR];
#pragma unroll
    for (int u = 0; u < BR; ++u) acc[u] = 0.f;
    for (int j = 0; j < nX; ++j) {
      const float x = xb[(size_t)j * D] * ci;
#pragma unroll
      for (int u = 0; u < BR; ++u) {
        const int i = min(i0 + u, nO - 1);
        const float m = trans ? sm[j * Cmax + i] : sm[i * Cmax + j];
        acc[u] = fmaf(m, x, acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < BR; ++u)
      if (i0 + u < nO) ob[(size_t)(i0 + u) * D] = acc[u] * co;
  }
}

__global__ void __launch_bounds__(128)
segment_rowdot_kernel(const float* __restrict__ X, const float* __restrict__ Y, const int64_t* __restrict__ ptr,
                      float* __restrict__ out, int D) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * 128 + threadIdx.x;
  if (d >= D) return;
  const int64_t r0 = ptr[b], r1 = ptr[b + 1];
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc = fmaf(X[(size_t)r * D + d], Y[(size_t)r * D + d], acc);
  out[(size_t)b * D + d] = acc;
}

}  // namespace fpm

extern "C" int fpm_transpose_f32(const float* src, float* dst, int R, int C, int ldo, void* stream) {
  FPM_CHECK_ARG(src && dst, "fpm_transpose_f32: null tensor");
  FPM_CHECK_ARG(R >= 0 && C > 0 && ldo >= R, "fpm_transpose_f32: bad sizes");
  if (ldo == 0) return FPM_OK;
  dim3 grid(fpm_cdiv(C, 32), fpm_cdiv(ldo, 32));
  FPM_CHECK_ARG(grid.y <= 65535, "fpm_transpose_f32: too many rows");
  fpm::transpose_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, R, C, ldo);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_bmm_ragged(const float* Mat, int B, int Rmax, int Cmax, int trans, const float* X,
                              const long long* ptrX, const long long* ptrO, const float* coeff_in,
                              const float* coeff_out, float* Out, int D, void* stream) {
  FPM_CHECK_ARG(Mat && X && ptrX && ptrO && Out, "fpm_bmm_ragged: null tensor");
  FPM_CHECK_ARG(B >= 0 && Rmax > 0 && Cmax > 0 && D > 0, "fpm_bmm_ragged: bad sizes");
  if (B == 0) return FPM_OK;
  const size_t smem = (size_t)Rmax * Cmax * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024 && B <= 65535, "fpm_bmm_ragged: matrix or batch too large");
  FPM_CUDA(cudaFuncSetAttribute(fpm::bmm_ragged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int Omax = trans ? Cmax : Rmax;               // output rows of a pair never exceed the matrix side
  fpm::bmm_ragged_kernel<<<dim3(fpm_cdiv(D, 128), B, fpm_cdiv(Omax, fpm::kRowsPerCta)), 128, smem, (cudaStream_t)stream>>>(
      Mat, Rmax, Cmax, trans, X, (const int64_t*)ptrX, (const int64_t*)ptrO, coeff_in, coeff_out, Out, D);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_segment_rowdot(const float* X, const float* Y, const long long* ptr, float* out, int B, int D,
                                  void* stream) {
  FPM_CHECK_ARG(X && Y && ptr && out, "fpm_segment_rowdot: null tensor");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_segment_rowdot: batch too large");
  fpm::segment_rowdot_kernel<<<dim3(fpm_cdiv(D, 128), B), 128, 0, (cudaStream_t)stream>>>(
      X, Y, (const int64_t*)ptr, out, D);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
