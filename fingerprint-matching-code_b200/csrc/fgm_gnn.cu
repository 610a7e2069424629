// Dense NGM-v1 message passing: the aggregation of GNNLayer.forward
// (/root/reference/src/model/gnn.py:54-68)
//     A  <- F.normalize(A, p=1, dim=2)                      (row scale 1 / max(sum_j |A[i,j]|, 1e-12), when norm)
//     x2[b,i,c] = sum_j A[b,i,j] * W[b,i,j,c'] * x1[b,j,c]   c' = c when the edge tensor has one channel per node
//                                                            channel (edge_emb), c' = 0 when it has a single channel
// which the reference writes as a permuted torch.matmul over a materialised [b, N, N, fe] product.  Here one CTA owns
// one output row: the coefficient a_ij * w_ij is formed on the fly, nothing of size N x N x fe is written.
// The same kernel evaluates the transposed product for the backward (dx1 = coef^T dx2), and a second one the
// gradient of the edge tensor.  Memory-bound: A and W are read once per call.
#include "common.cuh"

namespace fpm {

constexpr int kFgmMaxF = 32;
constexpr int kFgmThreads = 128;

// grid (N, B).  trans = 0: out[b,r,:] = sum_q A[r,q] inv[r] W[r,q,:] X[q,:]   (inv computed here when norm, and stored)
//               trans = 1: out[b,r,:] = sum_q A[q,r] inv[q] W[q,r,:] X[q,:]   (inv read; 1 when null)
template <int F>
__global__ void __launch_bounds__(kFgmThreads)
fgm_aggregate_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ X,
                     float* __restrict__ inv, float* __restrict__ out, int N, int fe, int norm, int trans) {
  __shared__ float red[32];
  __shared__ float part[kFgmThreads / 32][F];
  const int b = blockIdx.y, r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* Ab = A + (size_t)b * N * N;
  const float* Wb = W + (size_t)b * N * N * fe;
  const float* Xb = X + (size_t)b * N * F;
  float scale = 1.f;
  if (!trans && norm) {
    float s = 0.f;
    for (int q = tid; q < N; q += kFgmThreads) s += fabsf(Ab[(size_t)r * N + q]);
    s = block_sum(s, red);
    scale = 1.f / fmaxf(s, 1e-12f);
    if (inv && tid == 0) inv[(size_t)b * N + r] = scale;
  }
  float acc[F];
#pragma unroll
  for (int c = 0; c < F; ++c) acc[c] = 0.f;
  for (int q = tid; q < N; q += kFgmThreads) {
    const size_t e = trans ? (size_t)q * N + r : (size_t)r * N + q;
    float a = Ab[e];
    if (a == 0.f) continue;                                  // adjacency masks are sparse
    a *= trans ? (inv ? inv[(size_t)b * N + q] : 1.f) : scale;
    const float* w = Wb + e * fe;
    const float* x = Xb + (size_t)q * F;
    if (fe == 1) {
      const float aw = a * w[0];
#pragma unroll
      for (int c = 0; c < F; ++c) acc[c] = fmaf(aw, x[c], acc[c]);
    } else {
#pragma unroll
      for (int c = 0; c < F; ++c) acc[c] = fmaf(a * w[c], x[c], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < F; ++c) {
    const float v = warp_sum(acc[c]);
    if (lane == 0) part[warp][c] = v;
  }
  __syncthreads();
  if (tid < F) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kFgmThreads / 32; ++w) v += part[w][tid];
    out[((size_t)b * N + r) * F + tid] = v;
  }
}

// grid (cdiv(N, 128), N, B): dW[b,i,j,c'] = A[i,j] inv[i] * (fe == 1 ? sum_c dx2[i,c] x1[j,c] : dx2[i,c'] x1[j,c'])
template <int F>
__global__ void __launch_bounds__(128)
fgm_aggregate_dw_kernel(const float* __restrict__ A, const float* __restrict__ inv, const float* __restrict__ dx2,
                        const float* __restrict__ x1, float* __restrict__ dW, int N, int fe) {
  __shared__ float g[F];
  const int b = blockIdx.z, i = blockIdx.y, j = blockIdx.x * 128 + threadIdx.x;
  if (threadIdx.x < F) g[threadIdx.x] = dx2[((size_t)b * N + i) * F + threadIdx.x];
  __syncthreads();
  if (j >= N) return;
  const size_t e = ((size_t)b * N + i) * N + j;
  const float a = A[e] * (inv ? inv[(size_t)b * N + i] : 1.f);
  const float* x = x1 + ((size_t)b * N + j) * F;
  if (fe == 1) {
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < F; ++c) d = fmaf(g[c], x[c], d);
    dW[e] = a * d;
  } else {
#pragma unroll
    for (int c = 0; c < F; ++c) dW[e * F + c] = a * g[c] * x[c];
  }
}

}  // namespace fpm

#define FPM_FGM_DISPATCH(F, CALL)              \
  switch (F) {                                 \
    case 1: { constexpr int kF_ = 1; CALL; } break;   \
    case 2: { constexpr int kF_ = 2; CALL; } break;   \
    case 4: { constexpr int kF_ = 4; CALL; } break;   \
    case 8: { constexpr int kF_ = 8; CALL; } break;   \
    case 16: { constexpr int kF_ = 16; CALL; } break; \
    case 32: { constexpr int kF_ = 32; CALL; } break; \
    default:                                   \
      fpm_set_error("fpm_fgm_aggregate: feature width must be 1, 2, 4, 8, 16 or 32"); \
      return FPM_ERR_UNSUPPORTED;              \
  }

extern "C" int fpm_fgm_aggregate(const float* A, const float* W, const float* X, float* inv, float* out, int B, int N,
                                 int F, int fe, int norm, int trans, void* stream) {
  FPM_CHECK_ARG(A && W && X && out, "fpm_fgm_aggregate: null tensor");
  FPM_CHECK_ARG(B >= 0 && N > 0 && (fe == 1 || fe == F), "fpm_fgm_aggregate: edge channels must be 1 or the node width");
  FPM_CHECK_ARG(B <= 65535, "fpm_fgm_aggregate: batch too large");
  if (B == 0) return FPM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FPM_FGM_DISPATCH(F, (fpm::fgm_aggregate_kernel<kF_><<<dim3(N, B), fpm::kFgmThreads, 0, st>>>(A, W, X, inv, out, N, fe,
                                                                                               norm, trans)));
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_fgm_aggregate_dw(const float* A, const float* inv, const float* dx2, const float* x1, float* dW,
                                    int B, int N, int F, int fe, void* stream) {
  FPM_CHECK_ARG(A && dx2 && x1 && dW, "fpm_fgm_aggregate_dw: null tensor");
  FPM_CHECK_ARG(B >= 0 && N > 0 && N <= 65535 && (fe == 1 || fe == F), "fpm_fgm_aggregate_dw: bad sizes");
  FPM_CHECK_ARG(B <= 65535, "fpm_fgm_aggregate_dw: batch too large");
  if (B == 0) return FPM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FPM_FGM_DISPATCH(F, (fpm::fgm_aggregate_dw_kernel<kF_><<<dim3(fpm_cdiv(N, 128), N, B), 128, 0, st>>>(A, inv, dx2, x1, dW,
                                                                                                     N, fe)));
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
