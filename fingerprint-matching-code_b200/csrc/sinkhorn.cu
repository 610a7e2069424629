// Batched log-domain Sinkhorn and soft-top-k: one CTA per fingerprint pair, the pair's score
// matrix resident in shared memory for every iteration (global scratch only when it cannot fit).
//
// Replaces, on the hot path of /root/reference/src/model/ngm.py:
//   * Sinkhorn.forward -> pygmtools.sinkhorn(backend='pytorch')      (src/model/sinkhorn.py:85-87;
//     call sites src/model/gnn.py:219 (20 it) and src/model/ngm.py:371 (10 it))
//   * soft_topk + Sinkhorn_m.forward_log                              (src/model/soft_topk.py:8-53,166-255)
// HBM-bound by design: algorithmic traffic is one read and one write of the n1 x n2 matrix per call
// (SURVEY.md section 8d); all 10/20 iterations run on chip.
#include "common.cuh"

namespace fpm {

// ------------------------------------------------------------------------------------------
// Sinkhorn.  Semantics of pygmtools 0.5.3 `sinkhorn` with batched_operation=False:
//   frame transpose when C < R, per-sample transpose when n1_b > n2_b, log_s = s / tau,
//   dummy rows (= -100) up to a square n2_b x n2_b problem, alternate row / column
//   log-normalisation starting with rows, crop, exp.  Padding comes back as exact zeros.
// ------------------------------------------------------------------------------------------
template <bool kGlobal>
__global__ void __launch_bounds__(512)
sinkhorn_log_kernel(const float* __restrict__ s, const int64_t* __restrict__ n1,
                    const int64_t* __restrict__ n2, float* __restrict__ out,
                    float* __restrict__ out_t, float* __restrict__ workspace, int R, int C,
                    int max_iter, float tau, int dummy_row) {
  extern __shared__ float smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int D = R > C ? R : C;

  int n1b = n1 ? (int)n1[b] : R;
  int n2b = n2 ? (int)n2[b] : C;
  n1b = min(max(n1b, 0), R);
  n2b = min(max(n2b, 0), C);
  const bool frameT = C < R;
  const bool opT = frameT ? (n1b >= n2b) : (n1b > n2b);
  const int nr = opT ? n2b : n1b;          // rows of the per-sample problem (nr <= nc)
  const int nc = opT ? n1b : n2b;
  const int rows = dummy_row ? nc : nr;    // dummy rows square the problem

  float* M = kGlobal ? workspace + (size_t)b * D * D : smem;
  float* part = kGlobal ? smem : smem + (size_t)D * D;   // [nwarps][D] partials for column passes
  const float* sb = s + (size_t)b * R * C;
  const int ld = nc;

  for (int idx = tid; idx < rows * nc; idx += nthreads) {
    const int i = idx / nc, j = idx - i * nc;
    float v = -100.0f;
    if (i < nr) v = (opT ? sb[(size_t)j * C + i] : sb[(size_t)i * C + j]) / tau;
    M[idx] = v;
  }
  __syncthreads();

  const int chunks = (nc + 31) >> 5;
  const int wpc = max(1, nwarps / max(chunks, 1));     // warps cooperating on one 32-column chunk
  const int groups = nwarps / wpc;
  const int slot = warp / wpc, sub = warp % wpc;

  for (int it = 0; it < max_iter; ++it) {
    if ((it & 1) == 0) {
      // row normalisation: one warp per row
      for (int r = warp; r < rows; r += nwarps) {
        float* row = M + (size_t)r * ld;
        float mx = kNegInf;
        for (int j = lane; j < nc; j += 32) mx = fmaxf(mx, row[j]);
        mx = warp_max(mx);
        const float sh = (mx == kNegInf) ? 0.f : mx;
        float sum = 0.f;
        for (int j = lane; j < nc; j += 32) sum += expf(row[j] - sh);
        sum = warp_sum(sum);
        const float lse = logf(sum) + sh;
        for (int j = lane; j < nc; j += 32) row[j] = row[j] - lse;
      }
      __syncthreads();
    } else {
      // column normalisation: lane <-> column inside a 32-column chunk, `wpc` warps split the rows
      if (slot < groups) {
        for (int ch = slot; ch < chunks; ch += groups) {
          const int j = (ch << 5) + lane;
          float mx = kNegInf;
          if (j < nc)
            for (int r = sub; r < rows; r += wpc) mx = fmaxf(mx, M[(size_t)r * ld + j]);
          if (j < nc) part[sub * D + j] = mx;
        }
      }
      __syncthreads();
      if (slot < groups) {
        for (int ch = slot; ch < chunks; ch += groups) {
          const int j = (ch << 5) + lane;
          if (j < nc) {
            float mx = kNegInf;
            for (int w = 0; w < wpc; ++w) mx = fmaxf(mx, part[w * D + j]);
            const float sh = (mx == kNegInf) ? 0.f : mx;
            float sum = 0.f;
            for (int r = sub; r < rows; r += wpc) sum += expf(M[(size_t)r * ld + j] - sh);
            part[(wpc + sub) * D + j] = sum;
            if (sub == 0) part[2 * wpc * D + j] = sh;
          }
        }
      }
      __syncthreads();
      if (slot < groups) {
        for (int ch = slot; ch < chunks; ch += groups) {
          const int j = (ch << 5) + lane;
          if (j < nc) {
            float sum = 0.f;
            for (int w = 0; w < wpc; ++w) sum += part[(wpc + w) * D + j];
            const float lse = logf(sum) + part[2 * wpc * D + j];
            for (int r = sub; r < rows; r += wpc) M[(size_t)r * ld + j] -= lse;
          }
        }
      }
      __syncthreads();
    }
  }

  float* ob = out + (size_t)b * R * C;
  for (int idx = tid; idx < R * C; idx += nthreads) {
    const int a = idx / C, c = idx - a * C;
    float v = 0.f;
    if (a < n1b && c < n2b) {
      const int fi = opT ? c : a, fj = opT ? a : c;
      v = expf(M[(size_t)fi * ld + fj]);
    }
    ob[idx] = v;
  }
  if (out_t) {
    float* otb = out_t + (size_t)b * R * C;
    for (int idx = tid; idx < R * C; idx += nthreads) {
      const int c = idx / R, a = idx - c * R;       // out_t[b][c][a]
      float v = 0.f;
      if (a < n1b && c < n2b) {
        const int fi = opT ? c : a, fj = opT ? a : c;
        v = expf(M[(size_t)fi * ld + fj]);
      }
      otb[idx] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// soft-top-k: optimal transport between the n1_b*n2_b scores and the two anchors {min, max} with
// column marginals (N - k, k); returns exp(log-plan[:, 1]) reshaped to n1_b x n2_b.
// Follows Sinkhorn_m.forward_log's non-batched branch including the NaN -> -inf clean-up after every
// half-step and the "while any(log_s > 0)" continuation (soft_topk.py:217-243).
// ------------------------------------------------------------------------------------------
template <bool kGlobal>
__global__ void __launch_bounds__(512)
soft_topk_kernel(const float* __restrict__ scores, const float* __restrict__ ks,
                 const int64_t* __restrict__ n1, const int64_t* __restrict__ n2,
                 float* __restrict__ out, float* __restrict__ workspace, int R, int C,
                 int max_iter, float tau) {
  extern __shared__ float smem[];
  __shared__ float red[64];
  const int b = blockIdx.x, tid = threadIdx.x, nthreads = blockDim.x;
  int n1b = n1 ? (int)n1[b] : R;
  int n2b = n2 ? (int)n2[b] : C;
  n1b = min(max(n1b, 0), R);
  n2b = min(max(n2b, 0), C);
  const int N = n1b * n2b;
  float* L0 = kGlobal ? workspace + (size_t)b * 2 * R * C : smem;
  float* L1 = L0 + (size_t)R * C;
  const float* sb = scores + (size_t)b * R * C;

  // anchors = (min, max) over the valid block
  float mn = INFINITY, mx = kNegInf;
  for (int p = tid; p < N; p += nthreads) {
    const int i = p / n2b, j = p - i * n2b;
    const float v = sb[(size_t)i * C + j];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = block_min(mn, red);
  mx = block_max(mx, red + 32);
  for (int p = tid; p < N; p += nthreads) {
    const int i = p / n2b, j = p - i * n2b;
    const float v = sb[(size_t)i * C + j];
    L0[p] = (-fabsf(v - mn)) / tau;
    L1[p] = (-fabsf(v - mx)) / tau;
  }
  const float k = ks[b];
  const float lc0 = logf((float)((long long)n1b * (long long)n2b) - k);
  const float lc1 = logf(k);
  __syncthreads();

  int it = 0;
  while (true) {
    if (it >= max_iter) {
      int pos = 0;
      for (int p = tid; p < N; p += nthreads) pos |= (L0[p] > 0.f) | (L1[p] > 0.f);
      if (!__syncthreads_or(pos)) break;
    }
    if ((it & 1) == 0) {
      for (int p = tid; p < N; p += nthreads) {
        const float a = L0[p], c = L1[p];
        const float m = fmaxf(a, c);
        const float sh = (m == kNegInf || m == INFINITY) ? 0.f : m;
        const float lse = logf(expf(a - sh) + expf(c - sh)) + sh;
        float na = a - lse + 0.0f, nc_ = c - lse + 0.0f;
        L0[p] = isnan(na) ? kNegInf : na;
        L1[p] = isnan(nc_) ? kNegInf : nc_;
      }
      __syncthreads();
    } else {
      float m0 = kNegInf, m1 = kNegInf;
      for (int p = tid; p < N; p += nthreads) {
        m0 = fmaxf(m0, L0[p]);
        m1 = fmaxf(m1, L1[p]);
      }
      m0 = block_max(m0, red);
      m1 = block_max(m1, red + 32);
      const float sh0 = (m0 == kNegInf || m0 == INFINITY) ? 0.f : m0;
      const float sh1 = (m1 == kNegInf || m1 == INFINITY) ? 0.f : m1;
      float s0 = 0.f, s1 = 0.f;
      for (int p = tid; p < N; p += nthreads) {
        s0 += expf(L0[p] - sh0);
        s1 += expf(L1[p] - sh1);
      }
      s0 = block_sum(s0, red);
      s1 = block_sum(s1, red + 32);
      const float lse0 = logf(s0) + sh0, lse1 = logf(s1) + sh1;
      for (int p = tid; p < N; p += nthreads) {
        float na = L0[p] - lse0 + lc0, nc_ = L1[p] - lse1 + lc1;
        L0[p] = isnan(na) ? kNegInf : na;
        L1[p] = isnan(nc_) ? kNegInf : nc_;
      }
      __syncthreads();
    }
    ++it;
    if (it > max_iter + 64) break;   // cannot happen (a row step makes every entry <= 0)
  }

  float* ob = out + (size_t)b * R * C;
  for (int idx = tid; idx < R * C; idx += nthreads) {
    const int i = idx / C, j = idx - i * C;
    ob[idx] = (i < n1b && j < n2b) ? expf(L1[(size_t)i * n2b + j]) : 0.f;
  }
}

}  // namespace fpm

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
static const size_t kSmemLimit = 227 * 1024;

extern "C" long long fpm_sinkhorn_workspace_bytes(int B, int R, int C, int dummy_row) {
  const int D = R > C ? R : C;
  const int threads = 512;
  const size_t need = ((size_t)D * D + (size_t)(3 * (threads / 32)) * D) * sizeof(float);
  (void)dummy_row;
  return need <= kSmemLimit ? 0 : (long long)B * D * D * (long long)sizeof(float);
}

extern "C" int fpm_sinkhorn_log(const float* s, const long long* n1, const long long* n2, float* out,
                                float* out_t, void* workspace, int B, int R, int C, int max_iter,
                                float tau, int dummy_row, void* stream) {
  FPM_CHECK_ARG(s && out, "fpm_sinkhorn_log: null tensor");
  FPM_CHECK_ARG(B >= 0 && R > 0 && C > 0 && max_iter >= 0, "fpm_sinkhorn_log: bad sizes");
  FPM_CHECK_ARG(tau != 0.f, "fpm_sinkhorn_log: tau must be non-zero");
  if (B == 0) return FPM_OK;
  const int D = R > C ? R : C;
  const int threads = D <= 48 ? 128 : (D <= 96 ? 256 : 512);
  const int nwarps = threads / 32;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t part = (size_t)(3 * nwarps) * D * sizeof(float);
  const size_t full = (size_t)D * D * sizeof(float) + part;
  if (full <= kSmemLimit) {
    FPM_CUDA(cudaFuncSetAttribute(fpm::sinkhorn_log_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)full));
    fpm::sinkhorn_log_kernel<false><<<B, threads, full, st>>>(
        s, (const int64_t*)n1, (const int64_t*)n2, out, out_t, nullptr, R, C, max_iter, tau, dummy_row);
  } else {
    FPM_CHECK_ARG(workspace, "fpm_sinkhorn_log: matrix exceeds shared memory, workspace required");
    FPM_CUDA(cudaFuncSetAttribute(fpm::sinkhorn_log_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part));
    fpm::sinkhorn_log_kernel<true><<<B, threads, part, st>>>(
        s, (const int64_t*)n1, (const int64_t*)n2, out, out_t, (float*)workspace, R, C, max_iter, tau,
        dummy_row);
  }
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" long long fpm_soft_topk_workspace_bytes(int B, int R, int C) {
  const size_t need = (size_t)2 * R * C * sizeof(float);
  return need <= kSmemLimit - 1024 ? 0 : (long long)B * 2 * R * C * (long long)sizeof(float);
}

extern "C" int fpm_soft_topk(const float* scores, const float* ks, const long long* n1,
                             const long long* n2, float* out, void* workspace, int B, int R, int C,
                             int max_iter, float tau, void* stream) {
  FPM_CHECK_ARG(scores && ks && out, "fpm_soft_topk: null tensor");
  FPM_CHECK_ARG(B >= 0 && R > 0 && C > 0 && max_iter >= 0, "fpm_soft_topk: bad sizes");
  FPM_CHECK_ARG(tau != 0.f, "fpm_soft_topk: tau must be non-zero");
  if (B == 0) return FPM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = (R * C) <= 4096 ? 256 : 512;
  const size_t need = (size_t)2 * R * C * sizeof(float);
  if (need <= kSmemLimit - 1024) {
    FPM_CUDA(cudaFuncSetAttribute(fpm::soft_topk_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    fpm::soft_topk_kernel<false><<<B, threads, need, st>>>(
        scores, ks, (const int64_t*)n1, (const int64_t*)n2, out, nullptr, R, C, max_iter, tau);
  } else {
    FPM_CHECK_ARG(workspace, "fpm_soft_topk: matrix exceeds shared memory, workspace required");
    fpm::soft_topk_kernel<true><<<B, threads, 0, st>>>(
        scores, ks, (const int64_t*)n1, (const int64_t*)n2, out, (float*)workspace, R, C, max_iter, tau);
  }
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
