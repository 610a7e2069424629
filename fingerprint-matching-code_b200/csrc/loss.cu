// Loss and metric reductions that follow the forward in every train / validation step.
//
// fpm_permutation_loss replaces PermutationLoss.forward (/root/reference/src/loss_func.py:26-59): a python loop over
// pairs calling F.binary_cross_entropy(reduction='sum') on each pair's valid block, divided by sum(n1).  One CTA per
// pair reduces its block; the backward is the element-wise derivative torch uses (binary_cross_entropy_backward:
// (p - y) / max(p (1 - p), 1e-12)), zero outside the valid blocks.
// fpm_matching_stats replaces the per-pair loops of matching_recall / matching_precision
// (/root/reference/src/evaluation_metric.py:58-131): per pair sum(pred * gt), sum(gt), sum(pred) over rows < ns[b].
#include "common.cuh"

namespace fpm {

__global__ void __launch_bounds__(256)
perm_loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const int64_t* __restrict__ n1,
                     const int64_t* __restrict__ n2, float* __restrict__ pair_sum, int R, int C) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int r = min((int)n1[b], R), c = min((int)n2[b], C);
  const float* p = pred + (size_t)b * R * C;
  const float* y = gt + (size_t)b * R * C;
  float acc = 0.f;
  for (int idx = threadIdx.x; idx < r * c; idx += blockDim.x) {
    const int i = idx / c, j = idx - i * c;
    const float pv = p[(size_t)i * C + j], yv = y[(size_t)i * C + j];
    const float lp = fmaxf(logf(pv), -100.f), lq = fmaxf(log1pf(-pv), -100.f);     // torch clamps the logs at -100
    acc -= yv * lp + (1.f - yv) * lq;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) pair_sum[b] = acc;
}

__global__ void __launch_bounds__(256)
perm_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const int64_t* __restrict__ n1,
                     const int64_t* __restrict__ n2, const float* __restrict__ gscale, float* __restrict__ grad,
                     int R, int C) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * C) return;
  const int i = idx / C, j = idx - i * C;
  float g = 0.f;
  if (i < (int)n1[b] && j < (int)n2[b]) {
    const float pv = pred[(size_t)b * R * C + idx], yv = gt[(size_t)b * R * C + idx];
    g = gscale[0] * (pv - yv) / fmaxf((1.f - pv) * pv, 1e-12f);
  }
  grad[(size_t)b * R * C + idx] = g;
}

__global__ void __launch_bounds__(256)
matching_stats_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const int64_t* __restrict__ ns,
                      float* __restrict__ stats, int R, int C) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int r = min((int)ns[b], R);
  const float* p = pred + (size_t)b * R * C;
  const float* y = gt + (size_t)b * R * C;
  float hit = 0.f, ngt = 0.f, npred = 0.f;
  for (int idx = threadIdx.x; idx < r * C; idx += blockDim.x) {
    const float pv = p[idx], yv = y[idx];
    hit += pv * yv; ngt += yv; npred += pv;
  }
  hit = block_sum(hit, red);
  ngt = block_sum(ngt, red);
  npred = block_sum(npred, red);
  if (threadIdx.x == 0) { stats[b * 3] = hit; stats[b * 3 + 1] = ngt; stats[b * 3 + 2] = npred; }
}

// The scalar tail of Net.forward in eval mode (ngm.py:456-469): cls_prob = sigmoid(logits), the BCE-with-logits mean
// against the pair labels, the k-regression MSE (x k_factor) and L1 error against gt_ks = sum(gt_perm_mat) and
// min(n1, n2).  Stock torch spends ~17 launches of 1-CTA kernels on these 4 x B numbers.  CTA per pair (it sums its
// ground-truth matrix); the last CTA to finish adds the per-pair terms in a fixed order (deterministic) and rearms the
// ticket counter.  scalars: [0] cls_loss (0 without labels), [1] ks_loss, [2] ks_error.
__global__ void __launch_bounds__(256)
head_losses_kernel(const float* __restrict__ logits, const float* __restrict__ label, const float* __restrict__ ks,
                   const float* __restrict__ gt_perm, const int64_t* __restrict__ n1, const int64_t* __restrict__ n2,
                   float k_factor, float* __restrict__ cls_prob, float* __restrict__ terms, int* __restrict__ counter,
                   float* __restrict__ scalars, int B, int RC) {
  __shared__ float red[32];
  __shared__ int last;
  const int b = blockIdx.x;
  const float* g = gt_perm + (size_t)b * RC;
  float cnt = 0.f;
  for (int i = threadIdx.x; i < RC; i += blockDim.x) cnt += g[i];
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) {
    const float x = logits[b];
    cls_prob[b] = 1.0f / (1.0f + expf(-x));
    float bce = 0.f;
    if (label) {                                     // (1 - y) x + max(-x, 0) + log(exp(-max) + exp(-x - max))
      const float mv = fmaxf(-x, 0.f);
      bce = (1.0f - label[b]) * x + mv + logf(expf(-mv) + expf(-x - mv));
    }
    float mse = 0.f, l1 = 0.f;
    if (ks) {
      const float mp = (float)min(n1[b], n2[b]);
      const float d = ks[b] - cnt / mp;
      mse = d * d;
      l1 = fabsf(ks[b] * mp - cnt);
    }
    terms[b * 3] = bce; terms[b * 3 + 1] = mse; terms[b * 3 + 2] = l1;
    __threadfence();
    last = atomicAdd(counter, 1) == B - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float acc[3] = {0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const volatile float* t = terms + (size_t)i * 3;
    acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) acc[k] = block_sum(acc[k], red);
  if (threadIdx.x == 0) {
    scalars[0] = acc[0] / (float)B;
    scalars[1] = acc[1] / (float)B * k_factor;
    scalars[2] = acc[2] / (float)B;
    *counter = 0;
  }
}

}  // namespace fpm

extern "C" int fpm_permutation_loss(const float* pred, const float* gt, const long long* n1, const long long* n2,
                                    float* pair_sum, int B, int R, int C, void* stream) {
  FPM_CHECK_ARG(pred && gt && n1 && n2 && pair_sum, "fpm_permutation_loss: null tensor");
  if (B == 0) return FPM_OK;
  fpm::perm_loss_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(pred, gt, (const int64_t*)n1, (const int64_t*)n2,
                                                                 pair_sum, R, C);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_permutation_loss_bwd(const float* pred, const float* gt, const long long* n1, const long long* n2,
                                        const float* gscale, float* grad, int B, int R, int C, void* stream) {
  FPM_CHECK_ARG(pred && gt && n1 && n2 && gscale && grad, "fpm_permutation_loss_bwd: null tensor");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_permutation_loss_bwd: batch too large");
  fpm::perm_loss_bwd_kernel<<<dim3(fpm_cdiv((long long)R * C, 256), B), 256, 0, (cudaStream_t)stream>>>(
      pred, gt, (const int64_t*)n1, (const int64_t*)n2, gscale, grad, R, C);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_matching_stats(const float* pred, const float* gt, const long long* ns, float* stats, int B, int R,
                                  int C, void* stream) {
  FPM_CHECK_ARG(pred && gt && ns && stats, "fpm_matching_stats: null tensor");
  if (B == 0) return FPM_OK;
  fpm::matching_stats_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(pred, gt, (const int64_t*)ns, stats, R, C);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// workspace: 3 * B floats + one int (zero before the first use; the kernel rearms it), 16-byte aligned.
extern "C" int fpm_head_losses(const float* logits, const float* label, const float* ks, const float* gt_perm,
                               const long long* n1, const long long* n2, float k_factor, float* cls_prob,
                               float* workspace, float* scalars, int B, int R, int C, void* stream) {
  FPM_CHECK_ARG(logits && gt_perm && n1 && n2 && cls_prob && workspace && scalars, "fpm_head_losses: null tensor");
  FPM_CHECK_ARG(B > 0 && R > 0 && C > 0, "fpm_head_losses: bad sizes");
  fpm::head_losses_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logits, label, ks, gt_perm, (const int64_t*)n1,
                                                              (const int64_t*)n2, k_factor, cls_prob, workspace,
                                                              (int*)(workspace + (size_t)3 * B), scalars, B, R * C);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
