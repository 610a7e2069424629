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
