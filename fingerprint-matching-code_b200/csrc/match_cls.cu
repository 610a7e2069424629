// Genuine / imposter classifier over the matched-similarity map, inference path.
//
// Replaces MatchClassifier.forward in eval mode (/root/reference/src/model/ngm.py:75-106, called at :451-455):
//   x = (s * perm_mat)[:, None]                                       [B, 1, H, W]
//   x = MaxPool2(BN(ReLU(Conv3x3(1 -> 16, pad 1)(x))))                [B, 16, H/2, W/2]
//   x = MaxPool2(BN(ReLU(Conv3x3(16 -> 32, pad 1)(x))))               [B, 32, H/4, W/4]
//   logit = Linear(32 -> 1)(mean over pixels)                         [B]
// Stock torch runs this as ~12 cuDNN / elementwise launches (9 % of the 256-pair step, every intermediate through HBM).
// Here: stage 1 fuses the product, conv, bias, ReLU, BatchNorm (running statistics) and the pool into one pass;
// stage 2 does the same for the second block with the 16-channel input tile and the weights in shared memory and
// 128 fp32 accumulators per thread (one pooled pixel x 32 channels x 2x2 pre-pool positions), and reduces the
// average pool to per-tile partial sums; stage 3 adds the tiles in a fixed order and applies the linear layer.
// Everything is fp32 FMA (cuDNN's default for this conv is TF32); results are deterministic.
// Training mode (batch statistics, autograd) stays on stock torch.
#include "common.cuh"

namespace fpm {

constexpr int kC1 = 16, kC2 = 32;

// BatchNorm2d in eval mode: weight, bias, running_mean, running_var  ->  y = x * scale + shift
struct BnParams { const float* weight; const float* bias; const float* mean; const float* var; };
__device__ __forceinline__ void bn_affine(const BnParams& bn, int c, float eps, float& scale, float& shift) {
  const float inv = 1.0f / sqrtf(bn.var[c] + eps);
  scale = bn.weight[c] * inv;
  shift = bn.bias[c] - bn.mean[c] * scale;
}

// grid (cdiv(W1, 32), cdiv(H1, 8), B), block (32, 8): one thread per pooled pixel, all 16 channels.
__global__ void __launch_bounds__(256)
match_cls_stage1_kernel(const float* __restrict__ s, const float* __restrict__ perm, const float* __restrict__ w1,
                        const float* __restrict__ b1, BnParams bn1, float eps,
                        float* __restrict__ p1, int H, int W, int H1, int W1) {
  __shared__ float sw[kC1 * 9], sb[kC1], ssc[kC1], ssh[kC1];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < kC1 * 9) sw[tid] = w1[tid];
  if (tid < kC1) {
    sb[tid] = b1[tid];
    bn_affine(bn1, tid, eps, ssc[tid], ssh[tid]);
  }
  __syncthreads();
  const int b = blockIdx.z, y1 = blockIdx.y * 8 + threadIdx.y, x1 = blockIdx.x * 32 + threadIdx.x;
  if (y1 >= H1 || x1 >= W1) return;
  const float* sb_ = s + (size_t)b * H * W;
  const float* pb_ = perm ? perm + (size_t)b * H * W : nullptr;
  float p[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = 2 * y1 - 1 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int x = 2 * x1 - 1 + c;
      float v = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        v = sb_[(size_t)y * W + x];
        if (pb_) v *= pb_[(size_t)y * W + x];
      }
      p[r][c] = v;
    }
  }
#pragma unroll
  for (int oc = 0; oc < kC1; ++oc) {
    float best = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float a = sb[oc];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) a = fmaf(p[dy + ky][dx + kx], sw[oc * 9 + ky * 3 + kx], a);
        a = fmaxf(a, 0.f) * ssc[oc] + ssh[oc];
        best = fmaxf(best, a);
      }
    p1[(((size_t)b * kC1 + oc) * H1 + y1) * W1 + x1] = best;
  }
}

// grid (cdiv(W2, 8), cdiv(H2, 8), B), block 64: thread = one pooled output pixel, 32 channels x 4 pre-pool positions.
constexpr int kT2 = 8;                       // pooled pixels per tile side
constexpr int kIn2 = 2 * kT2 + 2;            // input tile side (pre-pool 16 + halo)
constexpr int kWpad = 12;                    // 9 taps padded to 12 floats (three 128-bit loads)

__global__ void __launch_bounds__(kT2 * kT2)
match_cls_stage2_kernel(const float* __restrict__ p1, const float* __restrict__ w2, const float* __restrict__ b2,
                        BnParams bn2, float eps, float* __restrict__ partial, int H1, int W1,
                        int H2, int W2) {
  extern __shared__ __align__(16) float sm2[];
  float* tile = sm2;                                   // [16][kIn2][kIn2]
  float* sw = tile + kC1 * kIn2 * kIn2;                // [16 ic][32 oc][kWpad]
  float* red = sw + kC1 * kC2 * kWpad;                 // [2 warps][32]
  const int tid = threadIdx.x, b = blockIdx.z;
  const int ty0 = blockIdx.y * kT2, tx0 = blockIdx.x * kT2;
  const int gy0 = 2 * ty0 - 1, gx0 = 2 * tx0 - 1;      // P1 coordinates of tile element (0, 0)
  for (int idx = tid; idx < kC1 * kIn2 * kIn2; idx += kT2 * kT2) {
    const int ic = idx / (kIn2 * kIn2), r = (idx / kIn2) % kIn2, c = idx % kIn2;
    const int y = gy0 + r, x = gx0 + c;
    tile[idx] = (y >= 0 && y < H1 && x >= 0 && x < W1) ? p1[(((size_t)b * kC1 + ic) * H1 + y) * W1 + x] : 0.f;
  }
  for (int idx = tid; idx < kC1 * kC2 * kWpad; idx += kT2 * kT2) {
    const int ic = idx / (kC2 * kWpad), oc = (idx / kWpad) % kC2, k = idx % kWpad;
    sw[idx] = k < 9 ? w2[((size_t)oc * kC1 + ic) * 9 + k] : 0.f;
  }
  __syncthreads();

  const int ty = tid / kT2, tx = tid % kT2;
  float acc[kC2][4];
#pragma unroll
  for (int oc = 0; oc < kC2; ++oc)
#pragma unroll
    for (int d = 0; d < 4; ++d) acc[oc][d] = 0.f;

  for (int ic = 0; ic < kC1; ++ic) {
    float p[4][4];
    const float* t = tile + (ic * kIn2 + 2 * ty) * kIn2 + 2 * tx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 a = *reinterpret_cast<const float2*>(t + r * kIn2);
      const float2 c = *reinterpret_cast<const float2*>(t + r * kIn2 + 2);
      p[r][0] = a.x; p[r][1] = a.y; p[r][2] = c.x; p[r][3] = c.y;
    }
    const float4* wv = reinterpret_cast<const float4*>(sw + ic * kC2 * kWpad);
#pragma unroll
    for (int oc = 0; oc < kC2; ++oc) {
      const float4 w0 = wv[oc * 3], w1 = wv[oc * 3 + 1], w2_ = wv[oc * 3 + 2];
      const float w[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2_.x};
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              acc[oc][dy * 2 + dx] = fmaf(p[dy + ky][dx + kx], w[ky * 3 + kx], acc[oc][dy * 2 + dx]);
    }
  }

  const bool valid = (ty0 + ty) < H2 && (tx0 + tx) < W2;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int oc = 0; oc < kC2; ++oc) {
    float scale, shift;
    bn_affine(bn2, oc, eps, scale, shift);
    const float bias = b2[oc];
    float best = -INFINITY;
#pragma unroll
    for (int d = 0; d < 4; ++d) best = fmaxf(best, fmaxf(acc[oc][d] + bias, 0.f) * scale + shift);
    float v = valid ? best : 0.f;
    v = warp_sum(v);
    if (lane == 0) red[warp * kC2 + oc] = v;
  }
  __syncthreads();
  if (tid < kC2) {
    const int tiles = gridDim.x * gridDim.y, tix = blockIdx.y * gridDim.x + blockIdx.x;
    partial[((size_t)b * tiles + tix) * kC2 + tid] = red[tid] + red[kC2 + tid];
  }
}

// grid (cdiv(B, 128)), block 128: fixed-order sum of the tile partials, mean, linear layer.
__global__ void __launch_bounds__(128)
match_cls_stage3_kernel(const float* __restrict__ partial, const float* __restrict__ fcw, const float* __restrict__ fcb,
                        float* __restrict__ logits, int B, int tiles, float inv_pixels) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float out = 0.f;
  for (int oc = 0; oc < kC2; ++oc) {
    float sum = 0.f;
    for (int t = 0; t < tiles; ++t) sum += partial[((size_t)b * tiles + t) * kC2 + oc];
    out = fmaf(sum * inv_pixels, fcw[oc], out);
  }
  logits[b] = out + fcb[0];
}

}  // namespace fpm

extern "C" long long fpm_match_classifier_workspace_floats(int B, int H, int W) {
  const long long H1 = H / 2, W1 = W / 2, H2 = H1 / 2, W2 = W1 / 2;
  const long long tiles = (long long)fpm_cdiv(W2 > 0 ? W2 : 1, fpm::kT2) * fpm_cdiv(H2 > 0 ? H2 : 1, fpm::kT2);
  return (long long)B * fpm::kC1 * H1 * W1 + (long long)B * tiles * fpm::kC2;
}

extern "C" int fpm_match_classifier(const float* s, const float* perm, const float* w1, const float* b1,
                                    const float* const* bn1, const float* w2, const float* b2,
                                    const float* const* bn2, const float* fcw, const float* fcb, float eps,
                                    float* workspace, float* logits, int B, int H, int W, void* stream) {
  FPM_CHECK_ARG(s && w1 && b1 && bn1 && w2 && b2 && bn2 && fcw && fcb && workspace && logits,
                "fpm_match_classifier: null tensor");
  for (int i = 0; i < 4; ++i) FPM_CHECK_ARG(bn1[i] && bn2[i], "fpm_match_classifier: null BatchNorm tensor");
  const fpm::BnParams q1{bn1[0], bn1[1], bn1[2], bn1[3]}, q2{bn2[0], bn2[1], bn2[2], bn2[3]};
  FPM_CHECK_ARG(B >= 0 && H >= 4 && W >= 4 && B <= 65535, "fpm_match_classifier: the map must be at least 4 x 4");
  if (B == 0) return FPM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int H1 = H / 2, W1 = W / 2, H2 = H1 / 2, W2 = W1 / 2;
  float* p1 = workspace;
  float* partial = workspace + (size_t)B * fpm::kC1 * H1 * W1;
  fpm::match_cls_stage1_kernel<<<dim3(fpm_cdiv(W1, 32), fpm_cdiv(H1, 8), B), dim3(32, 8), 0, st>>>(
      s, perm, w1, b1, q1, eps, p1, H, W, H1, W1);
  FPM_LAUNCH_CHECK();
  const dim3 g2(fpm_cdiv(W2, fpm::kT2), fpm_cdiv(H2, fpm::kT2), B);
  const size_t smem2 = sizeof(float) * ((size_t)fpm::kC1 * fpm::kIn2 * fpm::kIn2 + (size_t)fpm::kC1 * fpm::kC2 * fpm::kWpad +
                                        2 * fpm::kC2);
  FPM_CUDA(cudaFuncSetAttribute(fpm::match_cls_stage2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  fpm::match_cls_stage2_kernel<<<g2, fpm::kT2 * fpm::kT2, smem2, st>>>(p1, w2, b2, q2, eps, partial, H1, W1, H2, W2);
  FPM_LAUNCH_CHECK();
  fpm::match_cls_stage3_kernel<<<fpm_cdiv(B, 128), 128, 0, st>>>(partial, fcw, fcb, logits, B, (int)(g2.x * g2.y),
                                                                1.0f / ((float)H2 * (float)W2));
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
