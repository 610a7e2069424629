// Genuine / imposter classifier over the matched-similarity map, inference path.
//
// Replaces MatchClassifier.forward in eval mode (/root/reference/src/model/ngm.py:75-106, called at :451-455):
//   x = (s * perm_mat)[:, None]                                       [B, 1, H, W]
//   x = MaxPool2(BN(ReLU(Conv3x3(1 -> 16, pad 1)(x))))                [B, 16, H/2, W/2]
//   x = MaxPool2(BN(ReLU(Conv3x3(16 -> 32, pad 1)(x))))               [B, 32, H/4, W/4]
//   logit = Linear(32 -> 1)(mean over pixels)                         [B]
// Stock torch runs this as ~12 cuDNN / elementwise launches (9 % of the 256-pair step, every intermediate through HBM).
// Here: stage 1 fuses the product, conv, bias, ReLU, BatchNorm (running statistics) and the pool into one pass;
// stage 2 does the same for the second block with the weights in shared memory and 128 fp32 accumulators per thread
// (one pooled pixel x 32 channels x 2x2 pre-pool positions), and reduces the average pool to per-chunk partial
// sums; stage 3 adds the chunks in a fixed order and applies the linear layer.
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
                        float* __restrict__ p1, int H, int W, int H1, int W1, const float* __restrict__ w2,
                        float* __restrict__ wpack) {
  __shared__ float sw[kC1 * 9], sb[kC1], ssc[kC1], ssh[kC1];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    // second conv's weights [32][16][3][3] -> [16 ic][16 oc pairs][10 taps (9 used)][2] for stage 2 (one CTA, 20 KB):
    // the two output channels of a pair sit side by side, ready to be the 64-bit operand of a packed FMA
    for (int idx = tid; idx < kC1 * (kC2 / 2) * 20; idx += 256) {
      const int ic = idx / ((kC2 / 2) * 20), op = (idx / 20) % (kC2 / 2), k = (idx % 20) >> 1, u = idx & 1;
      wpack[idx] = k < 9 ? w2[((size_t)(2 * op + u) * kC1 + ic) * 9 + k] : 0.f;
    }
  }
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

// grid (cdiv(H2 * W2, 64), B), block 64: thread = one pooled output pixel (pixels taken in raster order, so that
// only the last chunk of a map has idle threads), 32 channels x 4 pre-pool positions = 128 accumulators held as 64
// packed pairs (two adjacent output channels).  The 4 x 4 input patch of every channel comes straight from P1 (L1 / L2
// resident; neighbouring threads share most of it); the weights are copied from the [ic][oc pair][10][2] repack stage 1
// left in the workspace and read back as 128-bit broadcasts (two taps of a channel pair each).  The 36 FMAs of an
// (input channel, output channel) are 36 FFMA2 per channel PAIR: half the issue slots of the scalar form
// (0.24 -> 0.14 ms at 256 pairs x 100 x 100).
constexpr int kChunk = 64;                   // pooled pixels per CTA
constexpr int kWpad = 20;                    // floats per (ic, oc pair): 9 taps x 2 channels, padded to five 128-bit loads

__global__ void __launch_bounds__(kChunk)
match_cls_stage2_kernel(const float* __restrict__ p1, const float* __restrict__ wpack, const float* __restrict__ b2,
                        BnParams bn2, float eps, float* __restrict__ partial, int H1, int W1, int H2, int W2) {
  extern __shared__ __align__(16) float sm2[];
  float* sw = sm2;                                     // [16 ic][16 oc pairs][kWpad]
  float* red = sw + kC1 * (kC2 / 2) * kWpad;           // [2 warps][32]
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int idx = tid; idx < kC1 * (kC2 / 2) * kWpad / 4; idx += kChunk)
    reinterpret_cast<float4*>(sw)[idx] = reinterpret_cast<const float4*>(wpack)[idx];
  __syncthreads();

  const int npix = H2 * W2;
  const int pix = blockIdx.x * kChunk + tid;
  const bool valid = pix < npix;
  const int pc = valid ? pix : npix - 1;               // idle threads shadow the last pixel
  const int y2 = pc / W2, x2 = pc - y2 * W2;
  const int y0 = 2 * y2 - 1, x0 = 2 * x2 - 1;          // P1 coordinates of patch element (0, 0)
  // rows / columns 1, 2 of the patch are always inside P1; 0 and 3 are the conv's zero padding at the borders
  const bool top = y0 >= 0, bottom = y0 + 3 < H1, left = x0 >= 0, right = x0 + 3 < W1;
  const float* base = p1 + (size_t)b * kC1 * H1 * W1 + (size_t)(y0 + 1) * W1 + (x0 + 1);

  f32x2 acc[kC2 / 2][4];
#pragma unroll
  for (int op = 0; op < kC2 / 2; ++op)
#pragma unroll
    for (int d = 0; d < 4; ++d) acc[op][d] = pk2(0.f, 0.f);

  for (int ic = 0; ic < kC1; ++ic) {
    const float* t = base + (size_t)ic * H1 * W1;      // patch element (1, 1)
    f32x2 p[4][4];                                     // patch values broadcast to both lanes
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const bool rok = (r == 0) ? top : (r == 3) ? bottom : true;
      const float* row = t + (r - 1) * W1;
      const float a0 = (rok && left) ? __ldg(row - 1) : 0.f;
      const float a1 = rok ? __ldg(row) : 0.f;
      const float a2 = rok ? __ldg(row + 1) : 0.f;
      const float a3 = (rok && right) ? __ldg(row + 2) : 0.f;
      p[r][0] = pk2(a0, a0); p[r][1] = pk2(a1, a1); p[r][2] = pk2(a2, a2); p[r][3] = pk2(a3, a3);
    }
    const float4* wv = reinterpret_cast<const float4*>(sw + ic * (kC2 / 2) * kWpad);
#pragma unroll
    for (int op = 0; op < kC2 / 2; ++op) {
      f32x2 w[10];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const float4 x = wv[op * 5 + q];
        w[2 * q] = pk2(x.x, x.y);
        w[2 * q + 1] = pk2(x.z, x.w);
      }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              acc[op][dy * 2 + dx] = fma2(p[dy + ky][dx + kx], w[ky * 3 + kx], acc[op][dy * 2 + dx]);
    }
  }

  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int op = 0; op < kC2 / 2; ++op) {
    float a[4][2];
#pragma unroll
    for (int d = 0; d < 4; ++d) upk2(acc[op][d], a[d][0], a[d][1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int oc = 2 * op + u;
      float scale, shift;
      bn_affine(bn2, oc, eps, scale, shift);
      const float bias = b2[oc];
      float best = -INFINITY;
#pragma unroll
      for (int d = 0; d < 4; ++d) best = fmaxf(best, fmaxf(a[d][u] + bias, 0.f) * scale + shift);
      float v = valid ? best : 0.f;
      v = warp_sum(v);
      if (lane == 0) red[warp * kC2 + oc] = v;
    }
  }
  __syncthreads();
  if (tid < kC2) partial[((size_t)b * gridDim.x + blockIdx.x) * kC2 + tid] = red[tid] + red[kC2 + tid];
}

// grid (cdiv(B, 4)), block 128: one warp per pair, lane = channel; fixed-order sum of the chunk partials, mean,
// linear layer.
__global__ void __launch_bounds__(128)
match_cls_stage3_kernel(const float* __restrict__ partial, const float* __restrict__ fcw, const float* __restrict__ fcb,
                        float* __restrict__ logits, int B, int chunks, float inv_pixels) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), oc = threadIdx.x & 31;
  if (b >= B) return;
  float sum = 0.f;
  for (int t = 0; t < chunks; ++t) sum += partial[((size_t)b * chunks + t) * kC2 + oc];
  const float out = warp_sum(sum * inv_pixels * fcw[oc]);
  if (oc == 0) logits[b] = out + fcb[0];
}

}  // namespace fpm

static inline long long match_cls_chunks(int H, int W) {
  const long long H2 = H / 4, W2 = W / 4;
  return (H2 * W2 + fpm::kChunk - 1) / fpm::kChunk;
}

extern "C" long long fpm_match_classifier_workspace_floats(int B, int H, int W) {
  const long long H1 = H / 2, W1 = W / 2;
  return (long long)B * fpm::kC1 * H1 * W1 + (long long)B * match_cls_chunks(H, W) * fpm::kC2 +
         (long long)fpm::kC1 * (fpm::kC2 / 2) * fpm::kWpad;
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
  const int chunks = (int)match_cls_chunks(H, W);
  float* p1 = workspace;
  float* partial = p1 + (size_t)B * fpm::kC1 * H1 * W1;
  float* wpack = partial + (size_t)B * chunks * fpm::kC2;
  fpm::match_cls_stage1_kernel<<<dim3(fpm_cdiv(W1, 32), fpm_cdiv(H1, 8), B), dim3(32, 8), 0, st>>>(
      s, perm, w1, b1, q1, eps, p1, H, W, H1, W1, w2, wpack);
  FPM_LAUNCH_CHECK();
  const size_t smem2 = sizeof(float) * ((size_t)fpm::kC1 * (fpm::kC2 / 2) * fpm::kWpad + 2 * fpm::kC2);
  fpm::match_cls_stage2_kernel<<<dim3(chunks, B), fpm::kChunk, smem2, st>>>(p1, wpack, b2, q2, eps, partial, H1, W1,
                                                                          H2, W2);
  FPM_LAUNCH_CHECK();
  fpm::match_cls_stage3_kernel<<<fpm_cdiv(B, 4), 128, 0, st>>>(partial, fcw, fcb, logits, B, chunks,
                                                                1.0f / ((float)H2 * (float)W2));
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
