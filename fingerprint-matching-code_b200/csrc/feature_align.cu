// Keypoint feature gather: bilinear feature_align over the backbone feature maps.
//
// Replaces /root/reference/utils/feature_align.py:5-125 (a python loop of ~40 tiny ops per keypoint)
// and, in the fused head path, normalize_over_channels + concat_features + torch.cat of
// /root/reference/src/model/ngm.py:241-251.  HBM-bound: per pair it reads the two feature maps once
// and writes [n, 768] node features once (1.56 MB/pair at n = 100).
//
//   fpm_feature_align        drop-in for utils.feature_align.feature_align ([B,C,H,W] -> [B,C,nmax]),
//                            bit-exact: products and sums are rounded separately, left to right,
//                            exactly as  Ia*wa + Ib*wb + Ic*wc + Id*wd  evaluates in torch.
//   fpm_fmap_prep            NCHW -> channels-last with the per-position L2 norm divided out, so the
//                            gather below reads 128-bit vectors along C.
//   fpm_node_features        one warp per keypoint: 4 taps x (256 + 512) channels, float4 loads,
//                            writes row [768] of the concatenated node-feature matrix.
#include "common.cuh"

namespace fpm {

struct Taps {
  int y0, y1, x0, x1;       // clamped tap coordinates (fetch positions)
  float wa, wb, wc, wd;     // weights after the post-fetch edge adjustment
};

// feature_align.py:55-62 (coordinate transform with the (W,H)/(Hf,Wf) mix-up kept) and :79-118.
__device__ __forceinline__ Taps make_taps(float px, float py, float ori_w, float ori_h, int Hf, int Wf,
                                          bool feat_coords = false) {
  // ori_size = (ori_w, ori_h); feat_size = (Hf, Wf)  [sic]; step = ori / feat
  const float f0 = (float)Hf, f1 = (float)Wf;
  const float step0 = __fdiv_rn(ori_w, f0), step1 = __fdiv_rn(ori_h, f1);
  float x = __fmul_rn(__fdiv_rn(__fsub_rn(px, __fdiv_rn(step0, 2.f)), ori_w), f0);
  float y = __fmul_rn(__fdiv_rn(__fsub_rn(py, __fdiv_rn(step1, 2.f)), ori_h), f1);
  if (feat_coords) { x = px; y = py; }     // bilinear_interpolate(im, x, y): already feature-space
  float x0 = floorf(x), x1 = x0 + 1.f, y0 = floorf(y), y1 = y0 + 1.f;
  x0 = fminf(fmaxf(x0, 0.f), (float)(Wf - 1)); x1 = fminf(fmaxf(x1, 0.f), (float)(Wf - 1));
  y0 = fminf(fmaxf(y0, 0.f), (float)(Hf - 1)); y1 = fminf(fmaxf(y1, 0.f), (float)(Hf - 1));
  Taps t;
  t.x0 = (int)x0; t.x1 = (int)x1; t.y0 = (int)y0; t.y1 = (int)y1;
  int ax0 = t.x0, ax1 = t.x1, ay0 = t.y0, ay1 = t.y1;
  if (ax0 == ax1) { if (ax0 == 0) ax0 -= 1; else ax1 += 1; }
  if (ay0 == ay1) { if (ay0 == 0) ay0 -= 1; else ay1 += 1; }
  const float fx0 = (float)ax0, fx1 = (float)ax1, fy0 = (float)ay0, fy1 = (float)ay1;
  t.wa = __fmul_rn(__fsub_rn(fx1, x), __fsub_rn(fy1, y));
  t.wb = __fmul_rn(__fsub_rn(fx1, x), __fsub_rn(y, fy0));
  t.wc = __fmul_rn(__fsub_rn(x, fx0), __fsub_rn(fy1, y));
  t.wd = __fmul_rn(__fsub_rn(x, fx0), __fsub_rn(y, fy0));
  return t;
}

__device__ __forceinline__ float blend(float a, float b, float c, float d, const Taps& t) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, t.wa), __fmul_rn(b, t.wb)), __fmul_rn(c, t.wc)),
                   __fmul_rn(d, t.wd));
}

// Drop-in op.  grid (ceil(nmax/32), ceil(C/8), B), block (32, 8): lanes over points (coalesced
// stores along n), y over channels.
__global__ void feature_align_kernel(const float* __restrict__ fmap, const float* __restrict__ P,
                                     const int64_t* __restrict__ ns, float* __restrict__ out, int C,
                                     int Hf, int Wf, int nmax, float ori_w, float ori_h, int feat_coords) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int c = blockIdx.y * 8 + threadIdx.y;
  if (i >= nmax || c >= C) return;
  const int n = (int)ns[b];
  float v = 0.f;
  if (i < n) {
    const Taps t = make_taps(P[((size_t)b * nmax + i) * 2], P[((size_t)b * nmax + i) * 2 + 1], ori_w,
                             ori_h, Hf, Wf, feat_coords != 0);
    const float* im = fmap + ((size_t)b * C + c) * Hf * Wf;
    v = blend(im[t.y0 * Wf + t.x0], im[t.y1 * Wf + t.x0], im[t.y0 * Wf + t.x1], im[t.y1 * Wf + t.x1], t);
  }
  out[((size_t)b * C + c) * nmax + i] = v;
}

// NCHW raw -> NHWC divided by the channel L2 norm (normalize_over_channels, ngm.py:65-67).
// One CTA per (kTp-position tile, image); 256 threads; the tile [C][kTp + 1] is staged in shared memory: a warp load
// covers 32 / kTp channels x kTp consecutive positions (eight loads in flight per thread), the stores walk the
// channels of one position (coalesced rows of the NHWC map).  kTp = 16 for the small maps (8 x 10 positions: a
// 32-wide tile left 3 CTAs per image, the last one half empty, and 68 KB of shared memory per CTA at 512 channels).
template <int kTp, int kCpt>                   // kCpt = C / 256 when that is exact (channels per thread), else 0
__global__ void __launch_bounds__(256)
fmap_prep_kernel(const float* __restrict__ fmap, float* __restrict__ out, int C, int HW) {
  extern __shared__ float tile[];              // [C][kTp + 1]
  constexpr int kCpw = 32 / kTp;               // channels per warp load
  constexpr int kLd = kTp + 1;
  __shared__ float part[8 * kCpw][kLd];
  __shared__ float norm[kTp], rnorm[kTp];
  const int b = blockIdx.y, p0 = blockIdx.x * kTp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lp = lane % kTp, csub = lane / kTp;
  const int p = p0 + lp;
  const float* src = fmap + (size_t)b * C * HW;
  float ss = 0.f;
#pragma unroll 8
  for (int c = warp * kCpw + csub; c < C; c += 8 * kCpw) {
    const float v = (p < HW) ? src[(size_t)c * HW + p] : 0.f;
    tile[c * kLd + lp] = v;
    ss = fmaf(v, v, ss);
  }
  part[warp * kCpw + csub][lp] = ss;
  __syncthreads();
  if (threadIdx.x < kTp) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8 * kCpw; ++w) t += part[w][threadIdx.x];
    const float nrm = sqrtf(t);
    norm[threadIdx.x] = nrm;
    rnorm[threadIdx.x] = 1.0f / nrm;
  }
  __syncthreads();
  float* dst = out + ((size_t)b * HW + p0) * C;
  const int npos = min(kTp, HW - p0);
  // v / n from the position's correctly rounded reciprocal and one FMA correction (Markstein): three FMA-pipe
  // instructions per element instead of a full IEEE division sequence in every thread.  With the channel count a
  // multiple of 256 a thread owns fixed channels and walks the positions, four in flight (the generic loop spends
  // ~80 instructions of control per position for one or two elements).
  if (kCpt > 0) {
    float* d0 = dst + threadIdx.x;
#pragma unroll 4
    for (int q = 0; q < npos; ++q) {
      const float nq = norm[q], rq = rnorm[q];
#pragma unroll
      for (int u = 0; u < kCpt; ++u) {
        const float v = tile[(threadIdx.x + 256 * u) * kLd + q];
        const float d = v * rq;
        d0[(size_t)q * C + 256 * u] = fmaf(fmaf(-d, nq, v), rq, d);
      }
    }
  } else {
    for (int q = 0; q < npos; ++q) {
      const float nq = norm[q], rq = rnorm[q];
      for (int c = threadIdx.x; c < C; c += 256) {
        const float v = tile[c * kLd + q];
        const float d = v * rq;
        dst[(size_t)q * C + c] = fmaf(fmaf(-d, nq, v), rq, d);
      }
    }
  }
}

// global[b, c] = max over positions of the raw map (AdaptiveMaxPool2d(1, 1), feature_extractor.py:54).  The (b, c)
// rows are contiguous, so kG lanes share a row with 128-bit loads (kG = 8 for the 8 x 10 maps: a warp then reads four
// whole rows = 1280 contiguous bytes with three independent loads per lane); HW % 4 != 0 takes the scalar
// warp-per-row loop.  (First version: always one warp per row with scalar loads - 131 k warps of 320 bytes each,
// 1.4 TB/s.)
template <int kG>
__global__ void __launch_bounds__(256)
global_max_kernel(const float* __restrict__ fmap, float* __restrict__ out, int BC, int HW, int out_stride,
                  int out_offset, int C, int vec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = t / kG, sub = t % kG;
  float m = kNegInf;
  if (w < BC) {
    if (vec) {
      const float4* src = (const float4*)(fmap + (size_t)w * HW);
      for (int i = sub; i < HW / 4; i += kG) {
        const float4 v = src[i];
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
    } else {
      const float* src = fmap + (size_t)w * HW;
      for (int i = sub; i < HW; i += kG) m = fmaxf(m, src[i]);
    }
  }
#pragma unroll
  for (int o = kG / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (w < BC && sub == 0) out[(size_t)(w / C) * out_stride + out_offset + (w % C)] = m;
}

// Fused gather: X[ptr[b] + i, 0:C1] = align(nodes_nhwc), X[.., C1:C1+C2] = align(edges_nhwc).
// One warp per keypoint, float4 lanes along channels.
__global__ void __launch_bounds__(256)
node_features_kernel(const float* __restrict__ nodes, const float* __restrict__ edges,
                     const float* __restrict__ P, const int64_t* __restrict__ ns,
                     const int64_t* __restrict__ ptr, float* __restrict__ X, int B, int nmax, int C1,
                     int H1, int W1, int C2, int H2, int W2, float ori_w, float ori_h) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * nmax) return;
  const int b = gw / nmax, i = gw - b * nmax;
  if (i >= (int)ns[b]) return;
  const float px = P[((size_t)b * nmax + i) * 2], py = P[((size_t)b * nmax + i) * 2 + 1];
  float* xrow = X + ((size_t)ptr[b] + i) * (C1 + C2);
  {
    const Taps t = make_taps(px, py, ori_w, ori_h, H1, W1);
    const float* base = nodes + (size_t)b * H1 * W1 * C1;
    const float4* a = (const float4*)(base + (size_t)(t.y0 * W1 + t.x0) * C1);
    const float4* bq = (const float4*)(base + (size_t)(t.y1 * W1 + t.x0) * C1);
    const float4* c = (const float4*)(base + (size_t)(t.y0 * W1 + t.x1) * C1);
    const float4* d = (const float4*)(base + (size_t)(t.y1 * W1 + t.x1) * C1);
    float4* o = (float4*)xrow;
    for (int v = lane; v < C1 / 4; v += 32) {
      const float4 A = a[v], Bv = bq[v], Cv = c[v], Dv = d[v];
      float4 r;
      r.x = blend(A.x, Bv.x, Cv.x, Dv.x, t); r.y = blend(A.y, Bv.y, Cv.y, Dv.y, t);
      r.z = blend(A.z, Bv.z, Cv.z, Dv.z, t); r.w = blend(A.w, Bv.w, Cv.w, Dv.w, t);
      o[v] = r;
    }
  }
  {
    const Taps t = make_taps(px, py, ori_w, ori_h, H2, W2);
    const float* base = edges + (size_t)b * H2 * W2 * C2;
    const float4* a = (const float4*)(base + (size_t)(t.y0 * W2 + t.x0) * C2);
    const float4* bq = (const float4*)(base + (size_t)(t.y1 * W2 + t.x0) * C2);
    const float4* c = (const float4*)(base + (size_t)(t.y0 * W2 + t.x1) * C2);
    const float4* d = (const float4*)(base + (size_t)(t.y1 * W2 + t.x1) * C2);
    float4* o = (float4*)(xrow + C1);
    for (int v = lane; v < C2 / 4; v += 32) {
      const float4 A = a[v], Bv = bq[v], Cv = c[v], Dv = d[v];
      float4 r;
      r.x = blend(A.x, Bv.x, Cv.x, Dv.x, t); r.y = blend(A.y, Bv.y, Cv.y, Dv.y, t);
      r.z = blend(A.z, Bv.z, Cv.z, Dv.z, t); r.w = blend(A.w, Bv.w, Cv.w, Dv.w, t);
      o[v] = r;
    }
  }
}

// coeff[b, :] = tanh(A * normalize([g_src[b]; g_tgt[b]]) + a)   (ngm.py:262-268, affinity_layer.py:13)
// grid (B, cdiv(OUT, 64)): one CTA per (pair, 64 output rows of A) - the first version gave a pair's 768 rows to one CTA
// and took 0.2 ms whatever the batch; the 2*G-vector sits in shared memory; one warp per output row.
constexpr int kCoeffRows = 64;
__global__ void __launch_bounds__(256)
affinity_coeff_kernel(const float* __restrict__ gcat, const float* __restrict__ W,
                      const float* __restrict__ bias, float* __restrict__ coeff, int IN, int OUT) {
  extern __shared__ float g[];
  __shared__ float red[32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ss = 0.f;
  for (int i = threadIdx.x; i < IN; i += blockDim.x) {
    const float v = gcat[(size_t)b * IN + i];
    g[i] = v;
    ss = fmaf(v, v, ss);
  }
  ss = block_sum(ss, red);
  const float nrm = sqrtf(ss);
  __syncthreads();
  for (int i = threadIdx.x; i < IN; i += blockDim.x) g[i] = g[i] / nrm;
  __syncthreads();
  const int o_end = min(OUT, (int)(blockIdx.y + 1) * kCoeffRows);
  for (int o = blockIdx.y * kCoeffRows + warp; o < o_end; o += (blockDim.x >> 5)) {
    const float* w = W + (size_t)o * IN;
    float acc = 0.f;
    for (int i = lane; i < IN; i += 32) acc = fmaf(w[i], g[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) coeff[(size_t)b * OUT + o] = tanhf(acc + bias[o]);
  }
}


// ------------------------------------------------------------------------------------------
// Backward of the fused gather + normalisation (training, BASELINE.json config 3).  The reference gets these
// gradients from autograd through `out[:, i] = bilinear_interpolate(...)` (feature_align.py:62) and
// normalize_over_channels (ngm.py:65-67, 241-243).
// ------------------------------------------------------------------------------------------
// dprep[b, pos, c] = sum over the pair's keypoints and their 4 taps of w_tap * dX[ptr[b]+i, coff + c].
// One CTA per (pair, 64-channel chunk): the chunk's [HW][64] accumulator lives in shared memory, keypoints are
// walked in index order and every position is owned by one thread group -> deterministic, no atomics.
__global__ void __launch_bounds__(256)
node_features_bwd_kernel(const float* __restrict__ dX, const float* __restrict__ P,
                         const int64_t* __restrict__ ns, const int64_t* __restrict__ ptr,
                         float* __restrict__ dprep, int nmax, int C, int H, int W, int coff, int ctot,
                         float ori_w, float ori_h) {
  extern __shared__ float acc[];                 // [H*W][64]
  const int b = blockIdx.y, c0 = blockIdx.x * 64;
  const int c = threadIdx.x & 63, grp = threadIdx.x >> 6;     // 4 groups, position % 4 == grp
  const int HW = H * W;
  for (int i = threadIdx.x; i < HW * 64; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int n = (int)ns[b];
  const float* dx = dX + (size_t)ptr[b] * ctot + coff + c0 + c;
  for (int i = 0; i < n; ++i) {
    const Taps t = make_taps(P[((size_t)b * nmax + i) * 2], P[((size_t)b * nmax + i) * 2 + 1], ori_w, ori_h, H, W);
    const float g = dx[(size_t)i * ctot];
    const int pa = t.y0 * W + t.x0, pb = t.y1 * W + t.x0, pc = t.y0 * W + t.x1, pd = t.y1 * W + t.x1;
    if ((pa & 3) == grp) acc[pa * 64 + c] = fmaf(t.wa, g, acc[pa * 64 + c]);
    if ((pb & 3) == grp) acc[pb * 64 + c] = fmaf(t.wb, g, acc[pb * 64 + c]);
    if ((pc & 3) == grp) acc[pc * 64 + c] = fmaf(t.wc, g, acc[pc * 64 + c]);
    if ((pd & 3) == grp) acc[pd * 64 + c] = fmaf(t.wd, g, acc[pd * 64 + c]);
  }
  __syncthreads();
  float* dst = dprep + (size_t)b * HW * C + c0;
  for (int i = threadIdx.x; i < HW * 64; i += blockDim.x) dst[(size_t)(i >> 6) * C + (i & 63)] = acc[i];
}

// y = x / ||x||_c  ->  dx = (dy - y * <y, dy>) / ||x||.  x: raw NCHW map, dy: NHWC (gradient of the prepared
// map), dx: NCHW.  Same tiling as fmap_prep_kernel.
__global__ void __launch_bounds__(256)
fmap_prep_bwd_kernel(const float* __restrict__ fmap, const float* __restrict__ dy, float* __restrict__ dxo,
                     int C, int HW) {
  extern __shared__ float tile[];              // [C][33]
  __shared__ float part[8][33];
  __shared__ float norm[32];
  const int b = blockIdx.y, p0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = p0 + lane;
  const float* src = fmap + (size_t)b * C * HW;
  float ss = 0.f;
  for (int c = warp; c < C; c += 8) {
    const float v = (p < HW) ? src[(size_t)c * HW + p] : 0.f;
    tile[c * 33 + lane] = v;
    ss = fmaf(v, v, ss);
  }
  part[warp][lane] = ss;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    norm[lane] = sqrtf(t);
  }
  __syncthreads();
  const int npos = min(32, HW - p0);
  for (int q = warp; q < npos; q += 8) {       // warp per position, lanes along channels
    const float nq = norm[q];
    const float* g = dy + ((size_t)b * HW + p0 + q) * C;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot = fmaf(tile[c * 33 + q] / nq, g[c], dot);
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
      const float y = tile[c * 33 + q] / nq;
      tile[c * 33 + q] = (g[c] - y * dot) / nq;
    }
  }
  __syncthreads();
  float* dst = dxo + (size_t)b * C * HW;
  for (int c = warp; c < C; c += 8)
    if (p < HW) dst[(size_t)c * HW + p] = tile[c * 33 + lane];
}

}  // namespace fpm

extern "C" int fpm_feature_align(const float* fmap, const float* P, const long long* ns, float* out,
                                 int B, int C, int Hf, int Wf, int nmax, float ori_w, float ori_h,
                                 int feat_coords, void* stream) {
  FPM_CHECK_ARG(fmap && P && ns && out, "fpm_feature_align: null tensor");
  FPM_CHECK_ARG(B >= 0 && C > 0 && Hf > 0 && Wf > 0 && nmax >= 0, "fpm_feature_align: bad sizes");
  if (B == 0 || nmax == 0) return FPM_OK;
  dim3 grid(fpm_cdiv(nmax, 32), fpm_cdiv(C, 8), B), block(32, 8);
  FPM_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "fpm_feature_align: batch or channel count too large");
  fpm::feature_align_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
      fmap, P, (const int64_t*)ns, out, C, Hf, Wf, nmax, ori_w, ori_h, feat_coords);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_fmap_prep(const float* fmap, float* out_nhwc, int B, int C, int Hf, int Wf,
                             void* stream) {
  FPM_CHECK_ARG(fmap && out_nhwc, "fpm_fmap_prep: null tensor");
  FPM_CHECK_ARG(B >= 0 && C > 0 && Hf > 0 && Wf > 0, "fpm_fmap_prep: bad sizes");
  if (B == 0) return FPM_OK;
  const int HW = Hf * Wf;
  const bool narrow = HW <= 128;               // small maps: 16-position tiles
  const size_t smem = (size_t)C * (narrow ? 17 : 33) * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_fmap_prep: channel count too large");
  FPM_CHECK_ARG(B <= 65535, "fpm_fmap_prep: batch too large");
  const int cpt = (C % 256 == 0 && C <= 512) ? C / 256 : 0;
#define FPM_FMAP_PREP(TP, CPT)                                                                                     \
  do {                                                                                                             \
    FPM_CUDA(cudaFuncSetAttribute(fpm::fmap_prep_kernel<TP, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)smem));                                                                     \
    fpm::fmap_prep_kernel<TP, CPT><<<dim3(fpm_cdiv(HW, TP), B), 256, smem, (cudaStream_t)stream>>>(fmap, out_nhwc, \
                                                                                                   C, HW);         \
  } while (0)
  if (narrow) {
    if (cpt == 1) FPM_FMAP_PREP(16, 1); else if (cpt == 2) FPM_FMAP_PREP(16, 2); else FPM_FMAP_PREP(16, 0);
  } else {
    if (cpt == 1) FPM_FMAP_PREP(32, 1); else if (cpt == 2) FPM_FMAP_PREP(32, 2); else FPM_FMAP_PREP(32, 0);
  }
#undef FPM_FMAP_PREP
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_global_max(const float* fmap, float* out, int B, int C, int HW, int out_stride,
                              int out_offset, void* stream) {
  FPM_CHECK_ARG(fmap && out, "fpm_global_max: null tensor");
  if (B == 0) return FPM_OK;
  const long long rows = (long long)B * C;
  const int vec = (HW % 4 == 0 && ((uintptr_t)fmap & 15) == 0) ? 1 : 0;
  const int per_row = vec ? HW / 4 : HW;
  cudaStream_t st = (cudaStream_t)stream;
  if (per_row <= 24)
    fpm::global_max_kernel<8><<<fpm_cdiv(rows * 8, 256), 256, 0, st>>>(fmap, out, (int)rows, HW, out_stride, out_offset, C, vec);
  else if (per_row <= 48)
    fpm::global_max_kernel<16><<<fpm_cdiv(rows * 16, 256), 256, 0, st>>>(fmap, out, (int)rows, HW, out_stride, out_offset, C, vec);
  else
    fpm::global_max_kernel<32><<<fpm_cdiv(rows * 32, 256), 256, 0, st>>>(fmap, out, (int)rows, HW, out_stride, out_offset, C, vec);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_node_features(const float* nodes_nhwc, const float* edges_nhwc, const float* P,
                                 const long long* ns, const long long* ptr, float* X, int B, int nmax,
                                 int C1, int H1, int W1, int C2, int H2, int W2, float ori_w, float ori_h,
                                 void* stream) {
  FPM_CHECK_ARG(nodes_nhwc && edges_nhwc && P && ns && ptr && X, "fpm_node_features: null tensor");
  FPM_CHECK_ARG(C1 % 4 == 0 && C2 % 4 == 0, "fpm_node_features: channels must be multiples of 4");
  if (B == 0 || nmax == 0) return FPM_OK;
  const long long warps = (long long)B * nmax;
  fpm::node_features_kernel<<<fpm_cdiv(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      nodes_nhwc, edges_nhwc, P, (const int64_t*)ns, (const int64_t*)ptr, X, B, nmax, C1, H1, W1, C2, H2,
      W2, ori_w, ori_h);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_affinity_coeff(const float* gcat, const float* W, const float* bias, float* coeff,
                                  int B, int IN, int OUT, void* stream) {
  FPM_CHECK_ARG(gcat && W && bias && coeff, "fpm_affinity_coeff: null tensor");
  if (B == 0) return FPM_OK;
  fpm::affinity_coeff_kernel<<<dim3(B, fpm_cdiv(OUT, fpm::kCoeffRows)), 256, (size_t)IN * sizeof(float), (cudaStream_t)stream>>>(
      gcat, W, bias, coeff, IN, OUT);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_node_features_bwd(const float* dX, const float* P, const long long* ns, const long long* ptr,
                                     float* dnodes_nhwc, float* dedges_nhwc, int B, int nmax, int C1, int H1,
                                     int W1, int C2, int H2, int W2, float ori_w, float ori_h, void* stream) {
  FPM_CHECK_ARG(dX && P && ns && ptr && dnodes_nhwc && dedges_nhwc, "fpm_node_features_bwd: null tensor");
  FPM_CHECK_ARG(C1 % 64 == 0 && C2 % 64 == 0, "fpm_node_features_bwd: channels must be multiples of 64");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_node_features_bwd: batch too large");
  const size_t s1 = (size_t)H1 * W1 * 64 * sizeof(float), s2 = (size_t)H2 * W2 * 64 * sizeof(float);
  FPM_CHECK_ARG(s1 <= 200 * 1024 && s2 <= 200 * 1024, "fpm_node_features_bwd: feature map too large");
  FPM_CUDA(cudaFuncSetAttribute(fpm::node_features_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(s1 > s2 ? s1 : s2)));
  cudaStream_t st = (cudaStream_t)stream;
  fpm::node_features_bwd_kernel<<<dim3(C1 / 64, B), 256, s1, st>>>(dX, P, (const int64_t*)ns, (const int64_t*)ptr,
                                                                  dnodes_nhwc, nmax, C1, H1, W1, 0, C1 + C2, ori_w, ori_h);
  FPM_LAUNCH_CHECK();
  fpm::node_features_bwd_kernel<<<dim3(C2 / 64, B), 256, s2, st>>>(dX, P, (const int64_t*)ns, (const int64_t*)ptr,
                                                                  dedges_nhwc, nmax, C2, H2, W2, C1, C1 + C2, ori_w, ori_h);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_fmap_prep_bwd(const float* fmap_nchw, const float* dy_nhwc, float* dx_nchw, int B, int C,
                                 int Hf, int Wf, void* stream) {
  FPM_CHECK_ARG(fmap_nchw && dy_nhwc && dx_nchw, "fpm_fmap_prep_bwd: null tensor");
  FPM_CHECK_ARG(B >= 0 && C > 0 && Hf > 0 && Wf > 0, "fpm_fmap_prep_bwd: bad sizes");
  if (B == 0) return FPM_OK;
  const int HW = Hf * Wf;
  const size_t smem = (size_t)C * 33 * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024 && B <= 65535, "fpm_fmap_prep_bwd: channel count or batch too large");
  FPM_CUDA(cudaFuncSetAttribute(fpm::fmap_prep_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fpm::fmap_prep_bwd_kernel<<<dim3(fpm_cdiv(HW, 32), B), 256, smem, (cudaStream_t)stream>>>(fmap_nchw, dy_nhwc,
                                                                                           dx_nchw, C, HW);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
