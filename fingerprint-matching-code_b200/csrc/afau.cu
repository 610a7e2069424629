// AFA-U k-prediction module: fused mixed-score cross attention, add + InstanceNorm, k head.
//
// Replaces /root/reference/src/model/afau.py:231-300 (CrossSet_MultiHeadAttention, which materialises
// [B, n, 16, n, 16] fp32 intermediates = 2.6 GB per block at B=256, n=100), afau.py:145-176
// (AddAndInstanceNormalization) and the max-pool + MLP + sigmoid head of
// /root/reference/src/model/ngm.py:402-412.  The q/k/v, combine and feed-forward projections are plain
// dense GEMMs and go through gemm_*.cu.
#include "common.cuh"
#include <stdlib.h>

namespace fpm {

constexpr int kHeads = 16, kQkv = 16, kMs = 16;

// One CTA per (row tile of 128, head, pair); thread = one query row.  k_h / v_h of the pair are staged in
// shared memory ([nc][16] each); the softmax runs online (running maximum), so every score is evaluated once and
// the [n, n] score tile never leaves registers.  No masking: softmax runs over all nc columns (afau.py:288).
// cost is addressed as cost[b*cs_b + i*cs_r + j*cs_c] so the column block can pass cost^T without a copy.
__global__ void __launch_bounds__(128)
afau_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ cost, long long cs_b, long long cs_r, long long cs_c,
                      const float* __restrict__ mix1_w, const float* __restrict__ mix1_b,
                      const float* __restrict__ mix2_w, const float* __restrict__ mix2_b,
                      float* __restrict__ out, int nr, int nc, int q_zero) {
  extern __shared__ float sm[];
  float* ks = sm;                    // [nc][16]
  float* vs = sm + (size_t)nc * kQkv;
  __shared__ float w1a[kMs], w1b[kMs], b1[kMs], w2[kMs];
  __shared__ float b2s;
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int E = kHeads * kQkv;
  for (int idx = threadIdx.x; idx < nc * kQkv; idx += blockDim.x) {
    const int j = idx / kQkv, d = idx - j * kQkv;
    const size_t g = ((size_t)b * nc + j) * E + h * kQkv + d;
    ks[idx] = q_zero ? 0.f : k[g];
    vs[idx] = v[g];
  }
  if (threadIdx.x < kMs) {
    w1a[threadIdx.x] = mix1_w[(h * 2 + 0) * kMs + threadIdx.x];
    w1b[threadIdx.x] = mix1_w[(h * 2 + 1) * kMs + threadIdx.x];
    b1[threadIdx.x] = mix1_b[h * kMs + threadIdx.x];
    w2[threadIdx.x] = mix2_w[h * kMs + threadIdx.x];
  }
  if (threadIdx.x == 0) b2s = mix2_b[h];
  __syncthreads();
  if (i >= nr) return;

  float qv[kQkv];
  {
    const float4* qp = (const float4*)(q + ((size_t)b * nr + i) * E + h * kQkv);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float4 x = qp[t];
      qv[t * 4] = x.x; qv[t * 4 + 1] = x.y; qv[t * 4 + 2] = x.z; qv[t * 4 + 3] = x.w;
    }
  }
  const float* crow = cost + (size_t)b * cs_b + (size_t)i * cs_r;

  // q_zero: the caller guarantees q == 0 (the row block of Net.forward: the row embedding is all zeros, ngm.py:392),
  // so q . k = 0 exactly and the 16-term dot product is skipped - same bits, a third of the work.
  auto score = [&](int j) -> float {
    float dot = 0.f;
    if (!q_zero) {
#pragma unroll
      for (int d = 0; d < kQkv; ++d) dot = fmaf(qv[d], ks[j * kQkv + d], dot);
    }
    dot = dot / 4.0f;                                        // / sqrt(qkv_dim)
    const float c = crow[(size_t)j * cs_c];
    float s = 0.f;
#pragma unroll
    for (int m = 0; m < kMs; ++m) {
      const float h1 = fmaxf(fmaf(c, w1b[m], dot * w1a[m]) + b1[m], 0.f);
      s = fmaf(h1, w2[m], s);
    }
    return s + b2s;
  };

  // One pass with a running maximum (the score MLP is the expensive part: evaluating it once per column instead of
  // once for the maximum and once for the weights takes 40 % off the kernel).  When the maximum grows, the
  // denominator and the 16 accumulators are rescaled by exp(old - new); a row does that ~ln(nc) times.
  float mx = kNegInf;
  float den = 0.f;
  float acc[kQkv];
#pragma unroll
  for (int d = 0; d < kQkv; ++d) acc[d] = 0.f;
  for (int j = 0; j < nc; ++j) {
    const float sj = score(j);
    if (sj > mx) {
      const float r = expf(mx - sj);                          // exp(-inf) = 0 on the first column
      den *= r;
#pragma unroll
      for (int d = 0; d < kQkv; ++d) acc[d] *= r;
      mx = sj;
    }
    const float e = expf(sj - mx);
    den += e;
#pragma unroll
    for (int d = 0; d < kQkv; ++d) acc[d] = fmaf(e, vs[j * kQkv + d], acc[d]);
  }
  float4* op = (float4*)(out + ((size_t)b * nr + i) * E + h * kQkv);
#pragma unroll
  for (int t = 0; t < 4; ++t)
    op[t] = make_float4(acc[t * 4] / den, acc[t * 4 + 1] / den, acc[t * 4 + 2] / den, acc[t * 4 + 3] / den);
}

// The row block of Net.forward (ngm.py:392: the row embedding is all zeros, so q = 0 and the score of (i, j) is the
// mixing MLP of cost[i, j] alone).  CTA per (pair, kZh heads), thread = one (head, query row); the heads' values are
// staged in shared memory.  The 16-unit MLP runs on packed fp32: 8 FFMA2 form the hidden pairs (c * w1b + b1, one
// rounding instead of the generic kernel's two), 16 FMNMX, 8 FFMA2 fold them with w2; the 16 value accumulators are
// 8 FFMA2.  Scores are kept in the log2 domain (s * log2 e) so a softmax weight is one FFMA + one EX2.
// kStage = false (cost addressed with unit ROW stride, i.e. the caller holds cost^T - Net.forward passes the
// transposed copy the Sinkhorn kernel writes anyway): the threads of a warp read one column of their rows with a
// single coalesced load, four columns prefetched ahead, and the CTA needs only 13 KB of shared memory - it fits
// beside a GEMM CTA of another stream.  kStage = true: the pair's cost tile goes through shared memory first
// (odd row pitch: conflict-free), by warp-per-row loops with four loads in flight.
constexpr int kZh = 2;

template <bool kStage>
__global__ void __maxnreg__(96)       // 3 CTAs of 7 warps per SM; two of its warps fit a sub-partition beside a GEMM CTA
afau_attention_qzero_kernel(const float* __restrict__ v, const float* __restrict__ cost, long long cs_b, long long cs_r,
                            long long cs_c, const float* __restrict__ mix1_w, const float* __restrict__ mix1_b,
                            const float* __restrict__ mix2_w, const float* __restrict__ mix2_b,
                            float* __restrict__ out, int nr, int nc) {
  extern __shared__ float sm[];
  const int ldc = nc | 1;
  float* cs = sm;                                                           // [nr][ldc] (kStage only)
  float* vs = kStage ? sm + (((size_t)nr * ldc + 3) & ~(size_t)3) : sm;     // [kZh][nc][16], 16-byte aligned
  const int b = blockIdx.y, h0 = blockIdx.x * kZh;
  const int E = kHeads * kQkv;
  const float* cb = cost + (size_t)b * cs_b;
  if (kStage) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (cs_c == 1) {
      for (int i = warp; i < nr; i += nwarps) {
        const float* src = cb + (size_t)i * cs_r;
#pragma unroll 4
        for (int j = lane; j < nc; j += 32) cs[i * ldc + j] = src[j];
      }
    } else {
      for (int j = warp; j < nc; j += nwarps) {
        const float* src = cb + (size_t)j * cs_c;
#pragma unroll 4
        for (int i = lane; i < nr; i += 32) cs[i * ldc + j] = src[(size_t)i * cs_r];
      }
    }
  }
#pragma unroll
  for (int hl = 0; hl < kZh; ++hl) {
    const float* src = v + (size_t)b * nc * E + (h0 + hl) * kQkv;
#pragma unroll 4
    for (int e = threadIdx.x; e < nc * kQkv; e += blockDim.x)
      vs[hl * nc * kQkv + e] = src[(size_t)(e >> 4) * E + (e & 15)];
  }
  __syncthreads();
  const int hl = threadIdx.x / nr, i = threadIdx.x - hl * nr;
  if (hl >= kZh) return;
  const int h = h0 + hl;

  constexpr float kL2e = 1.4426950408889634f;
  f32x2 w1[kMs / 2], bb[kMs / 2], w2[kMs / 2];
#pragma unroll
  for (int m = 0; m < kMs / 2; ++m) {
    w1[m] = pk2(mix1_w[(h * 2 + 1) * kMs + 2 * m], mix1_w[(h * 2 + 1) * kMs + 2 * m + 1]);
    bb[m] = pk2(mix1_b[h * kMs + 2 * m], mix1_b[h * kMs + 2 * m + 1]);
    w2[m] = pk2(mix2_w[h * kMs + 2 * m] * kL2e, mix2_w[h * kMs + 2 * m + 1] * kL2e);
  }
  const float b2 = mix2_b[h] * kL2e;
  const float4* vh = (const float4*)(vs + (size_t)hl * nc * kQkv);

  float mx = kNegInf, den = 0.f;
  f32x2 acc[kQkv / 2];
#pragma unroll
  for (int d = 0; d < kQkv / 2; ++d) acc[d] = pk2(0.f, 0.f);
  auto column = [&](int j, float c) {
    f32x2 s2 = pk2(b2, 0.f);
#pragma unroll
    for (int m = 0; m < kMs / 2; ++m) {
      float h0v, h1v;
      upk2(fma2(pk2(c, c), w1[m], bb[m]), h0v, h1v);
      s2 = fma2(pk2(fmaxf(h0v, 0.f), fmaxf(h1v, 0.f)), w2[m], s2);
    }
    float sa, sb;
    upk2(s2, sa, sb);
    const float sj = sa + sb;                         // score * log2 e
    if (sj > mx) {                                    // a row does this ~ln(nc) times
      float r;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(mx - sj));     // 2^-inf = 0 on the first column
      den *= r;
      const f32x2 rr = pk2(r, r);
#pragma unroll
      for (int d = 0; d < kQkv / 2; ++d) acc[d] = fma2(acc[d], rr, pk2(0.f, 0.f));
      mx = sj;
    }
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(sj - mx));
    den += e;
    const f32x2 ee = pk2(e, e);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float4 x = vh[j * 4 + t];
      acc[2 * t] = fma2(ee, pk2(x.x, x.y), acc[2 * t]);
      acc[2 * t + 1] = fma2(ee, pk2(x.z, x.w), acc[2 * t + 1]);
    }
  };
  if (kStage) {
    const float* crow = cs + i * ldc;
    for (int j = 0; j < nc; ++j) column(j, crow[j]);
  } else {
    const float* crow = cb + (size_t)i * cs_r;        // cs_r == 1: lanes = consecutive rows, one line per column
    float nxt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) nxt[u] = u < nc ? __ldg(crow + (size_t)u * cs_c) : 0.f;
    for (int j0 = 0; j0 < nc; j0 += 4) {
      float cur[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
#pragma unroll
      for (int u = 0; u < 4; ++u) nxt[u] = j0 + 4 + u < nc ? __ldg(crow + (size_t)(j0 + 4 + u) * cs_c) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < nc) column(j0 + u, cur[u]);
    }
  }
  float4* op = (float4*)(out + ((size_t)b * nr + i) * E + h * kQkv);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float a0, a1, a2, a3;
    upk2(acc[2 * t], a0, a1);
    upk2(acc[2 * t + 1], a2, a3);
    op[t] = make_float4(a0 / den, a1 / den, a2 / den, a3 / den);
  }
}

// out[b, r, e] = InstanceNorm over r of (a + other)[b, :, e] * gamma[e] + beta[e]; other_mode: 0 none,
// 1 tensor [B, n, E], 2 row vector [E].  Optionally rowmax[b, e] = max_r out[b, r, e] (the
// "pad to 600 rows with -inf, MaxPool1d(600)" of ngm.py:402-405).  Thread per channel, coalesced along E.
__global__ void __launch_bounds__(128)
add_instnorm_kernel(const float* __restrict__ a, const float* __restrict__ other, int other_mode,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                    float* __restrict__ rowmax, int n, int E, float eps) {
  const int b = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float* ab = a + (size_t)b * n * E + e;
  const float* ob = other_mode == 1 ? other + (size_t)b * n * E + e : nullptr;
  const float ov = other_mode == 2 ? other[e] : 0.f;
  float sum = 0.f;
  for (int r = 0; r < n; ++r) {
    const float x = ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov);
    sum += x;
  }
  const float mean = sum / (float)n;
  float vs = 0.f;
  for (int r = 0; r < n; ++r) {
    const float x = ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov);
    const float d = x - mean;
    vs = fmaf(d, d, vs);
  }
  const float inv = 1.0f / sqrtf(vs / (float)n + eps);
  const float g = gamma[e], bt = beta[e];
  float mxv = kNegInf;
  float* outb = out + (size_t)b * n * E + e;
  for (int r = 0; r < n; ++r) {
    const float x = ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov);
    const float y = (x - mean) * inv * g + bt;
    outb[(size_t)r * E] = y;
    mxv = fmaxf(mxv, y);
  }
  if (rowmax) rowmax[(size_t)b * E + e] = mxv;
}

// Same computation for n <= NR rows with the column held in registers: one read of the inputs instead of three
// (all loads independent and in flight together), same operations in the same order -> identical bits.
template <int NR>
__global__ void __launch_bounds__(128)
add_instnorm_reg_kernel(const float* __restrict__ a, const float* __restrict__ other, int other_mode,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                        float* __restrict__ rowmax, int n, int E, float eps) {
  const int b = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float* ab = a + (size_t)b * n * E + e;
  const float* ob = other_mode == 1 ? other + (size_t)b * n * E + e : nullptr;
  const float ov = other_mode == 2 ? other[e] : 0.f;
  float x[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) x[r] = r < n ? ab[(size_t)r * E] : 0.f;
  if (ob) {
#pragma unroll
    for (int r = 0; r < NR; ++r) x[r] = r < n ? x[r] + ob[(size_t)r * E] : 0.f;
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) x[r] = r < n ? x[r] + ov : 0.f;
  }
  float sum = 0.f;
#pragma unroll
  for (int r = 0; r < NR; ++r) if (r < n) sum += x[r];
  const float mean = sum / (float)n;
  float vs = 0.f;
#pragma unroll
  for (int r = 0; r < NR; ++r) if (r < n) { const float d = x[r] - mean; vs = fmaf(d, d, vs); }
  const float inv = 1.0f / sqrtf(vs / (float)n + eps);
  const float g = gamma[e], bt = beta[e];
  float mxv = kNegInf;
  float* outb = out + (size_t)b * n * E + e;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (r < n) {
      const float y = (x[r] - mean) * inv * g + bt;
      outb[(size_t)r * E] = y;
      mxv = fmaxf(mxv, y);
    }
  }
  if (rowmax) rowmax[(size_t)b * E + e] = mxv;
}

// The kernel the matching head runs (n <= kWarps * kRows rows, E % 4 == 0): CTA per (pair, 128 channels).  A lane
// owns four adjacent channels (one 128-bit load per row, a warp reads 512 contiguous bytes of the row), a warp owns
// rows w, w+kWarps, ...; the per-channel sums over rows are folded across the warps through shared memory (mean, then
// the centred second moment - the two-pass form of torch's InstanceNorm, so near-constant channels stay accurate).
// The thread-per-channel kernels above kept a whole column (104 values, 168 registers) per thread: 17 % occupancy,
// 128-byte row segments, 16 % of the DRAM peak.  `out` may be null when only the row maximum is wanted (the second
// normalisation of an AFA-U block feeds nothing but the max-pool, ngm.py:402-405).  kMode: second operand none /
// tensor / row vector as above; 3 = row vector AND `a` is the one-hot column embedding of ngm.py:396-399 given by its
// row count hot[b] (never materialised).
template <int kWarps, int kRows, int kMode>
__global__ void __launch_bounds__(32 * kWarps, 1024 / (32 * kWarps))
add_instnorm_tile_kernel(const float* __restrict__ a, const float* __restrict__ other,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                         float* __restrict__ rowmax, int n, int E, float eps, const int64_t* __restrict__ hot) {
  __shared__ float4 red[kWarps][32];
  __shared__ float4 stat[32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * 128 + 4 * lane;
  const bool live = e < E;
  const size_t base = (size_t)b * n * E + e;
  float4 x[kRows];
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const int r = warp + kWarps * k;
    x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kMode == 3) {                                      // a = one-hot rows: a[b, r, c] = (c == r && r < hot[b])
      if (r < (int)hot[b] && (r >> 2) == (e >> 2)) {
        x[k].x = (r & 3) == 0 ? 1.f : 0.f; x[k].y = (r & 3) == 1 ? 1.f : 0.f;
        x[k].z = (r & 3) == 2 ? 1.f : 0.f; x[k].w = (r & 3) == 3 ? 1.f : 0.f;
      }
    } else if (live && r < n) {
      x[k] = *(const float4*)(a + base + (size_t)r * E);
    }
  }
  if (kMode == 1) {
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int r = warp + kWarps * k;
      if (live && r < n) {
        const float4 o = *(const float4*)(other + base + (size_t)r * E);
        x[k].x += o.x; x[k].y += o.y; x[k].z += o.z; x[k].w += o.w;
      }
    }
  } else if (kMode == 2 || kMode == 3) {
    float4 ov = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) ov = *(const float4*)(other + e);
#pragma unroll
    for (int k = 0; k < kRows; ++k) { x[k].x += ov.x; x[k].y += ov.y; x[k].z += ov.z; x[k].w += ov.w; }
  }
  // Combine the warps' partials: thread c < 128 owns channel c of the tile, adds the kWarps partials in warp order and
  // finishes the statistic (one division / square root per channel and CTA, not per thread), everybody reads it back.
  // kind 0: mean = sum / n;  1: 1 / sqrt(sum / n + eps);  2: max, written to rowmax by the owner (no read-back).
  const float fn = (float)n;
  auto fold = [&](float4 p, int kind) -> float4 {
    red[warp][lane] = p;
    __syncthreads();
    if (threadIdx.x < 128) {
      const float* rf = reinterpret_cast<const float*>(&red[0][0]) + threadIdx.x;
      float t = rf[0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) t = kind == 2 ? fmaxf(t, rf[w * 128]) : t + rf[w * 128];
      if (kind == 0) t = t / fn;
      if (kind == 1) t = 1.0f / sqrtf(t / fn + eps);
      if (kind == 2) {
        const int ec = blockIdx.x * 128 + threadIdx.x;
        if (ec < E) rowmax[(size_t)b * E + ec] = t;
      } else {
        reinterpret_cast<float*>(&stat[0])[threadIdx.x] = t;
      }
    }
    if (kind == 2) return p;
    __syncthreads();
    return stat[lane];
  };
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < kRows; ++k)
    if (warp + kWarps * k < n) { sum.x += x[k].x; sum.y += x[k].y; sum.z += x[k].z; sum.w += x[k].w; }
  const float4 mean = fold(sum, 0);
  float4 vs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    x[k].x -= mean.x; x[k].y -= mean.y; x[k].z -= mean.z; x[k].w -= mean.w;
    if (warp + kWarps * k < n) {
      vs.x = fmaf(x[k].x, x[k].x, vs.x); vs.y = fmaf(x[k].y, x[k].y, vs.y);
      vs.z = fmaf(x[k].z, x[k].z, vs.z); vs.w = fmaf(x[k].w, x[k].w, vs.w);
    }
  }
  const float4 inv = fold(vs, 1);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), bt = g;
  if (live) { g = *(const float4*)(gamma + e); bt = *(const float4*)(beta + e); }
  float4 mx = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const int r = warp + kWarps * k;
    if (live && r < n) {
      float4 y;
      y.x = x[k].x * inv.x * g.x + bt.x; y.y = x[k].y * inv.y * g.y + bt.y;
      y.z = x[k].z * inv.z * g.z + bt.z; y.w = x[k].w * inv.w * g.w + bt.w;
      if (out) *(float4*)(out + base + (size_t)r * E) = y;
      mx.x = fmaxf(mx.x, y.x); mx.y = fmaxf(mx.y, y.y); mx.z = fmaxf(mx.z, y.z); mx.w = fmaxf(mx.w, y.w);
    }
  }
  if (rowmax) fold(mx, 2);
}

// k[b, j, o] = j < n[b] ? W[o, j] : 0 : the projection of a one-hot embedding (ngm.py:396-399) is a
// column gather of the weight, exact because the remaining terms are products with 0.
__global__ void onehot_proj_kernel(const float* __restrict__ W, const int64_t* __restrict__ n,
                                   float* __restrict__ out, int nmax, int OUT, int IN) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nmax * OUT) return;
  const int j = idx / OUT, o = idx - j * OUT;
  out[(size_t)b * nmax * OUT + idx] = (j < (int)n[b] && j < IN) ? W[(size_t)o * IN + j] : 0.f;
}

// ks[b] = sigmoid((final_row(g_row[b]) + final_col(g_col[b])) / 2);  k_scaled[b] = ks[b] * min(n1_b, n2_b)
__global__ void __launch_bounds__(256)
k_head_kernel(const float* __restrict__ g_row, const float* __restrict__ g_col,
              const float* __restrict__ r0w, const float* __restrict__ r0b, const float* __restrict__ r2w,
              const float* __restrict__ r2b, const float* __restrict__ c0w, const float* __restrict__ c0b,
              const float* __restrict__ c2w, const float* __restrict__ c2b, const int64_t* __restrict__ n1,
              const int64_t* __restrict__ n2, float* __restrict__ ks, float* __restrict__ k_scaled, int E,
              int Hd, int mean_k) {
  __shared__ float hid[2][32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int u = warp; u < 2 * Hd; u += nw) {
    const int net = u / Hd, o = u - net * Hd;
    const float* g = (net ? g_col : g_row) + (size_t)b * E;
    const float* w = (net ? c0w : r0w) + (size_t)o * E;
    float acc = 0.f;
    for (int i = lane; i < E; i += 32) acc = fmaf(w[i], g[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) hid[net][o] = fmaxf(acc + (net ? c0b : r0b)[o], 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float kr = r2b[0], kc = c2b[0];
    for (int o = 0; o < Hd; ++o) { kr = fmaf(r2w[o], hid[0][o], kr); kc = fmaf(c2w[o], hid[1][o], kc); }
    const float logit = mean_k ? (kr + kc) / 2.f : kr;
    const float kv = 1.f / (1.f + expf(-logit));
    ks[b] = kv;
    if (k_scaled) {
      const long long m = n1[b] < n2[b] ? n1[b] : n2[b];
      k_scaled[b] = kv * (float)m;
    }
  }
}


// ------------------------------------------------------------------------------------------
// Backward of afau_attention_kernel (training of the AFA-U k-branch, stages 2-5 of train.py; the reference
// differentiates afau.py:253-297 with autograd through its [B,n,16,n,16] intermediates).
// Same decomposition as the forward: CTA per (row tile, head, pair), thread per query row, scores recomputed.
//   w_ij = softmax_j(s_ij),  out_i = sum_j w_ij v_j
//   ds_ij = w_ij (dout_i . v_j - dout_i . out_i);  dv_j += w_ij dout_i
//   s = sum_m relu(dot w1a_m + c w1b_m + b1_m) w2_m + b2  ->  gradients of the 65 per-head mixing parameters
//   (thread-local sums, block reduction, one atomic per parameter per CTA) and d dot -> dq_i, dk_j.
// dk / dv are accumulated with shared-memory atomics per CTA and added to global memory (zero-filled by the caller).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
afau_attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                          const float* __restrict__ cost, long long cs_b, long long cs_r, long long cs_c,
                          const float* __restrict__ mix1_w, const float* __restrict__ mix1_b,
                          const float* __restrict__ mix2_w, const float* __restrict__ mix2_b,
                          const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ dq,
                          float* __restrict__ dk, float* __restrict__ dv, float* __restrict__ dmix, int nr, int nc) {
  extern __shared__ float sm[];
  float* ks = sm;                                  // [nc][16]
  float* vs = ks + (size_t)nc * kQkv;
  float* dks = vs + (size_t)nc * kQkv;             // [nc][16] accumulators
  float* dvs = dks + (size_t)nc * kQkv;
  __shared__ float w1a[kMs], w1b[kMs], b1[kMs], w2[kMs];
  __shared__ float b2s;
  __shared__ float pred[4][4 * kMs + 1];
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int E = kHeads * kQkv;
  for (int idx = threadIdx.x; idx < nc * kQkv; idx += blockDim.x) {
    const int j = idx / kQkv, d = idx - j * kQkv;
    const size_t g = ((size_t)b * nc + j) * E + h * kQkv + d;
    ks[idx] = k[g];
    vs[idx] = v[g];
    dks[idx] = 0.f;
    dvs[idx] = 0.f;
  }
  if (threadIdx.x < kMs) {
    w1a[threadIdx.x] = mix1_w[(h * 2 + 0) * kMs + threadIdx.x];
    w1b[threadIdx.x] = mix1_w[(h * 2 + 1) * kMs + threadIdx.x];
    b1[threadIdx.x] = mix1_b[h * kMs + threadIdx.x];
    w2[threadIdx.x] = mix2_w[h * kMs + threadIdx.x];
  }
  if (threadIdx.x == 0) b2s = mix2_b[h];
  __syncthreads();

  // per-thread sums of the mixing-parameter gradients: w1a[16] w1b[16] b1[16] w2[16] b2
  float gw1a[kMs], gw1b[kMs], gb1[kMs], gw2[kMs], gb2 = 0.f;
#pragma unroll
  for (int m = 0; m < kMs; ++m) { gw1a[m] = 0.f; gw1b[m] = 0.f; gb1[m] = 0.f; gw2[m] = 0.f; }

  if (i < nr) {
    float qv[kQkv], go[kQkv], dqv[kQkv];
    float go_dot_out = 0.f;
    {
      const size_t base = ((size_t)b * nr + i) * E + h * kQkv;
#pragma unroll
      for (int d = 0; d < kQkv; ++d) {
        qv[d] = q[base + d]; go[d] = dout[base + d]; dqv[d] = 0.f;
        go_dot_out = fmaf(go[d], out[base + d], go_dot_out);
      }
    }
    const float* crow = cost + (size_t)b * cs_b + (size_t)i * cs_r;
    auto score = [&](int j, float& dot, float& c) -> float {
      dot = 0.f;
#pragma unroll
      for (int d = 0; d < kQkv; ++d) dot = fmaf(qv[d], ks[j * kQkv + d], dot);
      dot = dot / 4.0f;
      c = crow[(size_t)j * cs_c];
      float s = 0.f;
#pragma unroll
      for (int m = 0; m < kMs; ++m) {
        const float h1 = fmaxf(fmaf(c, w1b[m], dot * w1a[m]) + b1[m], 0.f);
        s = fmaf(h1, w2[m], s);
      }
      return s + b2s;
    };
    float mx = kNegInf, dot, c;
    for (int j = 0; j < nc; ++j) mx = fmaxf(mx, score(j, dot, c));
    float den = 0.f;
    for (int j = 0; j < nc; ++j) den += expf(score(j, dot, c) - mx);
    for (int j = 0; j < nc; ++j) {
      const float w = expf(score(j, dot, c) - mx) / den;
      float gv = 0.f;
#pragma unroll
      for (int d = 0; d < kQkv; ++d) gv = fmaf(go[d], vs[j * kQkv + d], gv);
      const float ds = w * (gv - go_dot_out);
#pragma unroll
      for (int d = 0; d < kQkv; ++d) atomicAdd(&dvs[j * kQkv + d], w * go[d]);
      float ddot = 0.f;
      gb2 += ds;
#pragma unroll
      for (int m = 0; m < kMs; ++m) {
        const float pre = fmaf(c, w1b[m], dot * w1a[m]) + b1[m];
        const float h1 = fmaxf(pre, 0.f);
        gw2[m] = fmaf(ds, h1, gw2[m]);
        const float dh = pre > 0.f ? ds * w2[m] : 0.f;
        gw1a[m] = fmaf(dh, dot, gw1a[m]);
        gw1b[m] = fmaf(dh, c, gw1b[m]);
        gb1[m] += dh;
        ddot = fmaf(dh, w1a[m], ddot);
      }
      ddot = ddot / 4.0f;
      if (ddot != 0.f) {
#pragma unroll
        for (int d = 0; d < kQkv; ++d) {
          dqv[d] = fmaf(ddot, ks[j * kQkv + d], dqv[d]);
          atomicAdd(&dks[j * kQkv + d], ddot * qv[d]);
        }
      }
    }
    const size_t base = ((size_t)b * nr + i) * E + h * kQkv;
#pragma unroll
    for (int d = 0; d < kQkv; ++d) dq[base + d] = dqv[d];
  }
  // block reduction of the 65 mixing-parameter sums, then one atomic each
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < kMs; ++m) {
    const float a0 = warp_sum(gw1a[m]), a1 = warp_sum(gw1b[m]), a2 = warp_sum(gb1[m]), a3 = warp_sum(gw2[m]);
    if (lane == 0) { pred[warp][m] = a0; pred[warp][kMs + m] = a1; pred[warp][2 * kMs + m] = a2; pred[warp][3 * kMs + m] = a3; }
  }
  {
    const float a4 = warp_sum(gb2);
    if (lane == 0) pred[warp][4 * kMs] = a4;
  }
  __syncthreads();
  if (threadIdx.x < 4 * kMs + 1) {
    const float t = pred[0][threadIdx.x] + pred[1][threadIdx.x] + pred[2][threadIdx.x] + pred[3][threadIdx.x];
    if (t != 0.f) atomicAdd(dmix + (size_t)h * (4 * kMs + 1) + threadIdx.x, t);
  }
  for (int idx = threadIdx.x; idx < nc * kQkv; idx += blockDim.x) {
    const int j = idx / kQkv, d = idx - j * kQkv;
    const size_t g = ((size_t)b * nc + j) * E + h * kQkv + d;
    if (dks[idx] != 0.f) atomicAdd(dk + g, dks[idx]);
    if (dvs[idx] != 0.f) atomicAdd(dv + g, dvs[idx]);
  }
}

// Backward of add_instnorm_kernel.  dy [B,n,E] and/or drowmax [B,E] (gradient of the row maximum, routed to the
// first arg-max row) -> dx [B,n,E] (= gradient of `a` and of a tensor `other`); dgamma/dbeta [E] and, for a
// row-vector `other` (mode 2), dvec [E] are accumulated with atomics (zero-filled by the caller).
__global__ void __launch_bounds__(128)
add_instnorm_bwd_kernel(const float* __restrict__ a, const float* __restrict__ other, int other_mode,
                        const float* __restrict__ gamma, const float* __restrict__ dy,
                        const float* __restrict__ drowmax, float* __restrict__ dx, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, float* __restrict__ dvec, int n, int E, float eps) {
  const int b = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float* ab = a + (size_t)b * n * E + e;
  const float* ob = other_mode == 1 ? other + (size_t)b * n * E + e : nullptr;
  const float ov = other_mode == 2 ? other[e] : 0.f;
  const float* dyb = dy ? dy + (size_t)b * n * E + e : nullptr;
  float sum = 0.f;
  for (int r = 0; r < n; ++r) sum += ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov);
  const float mean = sum / (float)n;
  float vs = 0.f;
  for (int r = 0; r < n; ++r) {
    const float d = ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov) - mean;
    vs = fmaf(d, d, vs);
  }
  const float inv = 1.0f / sqrtf(vs / (float)n + eps);
  const float g = gamma[e];
  int arg = -1;
  float gmax = 0.f;
  if (drowmax) {
    gmax = drowmax[(size_t)b * E + e];
    float mxv = kNegInf;
    for (int r = 0; r < n; ++r) {
      const float x = ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov);
      const float y = (x - mean) * inv * g;            // + beta: constant, does not move the arg-max
      if (y > mxv) { mxv = y; arg = r; }
    }
  }
  float s1 = 0.f, s2 = 0.f;                            // sum dy, sum dy * xhat
  for (int r = 0; r < n; ++r) {
    const float xh = (ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov) - mean) * inv;
    const float d = (dyb ? dyb[(size_t)r * E] : 0.f) + (r == arg ? gmax : 0.f);
    s1 += d;
    s2 = fmaf(d, xh, s2);
  }
  atomicAdd(dgamma + e, s2);
  atomicAdd(dbeta + e, s1);
  const float scale = g * inv / (float)n;
  float* dxb = dx + (size_t)b * n * E + e;
  float tot = 0.f;
  for (int r = 0; r < n; ++r) {
    const float xh = (ab[(size_t)r * E] + (ob ? ob[(size_t)r * E] : ov) - mean) * inv;
    const float d = (dyb ? dyb[(size_t)r * E] : 0.f) + (r == arg ? gmax : 0.f);
    const float v = scale * ((float)n * d - s1 - xh * s2);
    dxb[(size_t)r * E] = v;
    tot += v;
  }
  if (other_mode == 2 && dvec) atomicAdd(dvec + e, tot);
}

}  // namespace fpm

extern "C" int fpm_afau_attention(const float* q, const float* k, const float* v, const float* cost,
                                  long long cs_b, long long cs_r, long long cs_c, const float* mix1_w,
                                  const float* mix1_b, const float* mix2_w, const float* mix2_b, float* out,
                                  int B, int nr, int nc, int q_zero, void* stream) {
  FPM_CHECK_ARG(v && cost && mix1_w && mix1_b && mix2_w && mix2_b && out, "fpm_afau_attention: null tensor");
  FPM_CHECK_ARG(q_zero || (q && k), "fpm_afau_attention: q and k may only be omitted with q_zero");
  FPM_CHECK_ARG(B >= 0 && nr > 0 && nc > 0, "fpm_afau_attention: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_afau_attention: batch too large");
  static const bool qzero_path = [] {
    const char* e = getenv("FPMATCH_AFAU_QZERO");              // 0: the generic kernel also for q == 0 (A/B runs)
    return !(e && e[0] == '0');
  }();
  const bool direct = cs_r == 1;                             // cost^T: coalesced column reads, nothing staged
  const size_t vbytes = (size_t)fpm::kZh * nc * fpm::kQkv * sizeof(float);
  const size_t zsmem = direct ? vbytes : (((size_t)nr * (nc | 1) + 3) & ~(size_t)3) * sizeof(float) + vbytes;
  if (q_zero && qzero_path && fpm::kZh * nr <= 256 && zsmem <= 72 * 1024) {
    dim3 zgrid(fpm::kHeads / fpm::kZh, B);
    const int threads = (fpm::kZh * nr + 31) / 32 * 32;
    if (direct) {
      FPM_CUDA(cudaFuncSetAttribute(fpm::afau_attention_qzero_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
      fpm::afau_attention_qzero_kernel<false><<<zgrid, threads, zsmem, (cudaStream_t)stream>>>(
          v, cost, cs_b, cs_r, cs_c, mix1_w, mix1_b, mix2_w, mix2_b, out, nr, nc);
    } else {
      FPM_CUDA(cudaFuncSetAttribute(fpm::afau_attention_qzero_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
      fpm::afau_attention_qzero_kernel<true><<<zgrid, threads, zsmem, (cudaStream_t)stream>>>(
          v, cost, cs_b, cs_r, cs_c, mix1_w, mix1_b, mix2_w, mix2_b, out, nr, nc);
    }
    FPM_LAUNCH_CHECK();
    return FPM_OK;
  }
  FPM_CHECK_ARG(q && k, "fpm_afau_attention: this shape runs the generic kernel, which reads q and k");
  const size_t smem = (size_t)2 * nc * fpm::kQkv * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_afau_attention: too many columns");
  FPM_CUDA(cudaFuncSetAttribute(fpm::afau_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid(fpm_cdiv(nr, 128), fpm::kHeads, B);
  fpm::afau_attention_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(
      q, k, v, cost, cs_b, cs_r, cs_c, mix1_w, mix1_b, mix2_w, mix2_b, out, nr, nc, q_zero);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_add_instnorm(const float* a, const float* other, int other_mode, const float* gamma,
                                const float* beta, float* out, float* rowmax, int B, int n, int E, float eps,
                                void* stream) {
  FPM_CHECK_ARG(a && gamma && beta && (out || rowmax), "fpm_add_instnorm: null tensor");
  FPM_CHECK_ARG(other_mode == 0 || other, "fpm_add_instnorm: other tensor missing");
  FPM_CHECK_ARG(other_mode >= 0 && other_mode <= 2, "fpm_add_instnorm: bad mode");
  FPM_CHECK_ARG(B >= 0 && n > 0 && E > 0, "fpm_add_instnorm: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_add_instnorm: batch too large");
  dim3 grid(fpm_cdiv(E, 128), B);
  const bool aligned = E % 4 == 0 && ((uintptr_t)a | (uintptr_t)other | (uintptr_t)gamma | (uintptr_t)beta |
                                      (uintptr_t)out | (uintptr_t)rowmax) % 16 == 0;
  static const bool tile_path = [] {
    const char* e = getenv("FPMATCH_INSTNORM_TILE");           // 0: thread-per-channel kernels (A/B runs)
    return !(e && e[0] == '0');
  }();
  if (tile_path && aligned && n <= 112) {
#define FPM_INSTNORM(W, MODE)                                                                                   \
  fpm::add_instnorm_tile_kernel<W, 7, MODE><<<grid, 32 * W, 0, (cudaStream_t)stream>>>(a, other, gamma, beta, out, \
                                                                                        rowmax, n, E, eps, nullptr)
    if (n <= 56) {
      if (other_mode == 0) FPM_INSTNORM(8, 0); else if (other_mode == 1) FPM_INSTNORM(8, 1); else FPM_INSTNORM(8, 2);
    } else {
      if (other_mode == 0) FPM_INSTNORM(16, 0); else if (other_mode == 1) FPM_INSTNORM(16, 1); else FPM_INSTNORM(16, 2);
    }
#undef FPM_INSTNORM
    FPM_LAUNCH_CHECK();
    return FPM_OK;
  }
  FPM_CHECK_ARG(out, "fpm_add_instnorm: this shape needs the out tensor");
  if (n <= 104)
    fpm::add_instnorm_reg_kernel<104><<<grid, 128, 0, (cudaStream_t)stream>>>(a, other, other_mode, gamma, beta, out,
                                                                               rowmax, n, E, eps);
  else
    fpm::add_instnorm_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a, other, other_mode, gamma, beta, out,
                                                                      rowmax, n, E, eps);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// add_instnorm of (one-hot rows [B, n, E] with hot[b] ones on the diagonal) + vec: the first normalisation of the
// AFA-U column block (ngm.py:396-399 feeds afau.py:154-176) without materialising the embedding.
extern "C" int fpm_onehot_instnorm(const long long* hot, const float* vec, const float* gamma, const float* beta,
                                   float* out, float* rowmax, int B, int n, int E, float eps, void* stream) {
  FPM_CHECK_ARG(hot && vec && gamma && beta && (out || rowmax), "fpm_onehot_instnorm: null tensor");
  FPM_CHECK_ARG(B >= 0 && n > 0 && n <= 112 && E > 0 && E % 4 == 0, "fpm_onehot_instnorm: needs n <= 112, E % 4 == 0");
  FPM_CHECK_ARG((((uintptr_t)vec | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)out | (uintptr_t)rowmax) & 15) == 0,
                "fpm_onehot_instnorm: 16-byte alignment required");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_onehot_instnorm: batch too large");
  dim3 grid(fpm_cdiv(E, 128), B);
  if (n <= 56)
    fpm::add_instnorm_tile_kernel<8, 7, 3><<<grid, 256, 0, (cudaStream_t)stream>>>(
        nullptr, vec, gamma, beta, out, rowmax, n, E, eps, (const int64_t*)hot);
  else
    fpm::add_instnorm_tile_kernel<16, 7, 3><<<grid, 512, 0, (cudaStream_t)stream>>>(
        nullptr, vec, gamma, beta, out, rowmax, n, E, eps, (const int64_t*)hot);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_onehot_proj(const float* W, const long long* n, float* out, int B, int nmax, int OUT,
                               int IN, void* stream) {
  FPM_CHECK_ARG(W && n && out, "fpm_onehot_proj: null tensor");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_onehot_proj: batch too large");
  dim3 grid(fpm_cdiv((long long)nmax * OUT, 256), B);
  fpm::onehot_proj_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, (const int64_t*)n, out, nmax, OUT, IN);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_k_head(const float* g_row, const float* g_col, const float* const* weights,
                          const long long* n1, const long long* n2, float* ks, float* k_scaled, int B, int E,
                          int Hd, int mean_k, void* stream) {
  FPM_CHECK_ARG(g_row && g_col && weights && ks, "fpm_k_head: null tensor");
  FPM_CHECK_ARG(!k_scaled || (n1 && n2), "fpm_k_head: k_scaled needs n1, n2");
  FPM_CHECK_ARG(Hd > 0 && Hd <= 32, "fpm_k_head: hidden width must be <= 32");
  for (int i = 0; i < 8; ++i) FPM_CHECK_ARG(weights[i], "fpm_k_head: null weight");
  if (B == 0) return FPM_OK;
  fpm::k_head_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(
      g_row, g_col, weights[0], weights[1], weights[2], weights[3], weights[4], weights[5], weights[6],
      weights[7], (const int64_t*)n1, (const int64_t*)n2, ks, k_scaled, E, Hd, mean_k);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// dq [B,nr,256] is overwritten; dk, dv [B,nc,256] and dmix [16 heads x 65] (mix1_w[0][16] mix1_w[1][16] mix1_b[16]
// mix2_w[16] mix2_b per head) are ACCUMULATED: the caller zero-fills them.
extern "C" int fpm_afau_attention_bwd(const float* q, const float* k, const float* v, const float* cost, long long cs_b,
                                      long long cs_r, long long cs_c, const float* mix1_w, const float* mix1_b,
                                      const float* mix2_w, const float* mix2_b, const float* out, const float* dout,
                                      float* dq, float* dk, float* dv, float* dmix, int B, int nr, int nc,
                                      void* stream) {
  FPM_CHECK_ARG(q && k && v && cost && mix1_w && mix1_b && mix2_w && mix2_b && out && dout && dq && dk && dv && dmix,
                "fpm_afau_attention_bwd: null tensor");
  FPM_CHECK_ARG(B >= 0 && nr > 0 && nc > 0, "fpm_afau_attention_bwd: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_afau_attention_bwd: batch too large");
  const size_t smem = (size_t)4 * nc * fpm::kQkv * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_afau_attention_bwd: too many columns");
  FPM_CUDA(cudaFuncSetAttribute(fpm::afau_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(fpm_cdiv(nr, 128), fpm::kHeads, B);
  fpm::afau_attention_bwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(
      q, k, v, cost, cs_b, cs_r, cs_c, mix1_w, mix1_b, mix2_w, mix2_b, out, dout, dq, dk, dv, dmix, nr, nc);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_add_instnorm_bwd(const float* a, const float* other, int other_mode, const float* gamma,
                                    const float* dy, const float* drowmax, float* dx, float* dgamma, float* dbeta,
                                    float* dvec, int B, int n, int E, float eps, void* stream) {
  FPM_CHECK_ARG(a && gamma && dx && dgamma && dbeta && (dy || drowmax), "fpm_add_instnorm_bwd: null tensor");
  FPM_CHECK_ARG(other_mode == 0 || other, "fpm_add_instnorm_bwd: other tensor missing");
  FPM_CHECK_ARG(other_mode >= 0 && other_mode <= 2, "fpm_add_instnorm_bwd: bad mode");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_add_instnorm_bwd: batch too large");
  dim3 grid(fpm_cdiv(E, 128), B);
  fpm::add_instnorm_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a, other, other_mode, gamma, dy, drowmax, dx,
                                                                        dgamma, dbeta, dvec, n, E, eps);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
