// NGM association-graph message passing with the Kronecker structure kept factorised.
//
// Replaces, per pair and per layer, the chain of /root/reference/src/model/ngm.py:326-348:
//   construct_sparse_aff_mat (utils/factorize_graph_matching.py:57-95: index lists of e1*e2 + n1*n2
//   entries) -> torch_sparse.SparseTensor -> PYGNNLayer.forward (src/model/gnn.py:207-226:
//   SAGEConv mean aggregation + lin_l/lin_r + n_self_func + classifier).
//
// The association graph's edge (i2,i1) -> (j2,j1) exists iff i1->j1 is an edge of graph 1 and i2->j2 an
// edge of graph 2 (every column of G2 (x) G1 / H2 (x) H1 holds a single one), and SAGEConv drops the edge
// values, so the mean aggregation factorises:
//     agg[(j2,j1)] = ( sum_{i2 in In2(j2)} sum_{i1 in In1(j1)} x[(i2,i1)]  +  [p < n1_b*n2_b] x[p] )
//                    / ( |In2(j2)| * |In1(j1)| + [p < n1_b*n2_b] ),          p = j2*n1max + j1
// The (n1 n2)^2 affinity matrix and its index lists are never built.  One CTA per (pair, j2): it sums the
// |In2(j2)| source rows of x into shared memory once, then every thread finishes one node j1.
// HBM/L2-bound: ~7 row reads of [n1max, 17] per CTA, one [n1max, 16] row written.
#include "common.cuh"

namespace fpm {

// In-neighbour lists from the per-pair edge tables [B, 2, emax] (int32, -1 padded; row 0 = G-node
// (source), row 1 = H-node (target) of every G/H column).  Thread per destination node, edges scanned in
// column order -> deterministic.  in_ptr: [B, nmax + 1] (offsets local to the pair), in_src: [B, emax].
__global__ void assoc_in_csr_kernel(const int* __restrict__ edges, int* __restrict__ in_ptr,
                                    int* __restrict__ in_src, int nmax, int emax) {
  extern __shared__ int sh[];            // [2 * emax] edge table, then [nmax + 1] counts
  int* ssrc = sh; int* sdst = sh + emax; int* cnt = sh + 2 * emax;
  const int b = blockIdx.x;
  const int* eb = edges + (size_t)b * 2 * emax;
  for (int k = threadIdx.x; k < emax; k += blockDim.x) { ssrc[k] = eb[k]; sdst[k] = eb[emax + k]; }
  __syncthreads();
  for (int j = threadIdx.x; j < nmax; j += blockDim.x) {
    int c = 0;
    for (int k = 0; k < emax; ++k) c += (sdst[k] == j);
    cnt[j] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int j = 0; j < nmax; ++j) { const int c = cnt[j]; cnt[j] = run; run += c; }
    cnt[nmax] = run;
  }
  __syncthreads();
  for (int j = threadIdx.x; j <= nmax; j += blockDim.x) in_ptr[(size_t)b * (nmax + 1) + j] = cnt[j];
  for (int j = threadIdx.x; j < nmax; j += blockDim.x) {
    int w = cnt[j];
    for (int k = 0; k < emax; ++k)
      if (sdst[k] == j) in_src[(size_t)b * emax + (w++)] = ssrc[k];
  }
}

struct GnnWeights {
  const float* lin_l_w; const float* lin_l_b;   // [16, CIN], [16]
  const float* lin_r_w;                         // [16, CIN]
  const float* self0_w; const float* self0_b;   // [16, CIN], [16]
  const float* self2_w; const float* self2_b;   // [16, 16], [16]
  const float* cls_w; const float* cls_b;       // [16], [1]
};

constexpr int kF = 16;   // GNN_FEAT, ngm.py:47

// CIN = 1 (layer 0: emb = vec(Kp)) or 17 (x1 of the previous layer + its Sinkhorn channel).
// xprev:   [B, N, 16]   (CIN == 17 only), N = n1max*n2max, p = i2*n1max + i1
// mprev_t: [B, n2max, n1max]  the matrix channel in p order (Kp^T or Sinkhorn^T)
// xout:    [B, N, 16];  score: [B, n1max, n2max] (classifier output, Sinkhorn-ready layout)
template <int CIN>
__global__ void __launch_bounds__(128, 3)
gnn_layer_kernel(const float* __restrict__ xprev, const float* __restrict__ mprev_t,
                 const int* __restrict__ in_ptr1, const int* __restrict__ in_src1,
                 const int* __restrict__ in_ptr2, const int* __restrict__ in_src2,
                 const int64_t* __restrict__ n1, const int64_t* __restrict__ n2, GnnWeights w,
                 float* __restrict__ xout, float* __restrict__ score, int n1max, int n2max, int e1max,
                 int e2max) {
  // Shared layout: every row is padded to CP floats (a multiple of 4) so that the per-node loops below
  // read weights and partial sums as 128-bit broadcasts: one LDS.128 per 4 FMAs instead of one LDS per
  // FMA (the first version of this kernel was LSU-bound: 1256 LDS for 1183 FFMA per node).
  constexpr int CP = (CIN + 3) / 4 * 4;
  extern __shared__ __align__(16) float sm[];
  float* Rsum = sm;                               // [n1max][CP] sum over In2(j2) rows
  float* Wsh = sm + (size_t)n1max * CP;           // weights
  const int b = blockIdx.y, j2 = blockIdx.x;
  const int N = n1max * n2max;
  const int tid = threadIdx.x;

  float* wl = Wsh;                    // [16][CP]
  float* wr = wl + kF * CP;           // [16][CP]
  float* w0 = wr + kF * CP;           // [16][CP]
  float* w2 = w0 + kF * CP;           // [16][16]
  float* bl = w2 + kF * kF;           // [16]
  float* b0 = bl + kF;
  float* b2 = b0 + kF;
  float* wc = b2 + kF;                // [16] + bias
  for (int i = tid; i < kF * CP; i += blockDim.x) {
    const int o = i / CP, c = i - o * CP;
    const bool in = c < CIN;
    wl[i] = in ? w.lin_l_w[o * CIN + c] : 0.f;
    wr[i] = in ? w.lin_r_w[o * CIN + c] : 0.f;
    w0[i] = in ? w.self0_w[o * CIN + c] : 0.f;
  }
  for (int i = tid; i < kF * kF; i += blockDim.x) w2[i] = w.self2_w[i];
  if (tid < kF) {
    bl[tid] = w.lin_l_b[tid]; b0[tid] = w.self0_b[tid]; b2[tid] = w.self2_b[tid]; wc[tid] = w.cls_w[tid];
  }
  if (tid == 0) wc[kF] = w.cls_b[0];

  // ---- stage 1: Rsum[i1, c] = sum_{i2 in In2(j2)} feat[(i2, i1), c]
  const int* ip2 = in_ptr2 + (size_t)b * (n2max + 1);
  const int beg2 = ip2[j2], end2 = ip2[j2 + 1];
  const int* is2 = in_src2 + (size_t)b * e2max;
  const float* xb = (CIN > 1) ? xprev + (size_t)b * N * kF : nullptr;
  const float* mb = mprev_t + (size_t)b * N;
  // Each source row (i2, :) is contiguous ([n1max][16] floats + [n1max] for the matrix channel): stream it
  // with independent 128-bit loads, 4 per thread in flight, accumulating over In2(j2) in registers.  (The
  // first version walked In2 per element with dependent scalar loads and was latency-bound: 1.6 ms/layer.)
  if (CIN > 1) {
    const int nvec = n1max * (kF / 4);
    for (int f0 = 0; f0 < nvec; f0 += 4 * blockDim.x) {
      float4 acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int q = beg2; q < end2; ++q) {
        const float4* row = (const float4*)(xb + (size_t)is2[q] * n1max * kF);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int f = f0 + u * blockDim.x + tid;
          if (f < nvec) {
            const float4 v = row[f];
            acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = f0 + u * blockDim.x + tid;
        if (f < nvec) *(float4*)&Rsum[(size_t)(f >> 2) * CP + (f & 3) * 4] = acc[u];
      }
    }
  }
  for (int i1 = tid; i1 < n1max; i1 += blockDim.x) {
    float a = 0.f;
#pragma unroll 4
    for (int q = beg2; q < end2; ++q) a += mb[(size_t)is2[q] * n1max + i1];
    Rsum[(size_t)i1 * CP + (CIN - 1)] = a;
#pragma unroll
    for (int c = CIN; c < CP; ++c) Rsum[(size_t)i1 * CP + c] = 0.f;
  }
  __syncthreads();

  // ---- stage 2: one thread per node (j2, j1)
  const int n1b = (int)n1[b], n2b = (int)n2[b];
  const long long ndiag = (long long)n1b * (long long)n2b;
  const int* ip1 = in_ptr1 + (size_t)b * (n1max + 1);
  const int* is1 = in_src1 + (size_t)b * e1max;
  const int d2 = end2 - beg2;
  for (int j1 = tid; j1 < n1max; j1 += blockDim.x) {
    const size_t p = (size_t)j2 * n1max + j1;
    float own[CP], agg[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) { own[c] = 0.f; agg[c] = 0.f; }
    if (CIN > 1) {
      const float4* xp = (const float4*)(xb + p * kF);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 v = xp[t];
        own[t * 4] = v.x; own[t * 4 + 1] = v.y; own[t * 4 + 2] = v.z; own[t * 4 + 3] = v.w;
      }
    }
    own[CIN - 1] = mb[p];
    const int beg1 = ip1[j1], end1 = ip1[j1 + 1];
    for (int q = beg1; q < end1; ++q) {
      const float4* r = (const float4*)(Rsum + (size_t)is1[q] * CP);
#pragma unroll
      for (int t = 0; t < CP / 4; ++t) {
        const float4 v = r[t];
        agg[t * 4] += v.x; agg[t * 4 + 1] += v.y; agg[t * 4 + 2] += v.z; agg[t * 4 + 3] += v.w;
      }
    }
    long long cnt = (long long)d2 * (long long)(end1 - beg1);
    if ((long long)p < ndiag) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) agg[c] += own[c];
      cnt += 1;
    }
    const float inv = cnt > 0 ? (float)cnt : 1.f;
#pragma unroll
    for (int c = 0; c < CIN; ++c) agg[c] = agg[c] / inv;

    // padded lanes (c >= CIN) multiply zeros: fmaf(0, 0, a) == a, so the sums keep their c-order value
    float h[kF];
#pragma unroll
    for (int o = 0; o < kF; ++o) {
      float a = b0[o];
#pragma unroll
      for (int t = 0; t < CP / 4; ++t) {
        const float4 wv = *(const float4*)&w0[o * CP + t * 4];
        a = fmaf(wv.x, own[t * 4], a); a = fmaf(wv.y, own[t * 4 + 1], a);
        a = fmaf(wv.z, own[t * 4 + 2], a); a = fmaf(wv.w, own[t * 4 + 3], a);
      }
      h[o] = fmaxf(a, 0.f);
    }
    float x1[kF];
    float sc = wc[kF];
#pragma unroll
    for (int o = 0; o < kF; ++o) {
      float a = bl[o];
      float r = 0.f;
#pragma unroll
      for (int t = 0; t < CP / 4; ++t) {
        const float4 lv = *(const float4*)&wl[o * CP + t * 4];
        const float4 rv = *(const float4*)&wr[o * CP + t * 4];
        a = fmaf(lv.x, agg[t * 4], a); a = fmaf(lv.y, agg[t * 4 + 1], a);
        a = fmaf(lv.z, agg[t * 4 + 2], a); a = fmaf(lv.w, agg[t * 4 + 3], a);
        r = fmaf(rv.x, own[t * 4], r); r = fmaf(rv.y, own[t * 4 + 1], r);
        r = fmaf(rv.z, own[t * 4 + 2], r); r = fmaf(rv.w, own[t * 4 + 3], r);
      }
      float s2 = b2[o];
#pragma unroll
      for (int t = 0; t < kF / 4; ++t) {
        const float4 wv = *(const float4*)&w2[o * kF + t * 4];
        s2 = fmaf(wv.x, h[t * 4], s2); s2 = fmaf(wv.y, h[t * 4 + 1], s2);
        s2 = fmaf(wv.z, h[t * 4 + 2], s2); s2 = fmaf(wv.w, h[t * 4 + 3], s2);
      }
      const float v = (a + r) + fmaxf(s2, 0.f);
      x1[o] = v;
      sc = fmaf(wc[o], v, sc);
    }
    float4* dst = (float4*)(xout + ((size_t)b * N + p) * kF);
    dst[0] = make_float4(x1[0], x1[1], x1[2], x1[3]);
    dst[1] = make_float4(x1[4], x1[5], x1[6], x1[7]);
    dst[2] = make_float4(x1[8], x1[9], x1[10], x1[11]);
    dst[3] = make_float4(x1[12], x1[13], x1[14], x1[15]);
    score[((size_t)b * n1max + j1) * n2max + j2] = sc;
  }
}

// s[b, i1, i2] = classifier([x1[b, p, :], sk[b, i1, i2]]),  p = i2*n1max + i1   (ngm.py:368-369)
__global__ void final_classifier_kernel(const float* __restrict__ x1, const float* __restrict__ sk_t,
                                        const float* __restrict__ cw, const float* __restrict__ cb,
                                        float* __restrict__ s, int n1max, int n2max) {
  const int b = blockIdx.y;
  const int N = n1max * n2max;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const float4* xp = (const float4*)(x1 + ((size_t)b * N + p) * kF);
  float acc = cb[0];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = xp[q];
    acc = fmaf(cw[q * 4 + 0], v.x, acc); acc = fmaf(cw[q * 4 + 1], v.y, acc);
    acc = fmaf(cw[q * 4 + 2], v.z, acc); acc = fmaf(cw[q * 4 + 3], v.w, acc);
  }
  acc = fmaf(cw[kF], sk_t[(size_t)b * N + p], acc);
  const int i2 = p / n1max, i1 = p - i2 * n1max;
  s[((size_t)b * n1max + i1) * n2max + i2] = acc;
}

}  // namespace fpm

extern "C" int fpm_assoc_in_csr(const int* edges, int* in_ptr, int* in_src, int B, int nmax, int emax,
                                void* stream) {
  FPM_CHECK_ARG(edges && in_ptr && in_src, "fpm_assoc_in_csr: null tensor");
  FPM_CHECK_ARG(B >= 0 && nmax > 0 && emax >= 0, "fpm_assoc_in_csr: bad sizes");
  if (B == 0) return FPM_OK;
  const size_t smem = (size_t)(2 * emax + nmax + 1) * sizeof(int);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_assoc_in_csr: graph too large for one CTA");
  FPM_CUDA(cudaFuncSetAttribute(fpm::assoc_in_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  fpm::assoc_in_csr_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(edges, in_ptr, in_src, nmax, emax);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// weights: 9 device pointers in the order of GnnWeights.
extern "C" int fpm_gnn_layer(const float* xprev, const float* mprev_t, const int* in_ptr1,
                             const int* in_src1, const int* in_ptr2, const int* in_src2,
                             const long long* n1, const long long* n2, const float* const* weights,
                             float* xout, float* score, int B, int n1max, int n2max, int e1max, int e2max,
                             int cin, void* stream) {
  FPM_CHECK_ARG(mprev_t && in_ptr1 && in_src1 && in_ptr2 && in_src2 && n1 && n2 && weights && xout && score,
                "fpm_gnn_layer: null tensor");
  FPM_CHECK_ARG(cin == 1 || (cin == 17 && xprev), "fpm_gnn_layer: cin must be 1 or 17 (with xprev)");
  FPM_CHECK_ARG(B >= 0 && n1max > 0 && n2max > 0, "fpm_gnn_layer: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_gnn_layer: batch too large");
  fpm::GnnWeights w{weights[0], weights[1], weights[2], weights[3], weights[4],
                    weights[5], weights[6], weights[7], weights[8]};
  for (int i = 0; i < 9; ++i) FPM_CHECK_ARG(weights[i], "fpm_gnn_layer: null weight");
  const int cp = (cin + 3) / 4 * 4;
  const size_t smem = ((size_t)n1max * cp + 3 * 16 * cp + 16 * 16 + 4 * 16 + 4) * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_gnn_layer: n1max too large");
  dim3 grid(n2max, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (cin == 1) {
    FPM_CUDA(cudaFuncSetAttribute(fpm::gnn_layer_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fpm::gnn_layer_kernel<1><<<grid, 128, smem, st>>>(xprev, mprev_t, in_ptr1, in_src1, in_ptr2, in_src2,
                                                      (const int64_t*)n1, (const int64_t*)n2, w, xout, score,
                                                      n1max, n2max, e1max, e2max);
  } else {
    FPM_CUDA(cudaFuncSetAttribute(fpm::gnn_layer_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fpm::gnn_layer_kernel<17><<<grid, 128, smem, st>>>(xprev, mprev_t, in_ptr1, in_src1, in_ptr2, in_src2,
                                                       (const int64_t*)n1, (const int64_t*)n2, w, xout, score,
                                                       n1max, n2max, e1max, e2max);
  }
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_final_classifier(const float* x1, const float* sk_t, const float* cw, const float* cb,
                                    float* s, int B, int n1max, int n2max, void* stream) {
  FPM_CHECK_ARG(x1 && sk_t && cw && cb && s, "fpm_final_classifier: null tensor");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_final_classifier: batch too large");
  dim3 grid(fpm_cdiv((long long)n1max * n2max, 256), B);
  fpm::final_classifier_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x1, sk_t, cw, cb, s, n1max, n2max);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
