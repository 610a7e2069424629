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
#include <stdlib.h>

namespace fpm {

// ------------------------------------------------------------------------------------------
// Effective association structure of a pair (what the reference's index lists actually describe).
//
// gmdataset.py:623-642 hands the model idxG = CSC indices of kron(G2, G1) and idxH = those of kron(H2, H1): one entry
// per NON-ZERO column, each matrix eliminating its own zero columns.  With a complete edge table the t-th entries of
// both lists belong to the same column t = k2*e1 + k1 and the association edge is (src2[k2], src1[k1]) ->
// (dst2[k2], dst1[k1]).  With a PARTIAL ground-truth permutation (gmdataset.py:345-352: G2 = perm^T G1, H2 = perm^T H1)
// G2 loses the columns whose source keypoint has no counterpart and H2 those whose target has none, so the t-th
// entries pair the a-th surviving G2 column with the a-th surviving H2 column: still a Kronecker structure, over the
// "effective" edge list (src2[gv(a)], dst2[hv(a)]).  ngm.py:333-342 then cuts [idx; 0..n1 n2-1] to
// common_len = min(len(row), len(col), e1_pyg*e2_pyg + n1*n2) (K_value is built from the PyG edge lists), which can
// drop the tail of the diagonal block and, if the lists are long enough, end inside a Kronecker column block.
// Per pair this kernel emits
//   eff1 / eff2 [2, emax]  effective edge tables (sources and targets compacted independently, -1 padded); eff2 holds
//                          only the column blocks that survive the cut completely,
//   part = (ps2, pd2, ccut, 0): the block the cut ends in - graph-2 edge ps2 -> pd2 combined with the graph-1 edges
//                          c < ccut only (ccut = 0: none),
//   ndiag                  number of diagonal (self-loop) entries that survive: association nodes p < ndiag,
//   status bit 0           the G / H lists of a graph have different lengths (asymmetric adjacency): the reference's
//                          pairing is then not a Kronecker structure; min(len) columns are used and the bit is set.
// For complete tables eff == edges, ndiag = n1*n2, ccut = 0.  One CTA per pair, one warp per list.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
assoc_effective_kernel(const int* __restrict__ edges1, const int* __restrict__ edges2,
                       const int64_t* __restrict__ eptr1, const int64_t* __restrict__ eptr2,
                       const int64_t* __restrict__ n1, const int64_t* __restrict__ n2, int* __restrict__ eff1,
                       int* __restrict__ eff2, int64_t* __restrict__ ndiag, int* __restrict__ part,
                       int* __restrict__ status, int e1max, int e2max) {
  extern __shared__ int sh[];            // gs1[e1max] hd1[e1max] gs2[e2max] hd2[e2max]
  __shared__ int cnt[4];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* lists[4] = {sh, sh + e1max, sh + 2 * e1max, sh + 2 * e1max + e2max};
  {
    const int emax = warp < 2 ? e1max : e2max;
    const int* src = (warp < 2 ? edges1 + (size_t)b * 2 * e1max : edges2 + (size_t)b * 2 * e2max) + (warp & 1) * emax;
    int* out = lists[warp];
    int w = 0;
    for (int k0 = 0; k0 < emax; k0 += 32) {
      const int k = k0 + lane;
      const int v = k < emax ? src[k] : -1;
      const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
      if (v >= 0) out[w + __popc(m & ((1u << lane) - 1u))] = v;
      w += __popc(m);
    }
    if (lane == 0) cnt[warp] = w;
  }
  __syncthreads();
  const int E1 = min(cnt[0], cnt[1]), E2 = min(cnt[2], cnt[3]);
  const long long L = (long long)E1 * E2;
  const long long nd = (long long)n1[b] * (long long)n2[b];
  const long long lval = (eptr1[b + 1] - eptr1[b]) * (eptr2[b + 1] - eptr2[b]) + nd;
  const long long common = min(L + nd, lval);
  const long long cut = min(L, common);
  const int afull = E1 > 0 ? (int)(cut / E1) : 0;
  const int ccut = E1 > 0 ? (int)(cut % E1) : 0;
  for (int k = threadIdx.x; k < e1max; k += blockDim.x) {
    const bool in = k < E1;
    eff1[(size_t)b * 2 * e1max + k] = in ? lists[0][k] : -1;
    eff1[(size_t)b * 2 * e1max + e1max + k] = in ? lists[1][k] : -1;
  }
  for (int k = threadIdx.x; k < e2max; k += blockDim.x) {
    const bool in = k < afull;
    eff2[(size_t)b * 2 * e2max + k] = in ? lists[2][k] : -1;
    eff2[(size_t)b * 2 * e2max + e2max + k] = in ? lists[3][k] : -1;
  }
  if (threadIdx.x == 0) {
    long long d = common - L;
    ndiag[b] = d < 0 ? 0 : (d > nd ? nd : d);
    const bool has = afull < E2 && ccut > 0;
    part[b * 4 + 0] = has ? lists[2][afull] : -1;
    part[b * 4 + 1] = has ? lists[3][afull] : -1;
    part[b * 4 + 2] = has ? ccut : 0;
    part[b * 4 + 3] = 0;
    if (cnt[0] != cnt[1] || cnt[2] != cnt[3]) atomicOr(status, 1);
  }
}

// In-neighbour lists from per-pair edge tables [B, 2, emax] (int32, -1 padded; row 0 = G-node (source), row 1 =
// H-node (target) of every column; a column counts only if BOTH ends are present).  One CTA per pair, a counting sort
// in shared memory: in-degree counts by atomics, scan, scatter of the column ids through atomic cursors, then one
// thread per node sorts its (short) list -> deterministic, ascending column ids inside every list whatever order the
// atomics produced.  (Before: one warp per destination node walked the whole edge table twice, O(n e): 23 us.)
// in_ptr: [B, nmax + 1] (offsets local to the pair), in_src: [B, emax], in_col (optional): [B, emax] column ids.
constexpr int kAssocThreads = 512;
__global__ void __launch_bounds__(kAssocThreads)
assoc_in_csr_kernel(const int* __restrict__ edges, int* __restrict__ in_ptr, int* __restrict__ in_src,
                    int* __restrict__ in_col, int nmax, int emax) {
  extern __shared__ int sh[];            // [3 * emax] sources, targets, sorted column ids; [2 * (nmax + 1)] cursors, offsets
  int* ssrc = sh; int* sdst = sh + emax; int* slist = sh + 2 * emax;
  int* cnt = sh + 3 * emax; int* off = cnt + nmax + 1;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int* eb = edges + (size_t)b * 2 * emax;
  for (int j = threadIdx.x; j <= nmax; j += blockDim.x) cnt[j] = 0;
  for (int k = threadIdx.x; k < emax; k += blockDim.x) {
    const int s_ = eb[k], d_ = eb[emax + k];
    ssrc[k] = s_;
    sdst[k] = (s_ >= 0 && s_ < nmax && d_ >= 0 && d_ < nmax) ? d_ : -1;   // a column without two valid ends is no edge
  }
  __syncthreads();
  for (int k = threadIdx.x; k < emax; k += blockDim.x)
    if (sdst[k] >= 0) atomicAdd(&cnt[sdst[k]], 1);
  __syncthreads();
  if (warp == 0) {                       // exclusive scan, 32 nodes at a time
    int base = 0;
    for (int j0 = 0; j0 < nmax; j0 += 32) {
      const int j = j0 + lane;
      const int c = j < nmax ? cnt[j] : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (j < nmax) { cnt[j] = base + inc - c; off[j] = base + inc - c; }
      base += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) { cnt[nmax] = base; off[nmax] = base; }
  }
  __syncthreads();
  for (int j = threadIdx.x; j <= nmax; j += blockDim.x) in_ptr[(size_t)b * (nmax + 1) + j] = off[j];
  for (int k = threadIdx.x; k < emax; k += blockDim.x)
    if (sdst[k] >= 0) slist[atomicAdd(&cnt[sdst[k]], 1)] = k;
  __syncthreads();
  for (int j = threadIdx.x; j < nmax; j += blockDim.x) {
    const int beg = off[j], end = off[j + 1];
    for (int a = beg + 1; a < end; ++a) {            // insertion sort: in-degrees are a handful
      const int key = slist[a];
      int q = a - 1;
      while (q >= beg && slist[q] > key) { slist[q + 1] = slist[q]; --q; }
      slist[q + 1] = key;
    }
  }
  __syncthreads();
  const int total = off[nmax];
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    const int k = slist[p];
    in_src[(size_t)b * emax + p] = ssrc[k];
    if (in_col) in_col[(size_t)b * emax + p] = k;
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

// The layer's weights live in CONSTANT memory: every thread of a warp multiplies by the same weight, so the FMAs take
// it from the constant bank through uniform registers and issue no per-thread load for it.  (With the weights in
// shared memory the kernel spent 27 % of its warp samples waiting on LDS - r1c ncu source view.)
// Layout, TRANSPOSED so that the weights of two adjacent OUTPUT channels for one input channel are an aligned pair
// (one operand of the packed FFMA2 of sm_100, see gnn_layer_kernel):
//   wlT[CP][16] wrT[CP][16] w0T[CP][16] w2T[16][16] bl[16] b0[16] b2[16] wc[16] cb     (xT[c][o] = x[o][c])
// gnn_pack_weights_kernel writes that layout into a device staging buffer and one cudaMemcpyToSymbolAsync
// (device to device, stream ordered) publishes it before the layer kernel.  The bank is one per device: launches
// from different streams are serialised against each other with an event (gnn_publish_weights).
constexpr int kGnnConstFloats = 3 * 16 * 20 + 16 * 16 + 4 * 16 + 4;
__constant__ __align__(16) float c_gnn[kGnnConstFloats];

template <int CP>
struct GnnOff {
  static constexpr int wl = 0, wr = 16 * CP, w0 = 32 * CP, w2 = 48 * CP, bl = w2 + 256, b0 = bl + 16, b2 = b0 + 16,
                       wc = b2 + 16, cb = wc + 16, total = cb + 1;
};
// element accessors (o = output channel, c = input channel)
#define GNN_WL(o, c) c_gnn[O::wl + (c) * 16 + (o)]
#define GNN_WR(o, c) c_gnn[O::wr + (c) * 16 + (o)]
#define GNN_W0(o, c) c_gnn[O::w0 + (c) * 16 + (o)]
#define GNN_W2(o, c) c_gnn[O::w2 + (c) * 16 + (o)]

template <int CIN>
__global__ void gnn_pack_weights_kernel(GnnWeights w, float* __restrict__ staging) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  using O = GnnOff<CP>;
  const int tid = threadIdx.x;
  for (int i = tid; i < 16 * CP; i += blockDim.x) {
    const int c = i / 16, o = i - c * 16;
    const bool in = c < CIN;
    staging[O::wl + i] = in ? w.lin_l_w[o * CIN + c] : 0.f;
    staging[O::wr + i] = in ? w.lin_r_w[o * CIN + c] : 0.f;
    staging[O::w0 + i] = in ? w.self0_w[o * CIN + c] : 0.f;
  }
  for (int i = tid; i < 256; i += blockDim.x) {
    const int c = i / 16, o = i - c * 16;
    staging[O::w2 + i] = w.self2_w[o * 16 + c];
  }
  if (tid < 16) {
    staging[O::bl + tid] = w.lin_l_b[tid]; staging[O::b0 + tid] = w.self0_b[tid];
    staging[O::b2 + tid] = w.self2_b[tid]; staging[O::wc + tid] = w.cls_w[tid];
  }
  if (tid == 0) staging[O::cb] = w.cls_b[0];
}

// acc[0..7] (output channel pairs (0,1) .. (14,15)) += WT[c][0..15] * x   - 8 FFMA2 with the weight pair as a
// uniform-register operand and x broadcast to both lanes
#define GNN_FMA_ROW(acc, base, c, x)                                                      \
  do {                                                                                    \
    const f32x2 xx__ = pk2((x), (x));                                                     \
    _Pragma("unroll") for (int p__ = 0; p__ < 8; ++p__)                                   \
      (acc)[p__] = fma2(pk2(c_gnn[(base) + (c) * 16 + 2 * p__], c_gnn[(base) + (c) * 16 + 2 * p__ + 1]), xx__, \
                        (acc)[p__]);                                                      \
  } while (0)

// Stage 1 of a (pair, j2) CTA, shared by the forward and backward kernels:
//   Rsum[i1, c] = sum_{i2 in In2(j2)} feat[(i2, i1), c]     (feat = 16 channels of xprev + the matrix channel)
//   Rp[i1, c]   = feat[(ps2, i1), c]  when the pair's cut-off block targets this j2 (see assoc_effective_kernel)
// Each source row (i2, :) is contiguous ([n1max][16] floats + [n1max] for the matrix channel): it is streamed with
// independent 128-bit loads, accumulated over In2(j2) in registers (packed adds).
template <int CIN>
__device__ __forceinline__ void gnn_stage1(const float* __restrict__ xb, const float* __restrict__ mb,
                                           const int* __restrict__ is2, int beg2, int end2, int ps2,
                                           float* __restrict__ Rsum, float* __restrict__ Rp, int n1max) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (CIN > 1) {
    const int nvec = n1max * (kF / 4);
    for (int f0 = 0; f0 < nvec; f0 += 2 * nt) {
      f32x2 acc[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) acc[u][0] = acc[u][1] = pk2(0.f, 0.f);
#pragma unroll 3
      for (int q = beg2; q < end2; ++q) {
        const ulonglong2* row = (const ulonglong2*)(xb + (size_t)is2[q] * n1max * kF);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int f = f0 + u * nt + tid;
          if (f < nvec) {
            const ulonglong2 v = row[f];
            acc[u][0] = add2(acc[u][0], v.x); acc[u][1] = add2(acc[u][1], v.y);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int f = f0 + u * nt + tid;
        if (f < nvec) *(ulonglong2*)&Rsum[(size_t)(f >> 2) * CP + (f & 3) * 4] = make_ulonglong2(acc[u][0], acc[u][1]);
      }
    }
    if (ps2 >= 0) {
      const float4* row = (const float4*)(xb + (size_t)ps2 * n1max * kF);
      for (int f = tid; f < nvec; f += nt) *(float4*)&Rp[(size_t)(f >> 2) * CP + (f & 3) * 4] = row[f];
    }
  }
  for (int i1 = tid; i1 < n1max; i1 += nt) {
    float a = 0.f;
#pragma unroll 4
    for (int q = beg2; q < end2; ++q) a += mb[(size_t)is2[q] * n1max + i1];
    Rsum[(size_t)i1 * CP + (CIN - 1)] = a;
#pragma unroll
    for (int c = CIN; c < CP; ++c) Rsum[(size_t)i1 * CP + c] = 0.f;
    if (ps2 >= 0) {
      Rp[(size_t)i1 * CP + (CIN - 1)] = mb[(size_t)ps2 * n1max + i1];
#pragma unroll
      for (int c = CIN; c < CP; ++c) Rp[(size_t)i1 * CP + c] = 0.f;
    }
  }
}

// own[] (the node's input features) and agg[] (mean over its association in-neighbours) of node (j2, j1); returns
// the divisor and whether the node carries a self loop.  Shared by the forward and backward kernels.
template <int CIN>
__device__ __forceinline__ float gnn_node_inputs(const float* __restrict__ xb, const float* __restrict__ mb,
                                                 const float* __restrict__ Rsum, const float* __restrict__ Rp,
                                                 const int* __restrict__ ip1, const int* __restrict__ is1,
                                                 const int* __restrict__ ic1, int j1, size_t p, int d2, int ccut,
                                                 long long ndiag, float* own, float* agg, bool& self) {
  constexpr int CP = (CIN + 3) / 4 * 4;
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
  if (ccut > 0) {                                   // CTA-uniform: the cut-off block of this pair targets j2
    for (int q = beg1; q < end1; ++q) {
      if (ic1[q] < ccut) {
        const float4* r = (const float4*)(Rp + (size_t)is1[q] * CP);
#pragma unroll
        for (int t = 0; t < CP / 4; ++t) {
          const float4 v = r[t];
          agg[t * 4] += v.x; agg[t * 4 + 1] += v.y; agg[t * 4 + 2] += v.z; agg[t * 4 + 3] += v.w;
        }
        cnt += 1;
      }
    }
  }
  self = (long long)p < ndiag;
  if (self) {
#pragma unroll
    for (int c = 0; c < CIN; ++c) agg[c] += own[c];
    cnt += 1;
  }
  const float inv = cnt > 0 ? (float)cnt : 1.f;
#pragma unroll
  for (int c = 0; c < CIN; ++c) agg[c] = agg[c] / inv;
  return inv;
}

// One NGM layer.  CIN = 1 (layer 0: emb = vec(Kp)) or 17 (x1 of the previous layer + its Sinkhorn channel).
// xprev:   [B, N, 16]   (CIN == 17 only), N = n1max*n2max, p = i2*n1max + i1
// mprev_t: [B, n2max, n1max]  the matrix channel in p order (Kp^T or Sinkhorn^T)
// xout:    [B, N, 16];  score: [B, n1max, n2max] (classifier output, Sinkhorn-ready layout)
//
// One CTA per (pair, group of `rows` consecutive j2), one thread per association node.  The kernel is bound by
// instruction issue (r2b ncu: 56 % issue slots busy at IPC 2.2 while the FMA pipe is 33 % busy), so the design
// minimises instructions per node:
//   * the 1 216 FMAs of a node (four 16 x {17,17,17,16} linears) are issued as 608 packed FFMA2 (sm_100: two fp32 FMAs
//     per instruction): accumulator pair = two adjacent output channels, multiplicand = the input value broadcast to
//     both lanes, multiplier = the transposed weight pair as a uniform-register operand (constant bank, no load),
//   * the In1 aggregation and stage 1 use packed adds, the mean is 3 instructions per channel (q = a r,
//     e = fma(-q, cnt, a), q += e r with r = RN(1 / cnt): the Markstein correction step of an IEEE division) instead of
//     17 full divisions,
//   * `rows` x n1max nodes per CTA fill whole warps (2 x 100 keypoints = 200 of 224 lanes instead of 100 of 128).
// History: one thread per node with scalar FMAs 0.49 ms per layer (r1); two half-threads per node 0.61 ms (r2b: the
// per-node bookkeeping - own load, aggregation, divisions - was duplicated and the instruction count doubled).
constexpr int kGnnMaxThreads = 256;

// MINB = CTAs per SM the register allocation is sized for: 3 -> 80 registers (12 bytes of spills), 4 -> 64 registers,
// no spills (ptxas finds the tighter schedule only when it is told to)
template <int CIN, int MINB>
__global__ void __launch_bounds__(kGnnMaxThreads, MINB)
gnn_layer_kernel(const float* __restrict__ xprev, const float* __restrict__ mprev_t,
                 const int* __restrict__ in_ptr1, const int* __restrict__ in_src1, const int* __restrict__ in_col1,
                 const int* __restrict__ in_ptr2, const int* __restrict__ in_src2,
                 const int64_t* __restrict__ ndiag_p, const int* __restrict__ part,
                 float* __restrict__ xout, float* __restrict__ score, int n1max, int n2max, int e1max,
                 int e2max, int rows) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  using O = GnnOff<CP>;
  extern __shared__ __align__(16) float sm[];
  float* Rsum = sm;                                      // [rows][n1max][CP] sums over In2(j2) rows
  float* Rp = Rsum + (size_t)rows * n1max * CP;          // [n1max][CP] row of the cut-off block (rarely used)
  int* sdeg = (int*)(Rp + (size_t)n1max * CP);           // [rows] |In2(j2)|
  const int b = blockIdx.y, j2base = blockIdx.x * rows;
  const int N = n1max * n2max;
  const int tid = threadIdx.x;

  const int* ip2 = in_ptr2 + (size_t)b * (n2max + 1);
  const int* is2 = in_src2 + (size_t)b * e2max;
  const float* xb = (CIN > 1) ? xprev + (size_t)b * N * kF : nullptr;
  const float* mb = mprev_t + (size_t)b * N;
  int pd2 = -1, ps2 = -1, ccut = 0;
  if (part != nullptr) { ps2 = part[b * 4]; pd2 = part[b * 4 + 1]; ccut = part[b * 4 + 2]; }
  if (ccut <= 0 || pd2 < j2base || pd2 >= j2base + rows) { pd2 = -1; ps2 = -1; ccut = 0; }
  const long long ndiag = ndiag_p[b];
  const int* ip1 = in_ptr1 + (size_t)b * (n1max + 1);
  const int* is1 = in_src1 + (size_t)b * e1max;
  const int* ic1 = in_col1 != nullptr ? in_col1 + (size_t)b * e1max : nullptr;

  // ---- stage 1: row sums over In2(j2) for every row of the group
  for (int r = 0; r < rows; ++r) {
    const int j2 = j2base + r;
    if (j2 >= n2max) break;
    const int beg2 = ip2[j2], end2 = ip2[j2 + 1];
    if (tid == 0) sdeg[r] = end2 - beg2;
    gnn_stage1<CIN>(xb, mb, is2, beg2, end2, j2 == pd2 ? ps2 : -1, Rsum + (size_t)r * n1max * CP, Rp, n1max);
  }
  __syncthreads();

  // ---- stage 2: one thread per node (j2, j1)
  const int nodes = rows * n1max;
  for (int t = tid; t < nodes; t += blockDim.x) {
    const int r = t / n1max, j1 = t - r * n1max;
    const int j2 = j2base + r;
    if (j2 >= n2max) break;
    const size_t p = (size_t)j2 * n1max + j1;
    const float* Rs = Rsum + (size_t)r * n1max * CP;
    // ---- (i) own features -> h0 = relu(W0 own + b0);  x1 accumulators = bl + Wr own.  The order of the four
    // linears is chosen for register pressure: own[] dies before the aggregation is formed, h0 before s2 is folded in.
    float own[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) own[c] = 0.f;
    if (CIN > 1) {
      const float4* xp = (const float4*)(xb + p * kF);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = xp[q];
        own[q * 4] = v.x; own[q * 4 + 1] = v.y; own[q * 4 + 2] = v.z; own[q * 4 + 3] = v.w;
      }
    }
    own[CIN - 1] = mb[p];
    f32x2 acc[8], xacc[8];
#pragma unroll
    for (int pp = 0; pp < 8; ++pp) {
      acc[pp] = pk2(c_gnn[O::b0 + 2 * pp], c_gnn[O::b0 + 2 * pp + 1]);
      xacc[pp] = pk2(c_gnn[O::bl + 2 * pp], c_gnn[O::bl + 2 * pp + 1]);
    }
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      GNN_FMA_ROW(acc, O::w0, c, own[c]);
      GNN_FMA_ROW(xacc, O::wr, c, own[c]);
    }
    const bool self = (long long)p < ndiag;
    // the self loop contributes own[] to the aggregation: keep it as packed pairs, own[] itself is dead from here on
    f32x2 ag[CP / 2];
#pragma unroll
    for (int c = 0; c < CP / 2; ++c) ag[c] = self ? pk2(own[2 * c], own[2 * c + 1]) : pk2(0.f, 0.f);
    // ---- (ii) s2 = relu(W2 h0 + b2), folded into the x1 accumulators
    {
      float h[kF];
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        upk2(acc[pp], h[2 * pp], h[2 * pp + 1]);
        h[2 * pp] = fmaxf(h[2 * pp], 0.f); h[2 * pp + 1] = fmaxf(h[2 * pp + 1], 0.f);
        acc[pp] = pk2(c_gnn[O::b2 + 2 * pp], c_gnn[O::b2 + 2 * pp + 1]);
      }
#pragma unroll
      for (int c = 0; c < kF; ++c) GNN_FMA_ROW(acc, O::w2, c, h[c]);
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        float s0, s1;
        upk2(acc[pp], s0, s1);
        xacc[pp] = add2(xacc[pp], pk2(fmaxf(s0, 0.f), fmaxf(s1, 0.f)));
      }
    }
    // ---- (iii) mean over the association in-neighbours: In1(j1) of the row sums (packed adds), the self loop
    const int beg1 = ip1[j1], end1 = ip1[j1 + 1];
    for (int q = beg1; q < end1; ++q) {
      const ulonglong2* rr = (const ulonglong2*)(Rs + (size_t)is1[q] * CP);
#pragma unroll
      for (int v = 0; v < CP / 4; ++v) {
        const ulonglong2 w = rr[v];
        ag[2 * v] = add2(ag[2 * v], w.x); ag[2 * v + 1] = add2(ag[2 * v + 1], w.y);
      }
    }
    long long cnt = (long long)sdeg[r] * (long long)(end1 - beg1) + (self ? 1 : 0);
    if (j2 == pd2) {                                    // rare: the cut-off block of this pair targets this row
      for (int q = beg1; q < end1; ++q) {
        if (ic1[q] < ccut) {
          const ulonglong2* rr = (const ulonglong2*)(Rp + (size_t)is1[q] * CP);
#pragma unroll
          for (int v = 0; v < CP / 4; ++v) {
            const ulonglong2 w = rr[v];
            ag[2 * v] = add2(ag[2 * v], w.x); ag[2 * v + 1] = add2(ag[2 * v + 1], w.y);
          }
          cnt += 1;
        }
      }
    }
    {
      const float inv = cnt > 0 ? (float)cnt : 1.f;
      const float rcp = __frcp_rn(inv);
#pragma unroll
      for (int c2 = 0; c2 < (CIN + 1) / 2; ++c2) {
        float a0, a1;
        upk2(ag[c2], a0, a1);
        // agg / inv: q = a r, e = fma(-q, inv, a), q += e r - the correction step of an IEEE division
        const float q0 = a0 * rcp, q1 = a1 * rcp;
        a0 = fmaf(fmaf(-q0, inv, a0), rcp, q0);
        a1 = fmaf(fmaf(-q1, inv, a1), rcp, q1);
        // ---- (iv) x1 += Wl agg
        GNN_FMA_ROW(xacc, O::wl, 2 * c2, a0);
        if (2 * c2 + 1 < CIN) GNN_FMA_ROW(xacc, O::wl, 2 * c2 + 1, a1);
      }
    }
    float x1[kF];
    float sc = c_gnn[O::cb];
#pragma unroll
    for (int pp = 0; pp < 8; ++pp) upk2(xacc[pp], x1[2 * pp], x1[2 * pp + 1]);
#pragma unroll
    for (int o = 0; o < kF; ++o) sc = fmaf(c_gnn[O::wc + o], x1[o], sc);
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


// ------------------------------------------------------------------------------------------
// Backward of gnn_layer_kernel (training).  The reference differentiates SAGEConv's mean aggregation,
// lin_l / lin_r, n_self_func and the classifier with autograd (src/model/gnn.py:207-218).  Here, per
// (pair, j2) CTA, the forward is recomputed from the layer inputs (nothing is saved but the inputs), then
// every node produces
//   * its direct input gradient  d_own = Wr^T dx1 + W0^T dh0 (+ d_agg / cnt on the self loop),
//   * gagg[p, :] = Wl^T dx1 / cnt  - to be pushed to the in-neighbours by assoc_aggregate_add_kernel,
//   * its outer-product contributions to the 9 weight gradients, reduced over the CTA's nodes in shared
//     memory and added to `grads` with one atomic per weight entry per CTA.
// grads layout (floats): Wl[16*CIN] bl[16] Wr[16*CIN] W0[16*CIN] b0[16] W2[256] b2[16] wc[16] cb[1].
// dxout: gradient wrt xout [B,N,16]; dscore: gradient wrt score [B,n1max,n2max].
// dxprev [B,N,16] (CIN == 17) and dm [B,n1max,n2max] receive d_own (overwritten, not accumulated).
// ------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(128, 2)
gnn_layer_bwd_kernel(const float* __restrict__ xprev, const float* __restrict__ mprev_t,
                     const int* __restrict__ in_ptr1, const int* __restrict__ in_src1,
                     const int* __restrict__ in_col1,
                     const int* __restrict__ in_ptr2, const int* __restrict__ in_src2,
                     const int64_t* __restrict__ ndiag_p, const int* __restrict__ part,
                     const float* __restrict__ dxout, const float* __restrict__ dscore,
                     float* __restrict__ dxprev, float* __restrict__ dm, float* __restrict__ gagg,
                     float* __restrict__ grads, int n1max, int n2max, int e1max, int e2max) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  constexpr int LV = 49;                       // dx1[16] ds2[16] dh0[16] dsc
  constexpr int RV = 2 * CP + 33;              // agg[CP] own[CP] h0[16] x1[16] one
  constexpr int VS = LV + RV;
  constexpr int NQ = 3 * kF * CIN + kF * kF + 4 * kF + 1;
  constexpr int QPT = (NQ + 127) / 128;
  extern __shared__ __align__(16) float sm[];
  float* Rsum = sm;                               // [n1max][CP]
  using O = GnnOff<CP>;                             // weights: constant bank c_gnn (packed by the host wrapper)
  const int b = blockIdx.y, j2 = blockIdx.x;
  const int N = n1max * n2max;
  const int tid = threadIdx.x;
  float* Rp = sm + (size_t)n1max * CP;             // [n1max][CP] row of the cut-off block
  float* V = sm + (size_t)2 * n1max * CP;          // [128][VS]
  short* qlo = (short*)(V + 128 * VS);   // [NQ]
  short* qro = qlo + NQ;
  // weight-gradient task table: entry q = <left vector component, right vector component>
  for (int q = tid; q < NQ; q += blockDim.x) {
    int lo, ro, r = q;
    constexpr int ONE = 2 * CP + 32;
    if (r < kF * CIN) { lo = r / CIN; ro = r % CIN; }                                   // Wl: dx1 x agg
    else if ((r -= kF * CIN) < kF) { lo = r; ro = ONE; }                                // bl
    else if ((r -= kF) < kF * CIN) { lo = r / CIN; ro = CP + r % CIN; }                 // Wr: dx1 x own
    else if ((r -= kF * CIN) < kF * CIN) { lo = 32 + r / CIN; ro = CP + r % CIN; }      // W0: dh0 x own
    else if ((r -= kF * CIN) < kF) { lo = 32 + r; ro = ONE; }                           // b0
    else if ((r -= kF) < kF * kF) { lo = 16 + r / kF; ro = 2 * CP + r % kF; }           // W2: ds2 x h0
    else if ((r -= kF * kF) < kF) { lo = 16 + r; ro = ONE; }                            // b2
    else if ((r -= kF) < kF) { lo = 48; ro = 2 * CP + 16 + r; }                         // wc: dsc x x1
    else { lo = 48; ro = ONE; }                                                         // cb
    qlo[q] = (short)lo; qro[q] = (short)(LV + ro);
  }

  // ---- stage 1 (as the forward): Rsum[i1, c] = sum_{i2 in In2(j2)} feat[(i2, i1), c]
  const int* ip2 = in_ptr2 + (size_t)b * (n2max + 1);
  const int beg2 = ip2[j2], end2 = ip2[j2 + 1];
  const int* is2 = in_src2 + (size_t)b * e2max;
  const float* xb = (CIN > 1) ? xprev + (size_t)b * N * kF : nullptr;
  const float* mb = mprev_t + (size_t)b * N;
  int ps2 = -1, ccut = 0;
  if (part != nullptr && part[b * 4 + 1] == j2) { ps2 = part[b * 4]; ccut = part[b * 4 + 2]; }
  if (ccut <= 0) { ps2 = -1; ccut = 0; }
  gnn_stage1<CIN>(xb, mb, is2, beg2, end2, ps2, Rsum, Rp, n1max);
  __syncthreads();

  const long long ndiag = ndiag_p[b];
  const int* ip1 = in_ptr1 + (size_t)b * (n1max + 1);
  const int* is1 = in_src1 + (size_t)b * e1max;
  const int* ic1 = in_col1 != nullptr ? in_col1 + (size_t)b * e1max : nullptr;
  const int d2 = end2 - beg2;
  float wacc[QPT];
#pragma unroll
  for (int u = 0; u < QPT; ++u) wacc[u] = 0.f;

  for (int j0 = 0; j0 < n1max; j0 += blockDim.x) {
    const int j1 = j0 + tid;
    float* v = V + (size_t)tid * VS;
    if (j1 < n1max) {
      const size_t p = (size_t)j2 * n1max + j1;
      float own[CP], agg[CP];
      bool self;
      const float inv = gnn_node_inputs<CIN>(xb, mb, Rsum, Rp, ip1, is1, ic1, j1, p, d2, ccut, ndiag, own, agg, self);

      // forward recompute: h0 = relu(W0 own + b0), s2 = W2 h0 + b2, x1 = Wl agg + bl + Wr own + relu(s2)
      float h0[kF], s2[kF], x1[kF];
#pragma unroll
      for (int o = 0; o < kF; ++o) {
        float a = c_gnn[O::b0 + o];
#pragma unroll
        for (int c = 0; c < CP; ++c) a = fmaf(GNN_W0(o, c), own[c], a);
        h0[o] = fmaxf(a, 0.f);
      }
#pragma unroll
      for (int o = 0; o < kF; ++o) {
        float a = c_gnn[O::bl + o], r = 0.f, q2 = c_gnn[O::b2 + o];
#pragma unroll
        for (int c = 0; c < CP; ++c) { a = fmaf(GNN_WL(o, c), agg[c], a); r = fmaf(GNN_WR(o, c), own[c], r); }
#pragma unroll
        for (int c = 0; c < kF; ++c) q2 = fmaf(GNN_W2(o, c), h0[c], q2);
        s2[o] = q2;
        x1[o] = (a + r) + fmaxf(q2, 0.f);
      }
      // incoming gradients
      const float dsc = dscore[((size_t)b * n1max + j1) * n2max + j2];
      float dx1[kF];
      {
        const float4* gp = (const float4*)(dxout + ((size_t)b * N + p) * kF);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 q4 = gp[t];
          dx1[t * 4] = q4.x; dx1[t * 4 + 1] = q4.y; dx1[t * 4 + 2] = q4.z; dx1[t * 4 + 3] = q4.w;
        }
      }
#pragma unroll
      for (int o = 0; o < kF; ++o) dx1[o] = fmaf(dsc, c_gnn[O::wc + o], dx1[o]);
      float ds2[kF], dh0[kF];
#pragma unroll
      for (int o = 0; o < kF; ++o) ds2[o] = s2[o] > 0.f ? dx1[o] : 0.f;
#pragma unroll
      for (int c = 0; c < kF; ++c) {
        float a = 0.f;
#pragma unroll
        for (int o = 0; o < kF; ++o) a = fmaf(GNN_W2(o, c), ds2[o], a);
        dh0[c] = h0[c] > 0.f ? a : 0.f;
      }
      float down[CP], dagg[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        float a = 0.f, g = 0.f;
#pragma unroll
        for (int o = 0; o < kF; ++o) {
          a = fmaf(GNN_WR(o, c), dx1[o], a);
          a = fmaf(GNN_W0(o, c), dh0[o], a);
          g = fmaf(GNN_WL(o, c), dx1[o], g);
        }
        g = g / inv;
        dagg[c] = g;
        down[c] = self ? a + g : a;
      }
      // outputs
      float4* gd = (float4*)(gagg + ((size_t)b * N + p) * CP);
#pragma unroll
      for (int t = 0; t < CP / 4; ++t) gd[t] = make_float4(dagg[t * 4], dagg[t * 4 + 1], dagg[t * 4 + 2], dagg[t * 4 + 3]);
      if (CIN > 1) {
        float4* dd = (float4*)(dxprev + ((size_t)b * N + p) * kF);
#pragma unroll
        for (int t = 0; t < 4; ++t) dd[t] = make_float4(down[t * 4], down[t * 4 + 1], down[t * 4 + 2], down[t * 4 + 3]);
      }
      dm[((size_t)b * n1max + j1) * n2max + j2] = down[CIN - 1];
      // per-node vectors for the weight-gradient reduction
#pragma unroll
      for (int o = 0; o < kF; ++o) { v[o] = dx1[o]; v[16 + o] = ds2[o]; v[32 + o] = dh0[o]; }
      v[48] = dsc;
#pragma unroll
      for (int c = 0; c < CP; ++c) { v[LV + c] = agg[c]; v[LV + CP + c] = own[c]; }
#pragma unroll
      for (int o = 0; o < kF; ++o) { v[LV + 2 * CP + o] = h0[o]; v[LV + 2 * CP + 16 + o] = x1[o]; }
      v[LV + 2 * CP + 32] = 1.f;
    }
    __syncthreads();
    const int nn = min((int)blockDim.x, n1max - j0);
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
      const int q = tid + u * 128;
      if (q < NQ) {
        const int lo = qlo[q], ro = qro[q];
        float a = 0.f;
        for (int nd = 0; nd < nn; ++nd) a = fmaf(V[(size_t)nd * VS + lo], V[(size_t)nd * VS + ro], a);
        wacc[u] += a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < QPT; ++u) {
    const int q = tid + u * 128;
    if (q < NQ && wacc[u] != 0.f) atomicAdd(grads + q, wacc[u]);
  }
}

// Pushes gagg (the mean-aggregation gradient of every destination node, already divided by its count) back to
// the sources:  d_feat[(i2,i1), c] += sum_{j2 in Out2(i2)} sum_{j1 in Out1(i1)} gagg[(j2,j1), c]  - the same
// factorised two-stage sum as the forward, over the OUT-neighbour lists.  Channels 0..15 go to dxprev
// (CIN == 17), channel CIN-1 to dm [B,n1max,n2max].
template <int CIN>
__global__ void __launch_bounds__(128, 4)
assoc_aggregate_add_kernel(const float* __restrict__ gagg, const int* __restrict__ out_ptr1,
                           const int* __restrict__ out_dst1, const int* __restrict__ out_col1,
                           const int* __restrict__ out_ptr2,
                           const int* __restrict__ out_dst2, const int* __restrict__ part,
                           float* __restrict__ dxprev,
                           float* __restrict__ dm, int n1max, int n2max, int e1max, int e2max) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  extern __shared__ __align__(16) float Rs[];       // [n1max][CP], then [n1max][CP] for the cut-off block
  const int b = blockIdx.y, i2 = blockIdx.x, tid = threadIdx.x;
  // the cut-off block (assoc_effective_kernel) is the association edge set (ps2, src1[c]) -> (pd2, dst1[c]), c < ccut:
  // its gradient flows from row pd2 of gagg to the sources in row ps2
  int pd2 = -1, ccut = 0;
  if (part != nullptr && part[b * 4] == (int)blockIdx.x) { pd2 = part[b * 4 + 1]; ccut = part[b * 4 + 2]; }
  if (ccut <= 0) pd2 = -1;
  float* Rq = Rs + (size_t)n1max * CP;
  const int N = n1max * n2max;
  const int* op2 = out_ptr2 + (size_t)b * (n2max + 1);
  const int beg2 = op2[i2], end2 = op2[i2 + 1];
  const int* od2 = out_dst2 + (size_t)b * e2max;
  const float* gb = gagg + (size_t)b * N * CP;
  const int nvec = n1max * (CP / 4);
  for (int f = tid; f < nvec; f += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = beg2; q < end2; ++q) {
      const float4 v = ((const float4*)(gb + (size_t)od2[q] * n1max * CP))[f];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    ((float4*)Rs)[f] = acc;
    if (pd2 >= 0) ((float4*)Rq)[f] = ((const float4*)(gb + (size_t)pd2 * n1max * CP))[f];
  }
  __syncthreads();
  const int* op1 = out_ptr1 + (size_t)b * (n1max + 1);
  const int* od1 = out_dst1 + (size_t)b * e1max;
  const int* oc1 = out_col1 != nullptr ? out_col1 + (size_t)b * e1max : nullptr;
  for (int i1 = tid; i1 < n1max; i1 += blockDim.x) {
    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = 0.f;
    for (int q = op1[i1]; q < op1[i1 + 1]; ++q) {
      const float4* r = (const float4*)(Rs + (size_t)od1[q] * CP);
#pragma unroll
      for (int t = 0; t < CP / 4; ++t) {
        const float4 v = r[t];
        acc[t * 4] += v.x; acc[t * 4 + 1] += v.y; acc[t * 4 + 2] += v.z; acc[t * 4 + 3] += v.w;
      }
    }
    if (pd2 >= 0) {
      for (int q = op1[i1]; q < op1[i1 + 1]; ++q) {
        if (oc1[q] < ccut) {
          const float4* r = (const float4*)(Rq + (size_t)od1[q] * CP);
#pragma unroll
          for (int t = 0; t < CP / 4; ++t) {
            const float4 v = r[t];
            acc[t * 4] += v.x; acc[t * 4 + 1] += v.y; acc[t * 4 + 2] += v.z; acc[t * 4 + 3] += v.w;
          }
        }
      }
    }
    const size_t p = (size_t)i2 * n1max + i1;
    if (CIN > 1) {
      float4* dd = (float4*)(dxprev + ((size_t)b * N + p) * kF);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float4 v = dd[t];
        v.x += acc[t * 4]; v.y += acc[t * 4 + 1]; v.z += acc[t * 4 + 2]; v.w += acc[t * 4 + 3];
        dd[t] = v;
      }
    }
    dm[((size_t)b * n1max + i1) * n2max + i2] += acc[CIN - 1];
  }
}

}  // namespace fpm

// weights -> padded staging buffer (device) -> constant bank c_gnn, all stream ordered (see the comment at c_gnn).
// The bank and the staging buffer exist once per device, so a (publish, layer kernel) sequence must not interleave
// with one issued on ANOTHER stream: gnn_bank_acquire makes the caller's stream wait for the last layer kernel that
// read the bank from a different stream, gnn_bank_release records that kernel's completion event.  Launches from
// several host threads are serialised by the mutex for the duration of the enqueue.
#include <mutex>
namespace {
struct GnnBank {
  std::mutex mu;
  float* staging = nullptr;
  cudaEvent_t last_use = nullptr;
  cudaStream_t last_stream = nullptr;
  bool used = false;
  bool capturing = false;
};
GnnBank g_bank[16];
}  // namespace

static int gnn_bank_acquire(int* dev_out, cudaStream_t st) {
  int dev = 0;
  FPM_CUDA(cudaGetDevice(&dev));
  FPM_CHECK_ARG(dev >= 0 && dev < 16, "fpm_gnn_layer: device index out of range");
  GnnBank& k = g_bank[dev];
  k.mu.lock();
  *dev_out = dev;
  cudaError_t e = cudaSuccess;
  if (!k.staging) e = cudaMalloc(&k.staging, fpm::kGnnConstFloats * sizeof(float));
  if (e == cudaSuccess && !k.last_use) e = cudaEventCreateWithFlags(&k.last_use, cudaEventDisableTiming);
  // A stream that is being captured into a CUDA graph must not wait on (or record) events of eager work: the graph
  // orders its own launches, and its replays must not overlap eager forwards of other streams on this device.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (e == cudaSuccess) e = cudaStreamIsCapturing(st, &cap);
  k.capturing = cap != cudaStreamCaptureStatusNone;
  if (e == cudaSuccess && !k.capturing && k.used && k.last_stream != st) e = cudaStreamWaitEvent(st, k.last_use, 0);
  if (e != cudaSuccess) {
    k.mu.unlock();
    fpm_set_error(cudaGetErrorString(e));
    return (int)e;
  }
  return FPM_OK;
}

static int gnn_bank_release(int dev, cudaStream_t st, int rc) {
  GnnBank& k = g_bank[dev];
  if (rc == FPM_OK && !k.capturing) {
    cudaError_t e = cudaEventRecord(k.last_use, st);
    if (e == cudaSuccess) { k.used = true; k.last_stream = st; }
    else { fpm_set_error(cudaGetErrorString(e)); rc = (int)e; }
  }
  k.mu.unlock();
  return rc;
}

template <int CIN>
static int gnn_publish_weights(const fpm::GnnWeights& w, int dev, cudaStream_t st) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  float* staging = g_bank[dev].staging;
  fpm::gnn_pack_weights_kernel<CIN><<<1, 256, 0, st>>>(w, staging);
  FPM_LAUNCH_CHECK();
  FPM_CUDA(cudaMemcpyToSymbolAsync(fpm::c_gnn, staging, fpm::GnnOff<CP>::total * sizeof(float), 0,
                                   cudaMemcpyDeviceToDevice, st));
  return FPM_OK;
}

extern "C" int fpm_assoc_effective(const int* edges1, const int* edges2, const long long* eptr1,
                                   const long long* eptr2, const long long* n1, const long long* n2, int* eff1,
                                   int* eff2, long long* ndiag, int* part, int* status, int B, int e1max, int e2max,
                                   void* stream) {
  FPM_CHECK_ARG(edges1 && edges2 && eptr1 && eptr2 && n1 && n2 && eff1 && eff2 && ndiag && part && status,
                "fpm_assoc_effective: null tensor");
  FPM_CHECK_ARG(B >= 0 && e1max >= 0 && e2max >= 0, "fpm_assoc_effective: bad sizes");
  if (B == 0) return FPM_OK;
  const size_t smem = (size_t)(2 * e1max + 2 * e2max + 4) * sizeof(int);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_assoc_effective: graphs too large for one CTA");
  FPM_CUDA(cudaFuncSetAttribute(fpm::assoc_effective_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fpm::assoc_effective_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(
      edges1, edges2, (const int64_t*)eptr1, (const int64_t*)eptr2, (const int64_t*)n1, (const int64_t*)n2, eff1, eff2,
      (int64_t*)ndiag, part, status, e1max, e2max);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_assoc_in_csr(const int* edges, int* in_ptr, int* in_src, int* in_col, int B, int nmax, int emax,
                                void* stream) {
  FPM_CHECK_ARG(edges && in_ptr && in_src, "fpm_assoc_in_csr: null tensor");
  FPM_CHECK_ARG(B >= 0 && nmax > 0 && emax >= 0, "fpm_assoc_in_csr: bad sizes");
  if (B == 0) return FPM_OK;
  const size_t smem = (size_t)(3 * emax + 2 * (nmax + 1)) * sizeof(int);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_assoc_in_csr: graph too large for one CTA");
  FPM_CUDA(cudaFuncSetAttribute(fpm::assoc_in_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  fpm::assoc_in_csr_kernel<<<B, fpm::kAssocThreads, smem, (cudaStream_t)stream>>>(edges, in_ptr, in_src, in_col, nmax,
                                                                                  emax);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

template <int CIN>
static int gnn_layer_launch(const fpm::GnnWeights& w, int dev, const float* xprev, const float* mprev_t,
                            const int* in_ptr1, const int* in_src1, const int* in_col1, const int* in_ptr2,
                            const int* in_src2, const long long* ndiag, const int* part, float* xout, float* score,
                            int B, int n1max, int n2max, int e1max, int e2max, cudaStream_t st) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  // rows of association nodes (consecutive j2) per CTA: as many as fit 256 threads, so that whole warps are busy
  int rows = fpm::kGnnMaxThreads / n1max;
  rows = rows < 1 ? 1 : (rows > 4 ? 4 : rows);
  if (rows > n2max) rows = n2max;
  int threads = (rows * n1max + 31) / 32 * 32;
  if (threads > fpm::kGnnMaxThreads) threads = fpm::kGnnMaxThreads;
  const size_t smem = ((size_t)(rows + 1) * n1max * CP + 8) * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_gnn_layer: n1max too large");
  int rc = gnn_publish_weights<CIN>(w, dev, st);
  if (rc != FPM_OK) return rc;
  static int minb = 0;                                 // FPMATCH_GNN_MINB=3|4 (A/B switch; default 4)
  if (minb == 0) {
    const char* e = getenv("FPMATCH_GNN_MINB");
    minb = (e && e[0] == '3') ? 3 : 4;
  }
  dim3 grid(fpm_cdiv(n2max, rows), B);
  if (minb == 4) {
    FPM_CUDA(cudaFuncSetAttribute(fpm::gnn_layer_kernel<CIN, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fpm::gnn_layer_kernel<CIN, 4><<<grid, threads, smem, st>>>(xprev, mprev_t, in_ptr1, in_src1, in_col1, in_ptr2,
                                                              in_src2, (const int64_t*)ndiag, part, xout, score, n1max,
                                                              n2max, e1max, e2max, rows);
  } else {
    FPM_CUDA(cudaFuncSetAttribute(fpm::gnn_layer_kernel<CIN, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fpm::gnn_layer_kernel<CIN, 3><<<grid, threads, smem, st>>>(xprev, mprev_t, in_ptr1, in_src1, in_col1, in_ptr2,
                                                              in_src2, (const int64_t*)ndiag, part, xout, score, n1max,
                                                              n2max, e1max, e2max, rows);
  }
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// weights: 9 device pointers in the order of GnnWeights.  ndiag [B] int64 and part [B,4] int32 come from
// fpm_assoc_effective (part may be NULL = no cut-off block; in_col1 is only read for pairs that have one).
extern "C" int fpm_gnn_layer(const float* xprev, const float* mprev_t, const int* in_ptr1,
                             const int* in_src1, const int* in_col1, const int* in_ptr2, const int* in_src2,
                             const long long* ndiag, const int* part, const float* const* weights,
                             float* xout, float* score, int B, int n1max, int n2max, int e1max, int e2max,
                             int cin, void* stream) {
  FPM_CHECK_ARG(mprev_t && in_ptr1 && in_src1 && in_ptr2 && in_src2 && ndiag && weights && xout && score,
                "fpm_gnn_layer: null tensor");
  FPM_CHECK_ARG(part == nullptr || in_col1 != nullptr, "fpm_gnn_layer: part needs in_col1");
  FPM_CHECK_ARG(cin == 1 || (cin == 17 && xprev), "fpm_gnn_layer: cin must be 1 or 17 (with xprev)");
  FPM_CHECK_ARG(B >= 0 && n1max > 0 && n2max > 0, "fpm_gnn_layer: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_gnn_layer: batch too large");
  for (int i = 0; i < 9; ++i) FPM_CHECK_ARG(weights[i], "fpm_gnn_layer: null weight");
  fpm::GnnWeights w{weights[0], weights[1], weights[2], weights[3], weights[4],
                    weights[5], weights[6], weights[7], weights[8]};
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  int rc = gnn_bank_acquire(&dev, st);
  if (rc != FPM_OK) return rc;
  if (cin == 1)
    rc = gnn_layer_launch<1>(w, dev, xprev, mprev_t, in_ptr1, in_src1, in_col1, in_ptr2, in_src2, ndiag, part, xout,
                             score, B, n1max, n2max, e1max, e2max, st);
  else
    rc = gnn_layer_launch<17>(w, dev, xprev, mprev_t, in_ptr1, in_src1, in_col1, in_ptr2, in_src2, ndiag, part, xout,
                              score, B, n1max, n2max, e1max, e2max, st);
  return gnn_bank_release(dev, st, rc);
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

// Backward of fpm_gnn_layer.  out_ptr*/out_dst*/out_col1: OUT-neighbour lists (fpm_assoc_in_csr on the effective edge
// tables with the two rows swapped).  gagg: scratch [B, N, CP] floats (CP = 4 for cin 1, 20 for cin 17).  grads: see
// the kernel comment; the caller zeroes it.  dxprev may be NULL for cin = 1.
template <int CIN>
static int gnn_layer_bwd_launch(const fpm::GnnWeights& w, int dev, const float* xprev, const float* mprev_t,
                                const int* in_ptr1, const int* in_src1, const int* in_col1, const int* in_ptr2,
                                const int* in_src2, const int* out_ptr1, const int* out_dst1, const int* out_col1,
                                const int* out_ptr2, const int* out_dst2, const long long* ndiag, const int* part,
                                const float* dxout, const float* dscore, float* dxprev, float* dm, float* gagg,
                                float* grads, int B, int n1max, int n2max, int e1max, int e2max, cudaStream_t st) {
  constexpr int CP = (CIN + 3) / 4 * 4;
  constexpr int NQ = 3 * 16 * CIN + 16 * 16 + 4 * 16 + 1;
  constexpr int VS = 49 + 2 * CP + 33;
  const size_t smem = ((size_t)2 * n1max * CP + (size_t)128 * VS) * sizeof(float) + (size_t)2 * NQ * sizeof(short) + 16;
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_gnn_layer_bwd: n1max too large");
  const size_t smem2 = (size_t)2 * n1max * CP * sizeof(float);
  dim3 grid(n2max, B);
  FPM_CUDA(cudaFuncSetAttribute(fpm::gnn_layer_bwd_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int rc = gnn_publish_weights<CIN>(w, dev, st);
  if (rc != FPM_OK) return rc;
  fpm::gnn_layer_bwd_kernel<CIN><<<grid, 128, smem, st>>>(xprev, mprev_t, in_ptr1, in_src1, in_col1, in_ptr2, in_src2,
                                                         (const int64_t*)ndiag, part, dxout, dscore, dxprev, dm, gagg,
                                                         grads, n1max, n2max, e1max, e2max);
  FPM_LAUNCH_CHECK();
  FPM_CUDA(cudaFuncSetAttribute(fpm::assoc_aggregate_add_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem2));
  fpm::assoc_aggregate_add_kernel<CIN><<<grid, 128, smem2, st>>>(gagg, out_ptr1, out_dst1, out_col1, out_ptr2, out_dst2,
                                                                part, dxprev, dm, n1max, n2max, e1max, e2max);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_gnn_layer_bwd(const float* xprev, const float* mprev_t, const int* in_ptr1, const int* in_src1,
                                 const int* in_col1, const int* in_ptr2, const int* in_src2, const int* out_ptr1,
                                 const int* out_dst1, const int* out_col1, const int* out_ptr2, const int* out_dst2,
                                 const long long* ndiag, const int* part, const float* const* weights,
                                 const float* dxout, const float* dscore, float* dxprev, float* dm, float* gagg,
                                 float* grads, int B, int n1max, int n2max, int e1max, int e2max, int cin,
                                 void* stream) {
  FPM_CHECK_ARG(mprev_t && in_ptr1 && in_src1 && in_ptr2 && in_src2 && out_ptr1 && out_dst1 && out_ptr2 && out_dst2 &&
                ndiag && weights && dxout && dscore && dm && gagg && grads, "fpm_gnn_layer_bwd: null tensor");
  FPM_CHECK_ARG(part == nullptr || (in_col1 != nullptr && out_col1 != nullptr),
                "fpm_gnn_layer_bwd: part needs in_col1 and out_col1");
  FPM_CHECK_ARG(cin == 1 || (cin == 17 && xprev && dxprev), "fpm_gnn_layer_bwd: cin must be 1 or 17 (with xprev, dxprev)");
  FPM_CHECK_ARG(B >= 0 && n1max > 0 && n2max > 0, "fpm_gnn_layer_bwd: bad sizes");
  if (B == 0) return FPM_OK;
  FPM_CHECK_ARG(B <= 65535, "fpm_gnn_layer_bwd: batch too large");
  for (int i = 0; i < 9; ++i) FPM_CHECK_ARG(weights[i], "fpm_gnn_layer_bwd: null weight");
  fpm::GnnWeights w{weights[0], weights[1], weights[2], weights[3], weights[4],
                    weights[5], weights[6], weights[7], weights[8]};
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  int rc = gnn_bank_acquire(&dev, st);
  if (rc != FPM_OK) return rc;
  if (cin == 1)
    rc = gnn_layer_bwd_launch<1>(w, dev, xprev, mprev_t, in_ptr1, in_src1, in_col1, in_ptr2, in_src2, out_ptr1,
                                 out_dst1, out_col1, out_ptr2, out_dst2, ndiag, part, dxout, dscore, dxprev, dm, gagg,
                                 grads, B, n1max, n2max, e1max, e2max, st);
  else
    rc = gnn_layer_bwd_launch<17>(w, dev, xprev, mprev_t, in_ptr1, in_src1, in_col1, in_ptr2, in_src2, out_ptr1,
                                  out_dst1, out_col1, out_ptr2, out_dst2, ndiag, part, dxout, dscore, dxprev, dm, gagg,
                                  grads, B, n1max, n2max, e1max, e2max, st);
  return gnn_bank_release(dev, st, rc);
}
