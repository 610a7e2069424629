// SplineConv node refinement, restructured for dense tensor-core work.
//
// Replaces torch_geometric.nn.SplineConv(768, 768, dim=2, kernel_size=5, aggr='max') as used by
// /root/reference/src/model/spline_conv.py:7-58 (torch_spline_conv basis + weighting, torch_scatter max).
//
// The reference evaluates, per edge, sum_s basis_s * (x_src @ W[wi_s]): a gather-GEMM of shape
// [4e, 768] x [768, 768] with a different weight slab per row.  Because every source node is used by
// ~6 edges touching almost all 25 slabs, the same FLOPs buy a plain dense GEMM instead:
//     Y[j, k, :] = x_j @ W[k]    for all 25 kernels (+ slab 25 = root weight)        -> gemm_*.cu
// and the per-edge work collapses to a 4-row blend of Y plus a max over in-edges (this file):
//     out_i = max_{e: j->i} sum_s basis_s(e) * Y[j, wi_s(e), :]  (0 if no in-edge) + Y[i, 25, :] + bias
// fused with ReLU (layer 0) or the residual x + 0.1 * out (layer 1, spline_conv.py:56).
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
#include <limits.h>

namespace fpm {

// In-edge lists of a PyG-style batch graph: for every node, the ids of the edges that end in it, in
// ascending edge order (deterministic).  One CTA per pair, a counting sort in O(e): count the in-degrees with atomics
// on this CTA's slice of in_ptr, scan them, scatter the edge ids (atomic cursors, arbitrary order inside a node),
// then one thread per node sorts its short list - the order the atomics produced never reaches the output.
// (Before: one warp per destination node scanned ALL the pair's edges twice, O(n e): 22 us per launch at 100
// keypoints; the first version was a single thread per destination, 0.24 ms at 400 keypoints.)
// in_ptr has [total_nodes + 1] entries (global offsets into in_eid).
constexpr int kCsrThreads = 512;
__global__ void __launch_bounds__(kCsrThreads)
csr_by_dst_kernel(const int64_t* __restrict__ edge_dst, const int64_t* __restrict__ ptr,
                  const int64_t* __restrict__ eptr, int* __restrict__ in_ptr, int* __restrict__ in_eid,
                  int total_nodes) {
  extern __shared__ int sm_csr[];                 // [e] local destinations, [e] scattered local edge ids
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = (int)ptr[b], n = (int)ptr[b + 1] - n0;
  const int e0 = (int)eptr[b], e = (int)eptr[b + 1] - e0;
  int* sdst = sm_csr;
  int* slist = sm_csr + e;
  int* cnt = in_ptr + n0;                         // counts -> offsets -> cursors live in the output (this CTA's rows)
  for (int j = threadIdx.x; j < n; j += blockDim.x) cnt[j] = 0;
  for (int k = threadIdx.x; k < e; k += blockDim.x) sdst[k] = (int)edge_dst[e0 + k] - n0;
  __syncthreads();
  for (int k = threadIdx.x; k < e; k += blockDim.x) atomicAdd(&cnt[sdst[k]], 1);
  __syncthreads();
  if (warp == 0) {                                // exclusive scan of the counts, 32 nodes at a time
    int base = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      const int c = j < n ? cnt[j] : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (j < n) cnt[j] = base + inc - c;         // local start of node j
      base += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < e; k += blockDim.x) slist[atomicAdd(&cnt[sdst[k]], 1)] = k;
  __syncthreads();
  // cnt[j] is now the END of node j's list = the start of node j + 1: sort every list, then shift the offsets
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int beg = j ? cnt[j - 1] : 0, end = cnt[j];
    for (int a = beg + 1; a < end; ++a) {         // insertion sort: in-degrees are a handful
      const int key = slist[a];
      int q = a - 1;
      while (q >= beg && slist[q] > key) { slist[q + 1] = slist[q]; --q; }
      slist[q + 1] = key;
    }
  }
  __syncthreads();
  int keep[4];                                    // n <= 4 * blockDim.x nodes per graph in registers, else a loop below
  const bool small = n <= 4 * (int)blockDim.x;
  if (small) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = threadIdx.x + u * blockDim.x;
      keep[u] = (j < n && j > 0) ? cnt[j - 1] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = threadIdx.x + u * blockDim.x;
      if (j < n) cnt[j] = e0 + keep[u];
    }
  } else if (threadIdx.x == 0) {                  // huge graphs: serial shift from the top
    for (int j = n - 1; j > 0; --j) cnt[j] = e0 + cnt[j - 1];
    cnt[0] = e0;
  }
  for (int k = threadIdx.x; k < e; k += blockDim.x) in_eid[e0 + k] = e0 + slist[k];
  if (b == (int)gridDim.x - 1 && threadIdx.x == 0) in_ptr[total_nodes] = e0 + e;
}

// open B-spline basis, degree 1, 2-d pseudo coordinates, kernel 5x5 (torch_spline_conv basis_cpu.cpp)
__device__ __forceinline__ void spline_basis4(float u0, float u1, int KS, float* bas, int* wi) {
  const float v0 = u0 * (float)(KS - 1), v1 = u1 * (float)(KS - 1);
  const float fl0 = floorf(v0), fl1 = floorf(v1);
  const float f0 = v0 - fl0, f1 = v1 - fl1;
  const int i0 = (int)fl0, i1 = (int)fl1;
  const float g0 = (float)(1.0 - (double)f0), g1 = (float)(1.0 - (double)f1);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int k0 = s & 1, k1 = s >> 1;
    bas[s] = __fmul_rn(k0 ? f0 : g0, k1 ? f1 : g1);
    wi[s] = ((i0 + k0) % KS) + ((i1 + k1) % KS) * KS;
  }
}

// One CTA (192 threads x float4 = 768 channels) per destination node.
// Y: [total_nodes, NS, C] with NS = KS*KS + 1 slabs.  mode 0: out = relu(conv); mode 1: out = x + 0.1*conv.
// argmax (training only): [total_nodes, C] int32, the edge id that won the max per channel (-1: no in-edge).
// The in-edges are handled in chunks of 16: thread t < chunk first resolves edge t (edge id -> source node ->
// pseudo-coordinates -> the 4 basis weights and slab ids) into shared memory, ONCE per edge; the channel loop then
// only streams Y: 4 independent 128-bit loads per edge, two edges in flight per thread.  (The first version walked
// in_eid -> edge_src -> pseudo -> Y as a dependent chain per edge in every one of the 192 threads, each re-deriving
// the same basis with its floor / double-precision / modulo arithmetic: 53 % of the DRAM peak, FMA pipe 34 % busy.)
constexpr int kGatherChunk = 16;      // 576 bytes of shared memory: small enough to sit beside a GEMM CTA
__global__ void __launch_bounds__(192)
spline_gather_max_kernel(const float* __restrict__ Y, const float* __restrict__ xin,
                         const int64_t* __restrict__ edge_src, const float* __restrict__ pseudo,
                         const int* __restrict__ in_ptr, const int* __restrict__ in_eid,
                         const float* __restrict__ bias, float* __restrict__ out, int* __restrict__ argmax,
                         __half* __restrict__ out_hi, __half* __restrict__ out_lo, float* __restrict__ out_inv,
                         int C, int KS, int mode) {
  __shared__ int s_e[kGatherChunk];
  __shared__ unsigned s_off[kGatherChunk][4];  // float4 offset of the slab row inside Y: (j * NS + wi) * C / 4
  __shared__ float s_bas[kGatherChunk][4];
  const int i = blockIdx.x;
  const int NS = KS * KS + 1;
  const int e_beg = in_ptr[i], e_end = in_ptr[i + 1];
  const int c4 = threadIdx.x;                 // float4 index along channels
  const bool live = c4 * 4 < C;
  float4 best = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
  int4 arg = make_int4(-1, -1, -1, -1);
  // the root / bias / residual operands do not depend on the edges: issue their loads first
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f), bi = r, x0 = r;
  if (live) {
    r = *(const float4*)(Y + ((size_t)i * NS + (NS - 1)) * C + c4 * 4);
    bi = *(const float4*)(bias + c4 * 4);
    if (mode == 1) x0 = *(const float4*)(xin + (size_t)i * C + c4 * 4);
  }
  for (int q0 = e_beg; q0 < e_end; q0 += kGatherChunk) {
    const int nq = min(kGatherChunk, e_end - q0);
    if (q0 != e_beg) __syncthreads();
    if ((int)threadIdx.x < nq) {
      const int e = in_eid[q0 + threadIdx.x];
      const long long j = edge_src[e];
      float bas[4]; int wi[4];
      spline_basis4(pseudo[(size_t)e * 2], pseudo[(size_t)e * 2 + 1], KS, bas, wi);
      s_e[threadIdx.x] = e;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        s_bas[threadIdx.x][s] = bas[s];
        s_off[threadIdx.x][s] = (unsigned)((j * NS + wi[s]) * (C / 4));   // < 2^32 float4 (the wrapper checks)
      }
    }
    __syncthreads();
    if (live) {
      const float4* yc = (const float4*)Y + c4;
      int q = 0;
      for (; q + 1 < nq; q += 2) {
        float4 y[8];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          y[s] = yc[s_off[q][s]];
          y[4 + s] = yc[s_off[q + 1][s]];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const float w = s_bas[q + u][s];
            const float4 v = y[4 * u + s];
            m.x = fmaf(w, v.x, m.x); m.y = fmaf(w, v.y, m.y); m.z = fmaf(w, v.z, m.z); m.w = fmaf(w, v.w, m.w);
          }
          if (argmax) {                              // first maximum wins (ascending edge order)
            const int e = s_e[q + u];
            if (m.x > best.x) arg.x = e;
            if (m.y > best.y) arg.y = e;
            if (m.z > best.z) arg.z = e;
            if (m.w > best.w) arg.w = e;
          }
          best.x = fmaxf(best.x, m.x); best.y = fmaxf(best.y, m.y);
          best.z = fmaxf(best.z, m.z); best.w = fmaxf(best.w, m.w);
        }
      }
      if (q < nq) {
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float w = s_bas[q][s];
          const float4 v = yc[s_off[q][s]];
          m.x = fmaf(w, v.x, m.x); m.y = fmaf(w, v.y, m.y); m.z = fmaf(w, v.z, m.z); m.w = fmaf(w, v.w, m.w);
        }
        if (argmax) {
          const int e = s_e[q];
          if (m.x > best.x) arg.x = e;
          if (m.y > best.y) arg.y = e;
          if (m.z > best.z) arg.z = e;
          if (m.w > best.w) arg.w = e;
        }
        best.x = fmaxf(best.x, m.x); best.y = fmaxf(best.y, m.y);
        best.z = fmaxf(best.z, m.z); best.w = fmaxf(best.w, m.w);
      }
    }
  }
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    if (argmax) *(int4*)(argmax + (size_t)i * C + c4 * 4) = arg;
    if (e_beg == e_end) best = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = best.x + r.x + bi.x; v.y = best.y + r.y + bi.y;
    v.z = best.z + r.z + bi.z; v.w = best.w + r.w + bi.w;
    if (mode == 0) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    } else if (mode == 1) {
      v.x = x0.x + 0.1f * v.x; v.y = x0.y + 0.1f * v.y; v.z = x0.z + 0.1f * v.z; v.w = x0.w + 0.1f * v.w;
    }
    if (out) *(float4*)(out + (size_t)i * C + c4 * 4) = v;
  }
  if (out_hi) {
    // The row is the A operand of the next layer's slab GEMM: emit its error-compensated fp16 split here (the same
    // arithmetic as f16_split_rows_kernel, gemm_tcgen05.cu) instead of writing fp32 and re-reading it in a split pass.
    __shared__ float s_amax[8];
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    amax = warp_max(amax);
    __syncthreads();                                  // the edge chunks' shared arrays are not read any more
    if ((threadIdx.x & 31) == 0) s_amax[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = s_amax[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) amax = fmaxf(amax, s_amax[w]);
    int e = 0;
    if (amax > 0.f && amax < INFINITY) frexpf(amax, &e);
    e = max(-100, min(100, e));
    const float sc = ldexpf(1.f, -e);
    if (threadIdx.x == 0) out_inv[i] = ldexpf(1.f, e);
    if (live) {
      const float x[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
      __half hh[4], ll[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hh[j] = __float2half_rn(x[j]);
        ll[j] = __float2half_rn((x[j] - __half2float(hh[j])) * 2048.f);
      }
      __half2* h2 = (__half2*)(out_hi + (size_t)i * C + c4 * 4);
      __half2* l2 = (__half2*)(out_lo + (size_t)i * C + c4 * 4);
      h2[0] = __halves2half2(hh[0], hh[1]); h2[1] = __halves2half2(hh[2], hh[3]);
      l2[0] = __halves2half2(ll[0], ll[1]); l2[1] = __halves2half2(ll[2], ll[3]);
    }
  }
}

// Backward of the gather/max: dY[j, k, :] = sum over out-edges e = (j -> i) whose message won the max at
// (i, c) of basis_s(e) * G[i, c] for the 4 slabs k = wi_s(e), plus the root slab dY[j, KS*KS, :] = G[j, :].
// G is the gradient of the conv output before relu / residual scaling.  One CTA per SOURCE node; its
// [NS][C] block of dY is accumulated in shared memory over the out-edge list (ascending edge order ->
// deterministic) and written once, zeros included, so dY needs no memset.
__global__ void __launch_bounds__(192)
spline_scatter_bwd_kernel(const float* __restrict__ G, const int* __restrict__ argmax,
                          const int64_t* __restrict__ edge_dst, const float* __restrict__ pseudo,
                          const int* __restrict__ out_ptr, const int* __restrict__ out_eid,
                          float* __restrict__ dY, int C, int KS) {
  extern __shared__ __align__(16) float blk[];       // [NS][C]
  const int j = blockIdx.x;
  const int NS = KS * KS + 1;
  const int c4 = threadIdx.x;
  if (c4 * 4 >= C) return;                           // threads own disjoint channel quads: no syncs needed
  for (int k = 0; k < NS - 1; ++k) *(float4*)(blk + (size_t)k * C + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  *(float4*)(blk + (size_t)(NS - 1) * C + c4 * 4) = *(const float4*)(G + (size_t)j * C + c4 * 4);
  const int e_beg = out_ptr[j], e_end = out_ptr[j + 1];
  for (int q = e_beg; q < e_end; ++q) {
    const int e = out_eid[q];
    const int i = (int)edge_dst[e];
    const int4 a = *(const int4*)(argmax + (size_t)i * C + c4 * 4);
    if (a.x != e && a.y != e && a.z != e && a.w != e) continue;
    float4 g = *(const float4*)(G + (size_t)i * C + c4 * 4);
    if (a.x != e) g.x = 0.f;
    if (a.y != e) g.y = 0.f;
    if (a.z != e) g.z = 0.f;
    if (a.w != e) g.w = 0.f;
    float bas[4]; int wi[4];
    spline_basis4(pseudo[(size_t)e * 2], pseudo[(size_t)e * 2 + 1], KS, bas, wi);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      float4* d = (float4*)(blk + (size_t)wi[s] * C + c4 * 4);
      float4 v = *d;
      v.x = fmaf(bas[s], g.x, v.x); v.y = fmaf(bas[s], g.y, v.y);
      v.z = fmaf(bas[s], g.z, v.z); v.w = fmaf(bas[s], g.w, v.w);
      *d = v;
    }
  }
  float* dst = dY + (size_t)j * NS * C;
  for (int k = 0; k < NS; ++k)
    *(float4*)(dst + (size_t)k * C + c4 * 4) = *(const float4*)(blk + (size_t)k * C + c4 * 4);
}

// Column-compacted variant of the scatter (training with the slab plan).  dY is zero outside the (node, slab) blocks
// some edge reads, and on keypoint graphs most slabs are read by few nodes or none: writing all 26 slabs of every
// node (1 GB at 64 pairs) and feeding them to the dX / dW GEMMs spends ~60 % of the backward on zeros.  The caller
// splits the slabs into a WIDE group (read by many nodes, plus the root slab) and a NARROW group (read by the nodes
// listed in `rowpos`); colmap[k] = g >= 0: column block g of dYd [T, nD, C];  = -(g + 1) < 0: column block g of
// dYs [nR, nS, C], row rowpos[j];  = INT_MIN: no edge reads slab k (nothing to write).
__global__ void __launch_bounds__(192)
spline_scatter_bwd_compact_kernel(const float* __restrict__ G, const int* __restrict__ argmax,
                                  const int64_t* __restrict__ edge_dst, const float* __restrict__ pseudo,
                                  const int* __restrict__ out_ptr, const int* __restrict__ out_eid,
                                  const int* __restrict__ colmap, const int* __restrict__ rowpos,
                                  float* __restrict__ dYd, float* __restrict__ dYs, int C, int KS, int nD, int nS) {
  extern __shared__ __align__(16) float blk[];       // [NS][C]
  const int j = blockIdx.x;
  const int NS = KS * KS + 1;
  const int c4 = threadIdx.x;
  if (c4 * 4 >= C) return;
  for (int k = 0; k < NS - 1; ++k) *(float4*)(blk + (size_t)k * C + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  *(float4*)(blk + (size_t)(NS - 1) * C + c4 * 4) = *(const float4*)(G + (size_t)j * C + c4 * 4);
  const int e_beg = out_ptr[j], e_end = out_ptr[j + 1];
  for (int q = e_beg; q < e_end; ++q) {
    const int e = out_eid[q];
    const int i = (int)edge_dst[e];
    const int4 a = *(const int4*)(argmax + (size_t)i * C + c4 * 4);
    if (a.x != e && a.y != e && a.z != e && a.w != e) continue;
    float4 g = *(const float4*)(G + (size_t)i * C + c4 * 4);
    if (a.x != e) g.x = 0.f;
    if (a.y != e) g.y = 0.f;
    if (a.z != e) g.z = 0.f;
    if (a.w != e) g.w = 0.f;
    float bas[4]; int wi[4];
    spline_basis4(pseudo[(size_t)e * 2], pseudo[(size_t)e * 2 + 1], KS, bas, wi);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      float4* d = (float4*)(blk + (size_t)wi[s] * C + c4 * 4);
      float4 v = *d;
      v.x = fmaf(bas[s], g.x, v.x); v.y = fmaf(bas[s], g.y, v.y);
      v.z = fmaf(bas[s], g.z, v.z); v.w = fmaf(bas[s], g.w, v.w);
      *d = v;
    }
  }
  const int rp = rowpos ? rowpos[j] : -1;
  for (int k = 0; k < NS; ++k) {
    const int cm = colmap[k];
    const float4 v = *(const float4*)(blk + (size_t)k * C + c4 * 4);
    if (cm >= 0) *(float4*)(dYd + ((size_t)j * nD + cm) * C + c4 * 4) = v;
    else if (cm != INT_MIN && rp >= 0) *(float4*)(dYs + ((size_t)rp * nS + (-cm - 1)) * C + c4 * 4) = v;
  }
}

// ------------------------------------------------------------------------------------------
// Slab planner: which (node, weight slab) products does the gather actually read?
//
// An edge src -> dst touches the 4 slabs wi_s(pseudo) of its SOURCE node, and the pseudo-coordinates of a keypoint
// graph are 0.5 + (P_src - P_dst) / 640: Delaunay neighbours are close, so almost every edge lands in the 3 x 3
// centre of the 5 x 5 kernel (measured on the synthetic pairs: 8.6 of 25 slabs per node).  Computing all 26 slab
// products for all nodes - the first design - spends 63 % of the tensor-core work on blocks nobody reads.
// The planner (device side, no host round trip) marks the slabs every node needs, then emits a tile table for the
// persistent GEMM (gemm_tcgen05.cu, PairTile):
//   * a slab needed by >= 1/4 of the nodes is "dense": it is computed for ALL nodes straight from X (no gather);
//   * the other slabs are "sparse": their nodes are compacted into 256-row groups behind X in the A buffer and the
//     GEMM scatters the result rows back to Y[node, slab, :] through the row map.
// The products that are computed are the same numbers as before; the ones skipped are never read.
// ------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;
constexpr int kTileM = 256, kTileN = 128;             // PairTile geometry of gemm_tc_pair_kernel

__global__ void slab_mask_kernel(const int64_t* __restrict__ edge_src, const float* __restrict__ pseudo,
                                 unsigned* __restrict__ mask, int E, int KS) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float bas[4]; int wi[4];
  spline_basis4(pseudo[(size_t)e * 2], pseudo[(size_t)e * 2 + 1], KS, bas, wi);
  const unsigned m = (1u << wi[0]) | (1u << wi[1]) | (1u << wi[2]) | (1u << wi[3]);
  atomicOr(mask + (int)edge_src[e], m);
}

// meta (ints): [0] tile count, [1] sparse rows in use, [2 .. 2+NS) dense flag per slab, [2+NS .. 2+2NS] sparse row
// offset per slab (+ total), [2+2NS+1 .. 2+3NS+1) cursors, then 33 ints of scratch the caller zero-fills (per-slab node
// counts + a ticket).  The node counts are taken by all CTAs of the grid (one ballot per (32 nodes, slab), one global
// atomic per (warp, slab)); the LAST CTA to finish then lays out the tile table alone.  (First version: one CTA did
// everything, counting with one shared-memory atomic per (node, slab): 39 us at T = 25 600; with ballots in that one
// CTA still 57 us - 8 k instructions per warp on a single SM.)
__global__ void __launch_bounds__(kPlanThreads)
slab_plan_kernel(const unsigned* __restrict__ mask, int T, int KS, int C, int* __restrict__ meta,
                 int4* __restrict__ tab, int max_tiles) {
  const int NS = KS * KS + 1;
  __shared__ int cnt[32], dense[32], off[33], dlist[32], tstart[33];
  __shared__ int nd, total_sparse, is_last;
  const int tid = threadIdx.x;
  int* gcnt = meta + 2 + 3 * NS + 2;
  int* ticket = gcnt + 32;
  {
    const int lane = tid & 31;
    int mine = 0;
    for (int j0 = blockIdx.x * kPlanThreads + tid - lane; j0 < T; j0 += gridDim.x * kPlanThreads) {
      const unsigned m = (j0 + lane < T) ? mask[j0 + lane] : 0u;
      for (int k = 0; k < NS; ++k) {
        const int c = __popc(__ballot_sync(0xffffffffu, (m >> k) & 1u));
        if (lane == k) mine += c;
      }
    }
    if (lane < NS && mine) atomicAdd(&gcnt[lane], mine);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid < 32) cnt[tid] = tid < NS ? *((volatile int*)gcnt + tid) : 0;
  __syncthreads();
  if (tid == 0) {
    cnt[NS - 1] = T;                                  // the root slab serves every node
    int n = 0, run = 0, trun = 0;
    for (int k = 0; k < NS; ++k) {
      dense[k] = (k == NS - 1) || (4LL * cnt[k] >= (long long)T && cnt[k] > 0);
      if (dense[k]) dlist[n++] = k;
      off[k] = run;
      tstart[k] = trun;                               // first sparse tile of slab k (after the dense tiles)
      if (!dense[k]) {
        run += (cnt[k] + kTileM - 1) / kTileM * kTileM;
        trun += (cnt[k] + kTileM - 1) / kTileM * (C / kTileN);
      }
    }
    off[NS] = run;
    tstart[NS] = trun;
    nd = n; total_sparse = run;
  }
  __syncthreads();
  const int ntn = C / kTileN;                         // N tiles per slab
  const int tiles_m = (T + kTileM - 1) / kTileM, T_pad = tiles_m * kTileM;
  const int grp_rows = 16;
  const int per_group = grp_rows * nd * ntn;
  const int total_dense = tiles_m * nd * ntn;
  for (int t = tid; t < total_dense && t < max_tiles; t += kPlanThreads) {
    const int grp = t / per_group, rem = t - grp * per_group;
    const int gsize = min(grp_rows, tiles_m - grp * grp_rows);
    const int mi = grp * grp_rows + rem % gsize, sn = rem / gsize;
    const int k = dlist[sn / ntn], nt = sn % ntn;
    tab[t] = make_int4(mi * kTileM, k * C + nt * kTileN, k * C + nt * kTileN, -1);
  }
  for (int k = 0; k < NS; ++k) {
    if (dense[k] || cnt[k] == 0) continue;
    const int mt = (cnt[k] + kTileM - 1) / kTileM;
    for (int i = tid; i < mt * ntn; i += kPlanThreads) {
      const int nt = i / mt, m = i - nt * mt, t = total_dense + tstart[k] + i;
      if (t < max_tiles)
        tab[t] = make_int4(T_pad + off[k] + m * kTileM, k * C + nt * kTileN, k * C + nt * kTileN, off[k] + m * kTileM);
    }
  }
  if (tid == 0) {
    meta[0] = min(total_dense + tstart[NS], max_tiles);
    meta[1] = total_sparse;
  }
  if (tid < NS) { meta[2 + tid] = dense[tid]; meta[2 + 2 * NS + 1 + tid] = 0; }
  if (tid <= NS) meta[2 + NS + tid] = off[tid];
}

// rowmap[off[k] + pos] = node for every sparse slab k the node needs (rowmap pre-filled with -1).
__global__ void slab_compact_kernel(const unsigned* __restrict__ mask, int T, int KS, int* __restrict__ meta,
                                    int* __restrict__ rowmap, int rowmap_cap) {
  const int NS = KS * KS + 1;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T) return;
  unsigned m = mask[j];
  while (m) {
    const int k = __ffs(m) - 1;
    m &= m - 1;
    if (meta[2 + k]) continue;                        // dense slab: computed for everybody
    const int pos = meta[2 + NS + k] + atomicAdd(&meta[2 + 2 * NS + 1 + k], 1);
    if (pos < rowmap_cap) rowmap[pos] = j;
  }
}

// Copies the fp16 hi / lo halves and the row scale of every mapped node behind the dense rows of the A buffer.
// One warp per compact row; the row count is read from meta[1] (device), the grid is sized for the worst case.
__global__ void __launch_bounds__(256)
slab_gather_rows_kernel(const int* __restrict__ meta, const int* __restrict__ rowmap, uint4* __restrict__ hi,
                        uint4* __restrict__ lo, float* __restrict__ inv, int T_pad, int K) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int total = meta[1];
  const int vec = K / 8;                              // uint4 = 8 halves
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < total; r += nwarps) {
    const int node = rowmap[r];
    if (node < 0) continue;
    const uint4* sh = hi + (size_t)node * vec; const uint4* sl = lo + (size_t)node * vec;
    uint4* dh = hi + (size_t)(T_pad + r) * vec; uint4* dl = lo + (size_t)(T_pad + r) * vec;
    for (int i = lane; i < vec; i += 32) { dh[i] = sh[i]; dl[i] = sl[i]; }
    if (lane == 0) inv[T_pad + r] = inv[node];
  }
}

}  // namespace fpm

extern "C" int fpm_csr_by_dst(const long long* edge_dst, const long long* ptr, const long long* eptr,
                              int* in_ptr, int* in_eid, int B, int total_nodes, int max_edges_per_graph,
                              void* stream) {
  FPM_CHECK_ARG(edge_dst && ptr && eptr && in_ptr && in_eid, "fpm_csr_by_dst: null tensor");
  FPM_CHECK_ARG(B > 0 && max_edges_per_graph >= 0, "fpm_csr_by_dst: bad sizes");
  const size_t smem = (size_t)(2 * max_edges_per_graph + 2) * sizeof(int);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_csr_by_dst: too many edges per graph");
  FPM_CUDA(cudaFuncSetAttribute(fpm::csr_by_dst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  fpm::csr_by_dst_kernel<<<B, fpm::kCsrThreads, smem, (cudaStream_t)stream>>>(
      (const int64_t*)edge_dst, (const int64_t*)ptr, (const int64_t*)eptr, in_ptr, in_eid, total_nodes);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_spline_gather_max(const float* Y, const float* xin, const long long* edge_src,
                                     const float* pseudo, const int* in_ptr, const int* in_eid,
                                     const float* bias, float* out, int* argmax, void* out_hi, void* out_lo,
                                     float* out_inv, int total_nodes, int C, int kernel_size, int mode,
                                     void* stream) {
  FPM_CHECK_ARG(Y && edge_src && pseudo && in_ptr && in_eid && bias && (out || out_hi),
                "fpm_spline_gather_max: null tensor");
  FPM_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr) && (out_hi == nullptr) == (out_inv == nullptr),
                "fpm_spline_gather_max: the fp16 split needs hi, lo and inv");
  FPM_CHECK_ARG(mode == 0 || mode == 2 || (mode == 1 && xin), "fpm_spline_gather_max: bad mode / residual mode needs xin");
  FPM_CHECK_ARG(C % 4 == 0 && C <= 768, "fpm_spline_gather_max: C must be a multiple of 4, at most 768");
  if (total_nodes == 0) return FPM_OK;
  FPM_CHECK_ARG((long long)total_nodes * (kernel_size * kernel_size + 1) * (C / 4) < (1ll << 32),
                "fpm_spline_gather_max: Y exceeds 2^32 float4 (split the batch)");
  // FPMATCH_GATHER_CARVEOUT=1 (experiment, off): same shared-memory carve-out as the slab GEMM.
  static int carve = -1;
  if (carve < 0) {
    const char* e = getenv("FPMATCH_GATHER_CARVEOUT");
    carve = (e && e[0] == '1') ? 1 : 0;        // measured: no help, and the gather alone slows from 0.18 to 0.23 ms
  }
  if (carve) {
    static bool done[16] = {false};
    int dev = 0;
    FPM_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 16 && !done[dev]) {
      FPM_CUDA(cudaFuncSetAttribute(fpm::spline_gather_max_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
      done[dev] = true;
    }
  }
  fpm::spline_gather_max_kernel<<<total_nodes, 192, 0, (cudaStream_t)stream>>>(
      Y, xin, (const int64_t*)edge_src, pseudo, in_ptr, in_eid, bias, out, argmax, (__half*)out_hi, (__half*)out_lo,
      out_inv, C, kernel_size, mode);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_spline_scatter_bwd(const float* G, const int* argmax, const long long* edge_dst,
                                      const float* pseudo, const int* out_ptr, const int* out_eid, float* dY,
                                      int total_nodes, int C, int kernel_size, void* stream) {
  FPM_CHECK_ARG(G && argmax && edge_dst && pseudo && out_ptr && out_eid && dY, "fpm_spline_scatter_bwd: null tensor");
  FPM_CHECK_ARG(C % 4 == 0 && C <= 768, "fpm_spline_scatter_bwd: C must be a multiple of 4, at most 768");
  if (total_nodes == 0) return FPM_OK;
  const size_t smem = (size_t)(kernel_size * kernel_size + 1) * C * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_spline_scatter_bwd: kernel_size too large");
  FPM_CUDA(cudaFuncSetAttribute(fpm::spline_scatter_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fpm::spline_scatter_bwd_kernel<<<total_nodes, 192, smem, (cudaStream_t)stream>>>(
      G, argmax, (const int64_t*)edge_dst, pseudo, out_ptr, out_eid, dY, C, kernel_size);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_spline_scatter_bwd_compact(const float* G, const int* argmax, const long long* edge_dst,
                                              const float* pseudo, const int* out_ptr, const int* out_eid,
                                              const int* colmap, const int* rowpos, float* dYd, float* dYs,
                                              int total_nodes, int C, int kernel_size, int nD, int nS, void* stream) {
  FPM_CHECK_ARG(G && argmax && edge_dst && pseudo && out_ptr && out_eid && colmap && dYd,
                "fpm_spline_scatter_bwd_compact: null tensor");
  FPM_CHECK_ARG(nD >= 1 && nS >= 0 && (nS == 0 || (dYs && rowpos)), "fpm_spline_scatter_bwd_compact: bad groups");
  FPM_CHECK_ARG(C % 4 == 0 && C <= 768, "fpm_spline_scatter_bwd_compact: C must be a multiple of 4, at most 768");
  if (total_nodes == 0) return FPM_OK;
  const size_t smem = (size_t)(kernel_size * kernel_size + 1) * C * sizeof(float);
  FPM_CHECK_ARG(smem <= 200 * 1024, "fpm_spline_scatter_bwd_compact: kernel_size too large");
  FPM_CUDA(cudaFuncSetAttribute(fpm::spline_scatter_bwd_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  fpm::spline_scatter_bwd_compact_kernel<<<total_nodes, 192, smem, (cudaStream_t)stream>>>(
      G, argmax, (const int64_t*)edge_dst, pseudo, out_ptr, out_eid, colmap, rowpos, dYd, dYs, C, kernel_size, nD, nS);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

// Builds the slab plan of one graph batch (see the planner comment above).  mask [T] uint32 and rowmap [rowmap_cap]
// int32 are scratch the caller zero- / -1-fills; meta needs 2 + 3*(KS*KS+1) + 2 + 33 ints, ZERO-FILLED; tab holds
// max_tiles int4.
extern "C" int fpm_spline_plan(const long long* edge_src, const float* pseudo, unsigned* mask, int* meta, int* tab,
                               int* rowmap, int T, int E, int C, int kernel_size, int max_tiles, int rowmap_cap,
                               void* stream) {
  FPM_CHECK_ARG(edge_src && pseudo && mask && meta && tab && rowmap, "fpm_spline_plan: null tensor");
  FPM_CHECK_ARG(T > 0 && E >= 0 && kernel_size * kernel_size + 1 <= 32 && C % 128 == 0 && max_tiles > 0,
                "fpm_spline_plan: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (E > 0) {
    fpm::slab_mask_kernel<<<fpm_cdiv(E, 256), 256, 0, st>>>((const int64_t*)edge_src, pseudo, mask, E, kernel_size);
    FPM_LAUNCH_CHECK();
  }
  const int plan_ctas = fpm_cdiv(T, fpm::kPlanThreads) < 64 ? fpm_cdiv(T, fpm::kPlanThreads) : 64;
  fpm::slab_plan_kernel<<<plan_ctas, fpm::kPlanThreads, 0, st>>>(mask, T, kernel_size, C, meta, (int4*)tab, max_tiles);
  FPM_LAUNCH_CHECK();
  fpm::slab_compact_kernel<<<fpm_cdiv(T, 256), 256, 0, st>>>(mask, T, kernel_size, meta, rowmap, rowmap_cap);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_spline_gather_rows(const int* meta, const int* rowmap, void* a_hi, void* a_lo, float* inv_a,
                                      int T_pad, int K, int rowmap_cap, void* stream) {
  FPM_CHECK_ARG(meta && rowmap && a_hi && a_lo && inv_a, "fpm_spline_gather_rows: null tensor");
  FPM_CHECK_ARG(K % 8 == 0, "fpm_spline_gather_rows: K must be a multiple of 8");
  if (rowmap_cap == 0) return FPM_OK;
  const int blocks = fpm_cdiv(rowmap_cap < 148 * 64 ? rowmap_cap : 148 * 64, 8);
  fpm::slab_gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(meta, rowmap, (uint4*)a_hi, (uint4*)a_lo, inv_a,
                                                                        T_pad, K);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
