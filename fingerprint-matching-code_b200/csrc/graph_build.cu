// Keypoint-graph construction on the device.
//
// The reference builds every graph on the host, one image at a time: scipy/Qhull Delaunay -> dense adjacency
// (/root/reference/utils/build_graphs.py:78-100), python double loop for the incidence factors G, H (:60-72),
// numpy for the PyG edge list and pseudo-coordinates (/root/reference/src/gmdataset.py:169-189), an O(n^3)
// hyper-edge list nobody reads (:180-181) and scipy kron for the index lists (:623-642).  Here the whole batch is
// built by a handful of launches from the padded keypoint tensor:
//
//   fpm_graph_adjacency   A[B,nmax,nmax]       Delaunay through the empty-circle property, or 'fc' / 'near'
//   fpm_graph_row_counts  nonzeros per row      (a torch cumsum turns them into global edge offsets)
//   fpm_graph_edges       edge_index, edge_attr, x, per-pair (src, dst) tables - all in np.nonzero(A) order
//   fpm_graph_permute     A2 = perm^T A1 perm and the mapped edge table (genuine pairs, gmdataset.py:345-352)
//   fpm_graph_incidence   dense one-hot G, H from an edge table
//   fpm_graph_kron_index  the KGHs_sparse lists idx = i2 * n1max + i1 in k2 * e1 + k1 order
//
// Delaunay: for points in general position (i, j) is an edge iff alpha + beta < pi, alpha (beta) the largest
// angle i-k-j over the points k left (right) of i->j; an empty side means a hull edge.  With
// cot(angle) = dot / |cross| this is  dot_L |cross_R| + dot_R |cross_L| > 0.  All arithmetic is fp64 with the
// roundings of the numpy statement in oracle/graphs.py (no fused multiply-add), so decisions agree bit for bit.
// Degenerate inputs follow the rules written in that file's header (segment blocking, fan rule for co-circular
// ties, duplicates isolated, collinear -> fully connected).  Cost: n^3 / 2 triple tests per graph (5e5 at n = 100,
// 3.2e7 at n = 400), a dozen fp64 operations each and no division, spread evenly over n CTAs.
#include "common.cuh"

namespace fpm {

__device__ __forceinline__ double dcross(double ax, double ay, double bx, double by) {
  return __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
}
__device__ __forceinline__ double ddot(double ax, double ay, double bx, double by) {
  return __dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by));
}

// grid (nmax, B), block 128 on a zero-initialised A.  Every unordered pair is decided once: CTA i owns the pairs
// {i, (i + c) mod n}, c = 1 .. n/2 (for even n the column c = n/2 is owned by the lower index only), so all rows
// carry the same load; the result is written to A[i,j] and A[j,i].
__global__ void __launch_bounds__(128)
graph_adjacency_kernel(const double* __restrict__ P, const int64_t* __restrict__ ns, float* __restrict__ A,
                       int nmax, int stg, double thre) {
  extern __shared__ double2 pts[];                // [nmax]
  __shared__ int first_distinct;
  const int b = blockIdx.y, i = blockIdx.x;
  const int n = min((int)ns[b], nmax);
  if (i >= n) return;
  const double2* Pb = reinterpret_cast<const double2*>(P) + (size_t)b * nmax;
  float* Ab = A + (size_t)b * nmax * nmax;
  for (int k = threadIdx.x; k < n; k += blockDim.x) pts[k] = Pb[k];
  if (threadIdx.x == 0) first_distinct = n;
  __syncthreads();

  bool full = (stg != 1) || n < 3;                // 'fc' / 'near', or too few points to triangulate
  bool drop_i = false;
  if (!full) {
    // Qhull fails on a flat input (all points collinear or identical): the reference falls back to 'fc'.
    const double2 p0 = pts[0];
    for (int k = threadIdx.x; k < n; k += blockDim.x)
      if (pts[k].x != p0.x || pts[k].y != p0.y) atomicMin(&first_distinct, k);
    __syncthreads();
    const int q = first_distinct;
    int off_line = 0, dup = 0;
    if (q < n) {
      const double qx = __dsub_rn(pts[q].x, p0.x), qy = __dsub_rn(pts[q].y, p0.y);
      for (int k = threadIdx.x; k < n; k += blockDim.x)
        off_line |= dcross(qx, qy, __dsub_rn(pts[k].x, p0.x), __dsub_rn(pts[k].y, p0.y)) != 0.0;
    }
    for (int k = threadIdx.x; k < i; k += blockDim.x) dup |= (pts[k].x == pts[i].x && pts[k].y == pts[i].y);
    full = !__syncthreads_or(off_line);
    drop_i = __syncthreads_or(dup);               // equal to a lower-indexed point: left isolated
  }

  const int half = n >> 1;
  for (int c = 1 + threadIdx.x; c <= half; c += blockDim.x) {
    if (2 * c == n && i >= half) continue;        // the antipodal pair of an even n belongs to the lower index
    int j = i + c; if (j >= n) j -= n;
    float v = 0.f;
    const int lo = min(i, j), hi = max(i, j);     // canonical orientation lo -> hi
    const double2 pl = pts[lo], ph = pts[hi];
    if (full) {
      v = 1.f;
      if (stg == 2) {
        const double dx = __dsub_rn(ph.x, pl.x), dy = __dsub_rn(ph.y, pl.y);   // hi - lo, the sign numpy's i > j loop uses
        if (sqrt(ddot(dx, dy, dx, dy)) > thre) v = 0.f;
      }
    } else if (!drop_i) {
      double dL = 0, cL = 1, dR = 0, cR = 1;
      int kL = -1, kR = -1;
      bool blocked = false, dup_j = false;
      const double2 pj = pts[j];
      for (int k = 0; k < n; ++k) {
        const double2 pk = pts[k];
        dup_j |= (k < j) && (pk.x == pj.x) && (pk.y == pj.y);
        if (k == lo || k == hi) continue;
        const double ax = __dsub_rn(pl.x, pk.x), ay = __dsub_rn(pl.y, pk.y);
        const double bx = __dsub_rn(ph.x, pk.x), by = __dsub_rn(ph.y, pk.y);
        const double cr = dcross(ax, ay, bx, by), dt = ddot(ax, ay, bx, by);
        // largest angle lo-k-hi = smallest dot / |cross|; dt / cr < dL / cL  <=>  dt * cL < dL * cr
        if (cr > 0.0) {
          if (kL < 0 || __dmul_rn(dt, cL) < __dmul_rn(dL, cr)) { kL = k; dL = dt; cL = cr; }
        } else if (cr < 0.0) {
          if (kR < 0 || __dmul_rn(dt, cR) < __dmul_rn(dR, -cr)) { kR = k; dR = dt; cR = -cr; }
        } else if (dt < 0.0) {
          blocked = true;                         // k lies strictly inside the segment
        }
      }
      bool ok = true;
      if (kL >= 0 && kR >= 0) {
        const double s = __dadd_rn(__dmul_rn(dL, cR), __dmul_rn(dR, cL));
        ok = s > 0.0 || (s == 0.0 && lo < min(kL, kR));
      }
      v = (ok && !blocked && !dup_j) ? 1.f : 0.f;
    }
    if (v != 0.f) {
      Ab[(size_t)i * nmax + j] = v;
      Ab[(size_t)j * nmax + i] = v;
    }
  }
}

// grid (B), block 256: nonzeros of every row (columns >= row only when upper_only), zero for padding rows.
__global__ void __launch_bounds__(256)
graph_row_counts_kernel(const float* __restrict__ A, const int64_t* __restrict__ ns, int* __restrict__ rowcnt,
                        int nmax, int upper_only) {
  const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int n = min((int)ns[b], nmax);
  for (int i = wid; i < nmax; i += nw) {
    int c = 0;
    if (i < n) {
      const float* row = A + ((size_t)b * nmax + i) * nmax;
      for (int j = (upper_only ? i : 0) + lane; j < n; j += 32) c += row[j] != 0.f;
    }
    c = warp_sum_int(c);
    if (lane == 0) rowcnt[(size_t)b * nmax + i] = c;
  }
}

// grid (cdiv(nmax, 4), B), block 128: one warp per adjacency row, ordered compaction by ballot.
__global__ void __launch_bounds__(128)
graph_edges_kernel(const float* __restrict__ A, const double* __restrict__ P, const int64_t* __restrict__ ns,
                   const int64_t* __restrict__ ptr, const int64_t* __restrict__ rowoff, int64_t* __restrict__ edge_index,
                   float* __restrict__ edge_attr, float* __restrict__ x, int* __restrict__ edge_list,
                   int nmax, long long E, int emax, int upper_only, double rescale) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int n = min((int)ns[b], nmax);
  if (i >= n) return;
  const double* Pb = P + (size_t)b * nmax * 2;
  const double xi = Pb[2 * i], yi = Pb[2 * i + 1];
  const int64_t node0 = ptr ? ptr[b] : 0;
  if (x && lane == 0) {
    x[2 * (node0 + i)] = (float)(xi / rescale);
    x[2 * (node0 + i) + 1] = (float)(yi / rescale);
  }
  const float* row = A + ((size_t)b * nmax + i) * nmax;
  int64_t off = rowoff[(size_t)b * nmax + i];
  const int64_t e0 = rowoff[(size_t)b * nmax];
  for (int j0 = upper_only ? (i & ~31) : 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    const bool hit = j < n && (!upper_only || j >= i) && row[j] != 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int64_t e = off + __popc(m & ((1u << lane) - 1u));
      if (edge_index) { edge_index[e] = node0 + i; edge_index[E + e] = node0 + j; }
      if (edge_attr) {
        // 0.5 * (P_i - P_j) / rescale + 0.5, clipped to [0, 1]; fp64 like numpy, rounded once at the end
        const double ux = __dadd_rn(__ddiv_rn(__dmul_rn(0.5, __dsub_rn(xi, Pb[2 * j])), rescale), 0.5);
        const double uy = __dadd_rn(__ddiv_rn(__dmul_rn(0.5, __dsub_rn(yi, Pb[2 * j + 1])), rescale), 0.5);
        edge_attr[2 * e] = (float)fmin(fmax(ux, 0.0), 1.0);
        edge_attr[2 * e + 1] = (float)fmin(fmax(uy, 0.0), 1.0);
      }
      if (edge_list) {
        const int64_t k = e - e0;
        if (k < emax) {
          edge_list[((size_t)b * 2) * emax + k] = i;
          edge_list[((size_t)b * 2 + 1) * emax + k] = j;
        }
      }
    }
    off += __popc(m);
  }
}

// grid (cdiv(max(n1max^2, emax), 256), B): A2[map[i], map[j]] = A1[i, j]; edge table mapped column by column.
__global__ void __launch_bounds__(256)
graph_permute_kernel(const float* __restrict__ A1, const int* __restrict__ map, const int* __restrict__ elist1,
                     float* __restrict__ A2, int* __restrict__ elist2, int n1max, int n2max, int emax) {
  const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  const int* mp = map + (size_t)b * n1max;
  if (t < n1max * n1max) {
    const int i = t / n1max, j = t - i * n1max;
    if (A1[((size_t)b * n1max + i) * n1max + j] != 0.f) {
      const int i2 = mp[i], j2 = mp[j];
      if (i2 >= 0 && j2 >= 0 && i2 < n2max && j2 < n2max) A2[((size_t)b * n2max + i2) * n2max + j2] = 1.f;
    }
  }
  if (elist1 && t < emax) {
    const int s = elist1[((size_t)b * 2) * emax + t], d = elist1[((size_t)b * 2 + 1) * emax + t];
    elist2[((size_t)b * 2) * emax + t] = s >= 0 ? mp[s] : -1;
    elist2[((size_t)b * 2 + 1) * emax + t] = d >= 0 ? mp[d] : -1;
  }
}

// grid (cdiv(emax, 256), B): G[b, src_k, k] = H[b, dst_k, k] = 1 on zero-initialised [B, npad, epad] buffers.
__global__ void __launch_bounds__(256)
graph_incidence_kernel(const int* __restrict__ elist, float* __restrict__ G, float* __restrict__ H, int emax,
                       int npad, int epad) {
  const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= emax || k >= epad) return;
  const int s = elist[((size_t)b * 2) * emax + k], d = elist[((size_t)b * 2 + 1) * emax + k];
  if (s >= 0 && s < npad) G[((size_t)b * npad + s) * epad + k] = 1.f;
  if (d >= 0 && d < npad) H[((size_t)b * npad + d) * epad + k] = 1.f;
}

// grid (cdiv(e1max * e2max, 256), B): column t = k2 * e1 + k1 of kron(G2, G1) has its one at row i2 * n1max + i1.
__global__ void __launch_bounds__(256)
graph_kron_index_kernel(const int* __restrict__ elist1, const int* __restrict__ elist2,
                        const int64_t* __restrict__ es1, const int64_t* __restrict__ es2,
                        const int64_t* __restrict__ koff, int64_t* __restrict__ idxG, int64_t* __restrict__ idxH,
                        int e1max, int e2max, int n1max) {
  const int b = blockIdx.y;
  const int e1 = (int)es1[b], e2 = (int)es2[b];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)e1 * e2) return;
  const int k2 = (int)(t / e1), k1 = (int)(t - (long long)k2 * e1);
  const int s1 = elist1[((size_t)b * 2) * e1max + k1], d1 = elist1[((size_t)b * 2 + 1) * e1max + k1];
  const int s2 = elist2[((size_t)b * 2) * e2max + k2], d2 = elist2[((size_t)b * 2 + 1) * e2max + k2];
  idxG[koff[b] + t] = (int64_t)s2 * n1max + s1;
  idxH[koff[b] + t] = (int64_t)d2 * n1max + d1;
}

}  // namespace fpm

extern "C" int fpm_graph_adjacency(const double* P, const long long* ns, float* A, int B, int nmax, int stg,
                                   double thre, void* stream) {
  FPM_CHECK_ARG(P && ns && A, "fpm_graph_adjacency: null tensor");
  FPM_CHECK_ARG(stg >= 0 && stg <= 2, "fpm_graph_adjacency: strategy must be 0 (fc), 1 (tri) or 2 (near)");
  FPM_CHECK_ARG(B >= 0 && nmax > 0 && B <= 65535, "fpm_graph_adjacency: bad sizes");
  FPM_CHECK_ARG((size_t)nmax * 16 <= 48 * 1024, "fpm_graph_adjacency: more than 3072 keypoints per graph");
  if (B == 0) return FPM_OK;
  FPM_CUDA(cudaMemsetAsync(A, 0, (size_t)B * nmax * nmax * sizeof(float), (cudaStream_t)stream));
  const int threads = nmax / 2 <= 32 ? 32 : nmax / 2 <= 64 ? 64 : 128;     // one thread per owned pair of a row
  fpm::graph_adjacency_kernel<<<dim3(nmax, B), threads, (size_t)nmax * 16, (cudaStream_t)stream>>>(
      P, (const int64_t*)ns, A, nmax, stg, thre);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_graph_row_counts(const float* A, const long long* ns, int* rowcnt, int B, int nmax,
                                    int upper_only, void* stream) {
  FPM_CHECK_ARG(A && ns && rowcnt, "fpm_graph_row_counts: null tensor");
  FPM_CHECK_ARG(B >= 0 && nmax > 0, "fpm_graph_row_counts: bad sizes");
  if (B == 0) return FPM_OK;
  fpm::graph_row_counts_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(A, (const int64_t*)ns, rowcnt, nmax, upper_only);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_graph_edges(const float* A, const double* P, const long long* ns, const long long* ptr,
                               const long long* rowoff, long long* edge_index, float* edge_attr, float* x,
                               int* edge_list, int B, int nmax, long long E, int emax, int upper_only,
                               double rescale, void* stream) {
  FPM_CHECK_ARG(A && P && ns && rowoff, "fpm_graph_edges: null tensor");
  FPM_CHECK_ARG(B >= 0 && nmax > 0 && E >= 0 && emax >= 0 && B <= 65535, "fpm_graph_edges: bad sizes");
  FPM_CHECK_ARG(rescale > 0, "fpm_graph_edges: rescale must be positive");
  FPM_CHECK_ARG(!x || ptr, "fpm_graph_edges: node offsets are needed to write x");
  if (B == 0) return FPM_OK;
  fpm::graph_edges_kernel<<<dim3(fpm_cdiv(nmax, 4), B), 128, 0, (cudaStream_t)stream>>>(
      A, P, (const int64_t*)ns, (const int64_t*)ptr, (const int64_t*)rowoff, (int64_t*)edge_index, edge_attr, x,
      edge_list, nmax, E, emax, upper_only, rescale);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_graph_permute(const float* A1, const int* map, const int* elist1, float* A2, int* elist2, int B,
                                 int n1max, int n2max, int emax, void* stream) {
  FPM_CHECK_ARG(A1 && map && A2, "fpm_graph_permute: null tensor");
  FPM_CHECK_ARG((elist1 == nullptr) == (elist2 == nullptr), "fpm_graph_permute: edge tables must come in pairs");
  FPM_CHECK_ARG(B >= 0 && n1max > 0 && n2max > 0 && emax >= 0 && B <= 65535, "fpm_graph_permute: bad sizes");
  if (B == 0) return FPM_OK;
  const long long work = (long long)n1max * n1max > emax ? (long long)n1max * n1max : emax;
  fpm::graph_permute_kernel<<<dim3(fpm_cdiv(work, 256), B), 256, 0, (cudaStream_t)stream>>>(
      A1, map, elist1, A2, elist2, n1max, n2max, elist1 ? emax : 0);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_graph_incidence(const int* edge_list, float* G, float* H, int B, int emax, int npad, int epad,
                                   void* stream) {
  FPM_CHECK_ARG(edge_list && G && H, "fpm_graph_incidence: null tensor");
  FPM_CHECK_ARG(B >= 0 && emax >= 0 && npad > 0 && epad >= 0 && B <= 65535, "fpm_graph_incidence: bad sizes");
  if (B == 0 || emax == 0 || epad == 0) return FPM_OK;
  fpm::graph_incidence_kernel<<<dim3(fpm_cdiv(emax, 256), B), 256, 0, (cudaStream_t)stream>>>(
      edge_list, G, H, emax, npad, epad);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_graph_kron_index(const int* elist1, const int* elist2, const long long* es1, const long long* es2,
                                    const long long* koff, long long* idxG, long long* idxH, int B, int e1max,
                                    int e2max, int n1max, void* stream) {
  FPM_CHECK_ARG(elist1 && elist2 && es1 && es2 && koff && idxG && idxH, "fpm_graph_kron_index: null tensor");
  FPM_CHECK_ARG(B >= 0 && e1max >= 0 && e2max >= 0 && n1max > 0 && B <= 65535, "fpm_graph_kron_index: bad sizes");
  if (B == 0 || e1max == 0 || e2max == 0) return FPM_OK;
  fpm::graph_kron_index_kernel<<<dim3(fpm_cdiv((long long)e1max * e2max, 256), B), 256, 0, (cudaStream_t)stream>>>(
      elist1, elist2, (const int64_t*)es1, (const int64_t*)es2, (const int64_t*)koff, (int64_t*)idxG,
      (int64_t*)idxH, e1max, e2max, n1max);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
