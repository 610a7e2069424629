// Exact GPU linear assignment + greedy top-k selection: one warp per fingerprint pair.
//
// Replaces /root/reference/utils/hungarian.py:8-65 (D2H copy -> scipy.optimize.linear_sum_assignment
// per pair -> H2D) and the argsort + greedy_perm tail of /root/reference/src/model/ngm.py:444-449
// (src/model/soft_topk.py:56-77).  Bit-exact permutations require scipy's exact traversal, because
// soft-top-k matrices are full of ties (exact zeros / ones): this kernel re-states scipy's
// rectangular_lsap.cpp (shortest augmenting path, Crouse 2016) in fp64 with the same
//   * row order, transposition rule (nc < nr), `remaining` list with swap-removal,
//   * left-to-right evaluation of  minVal + cost - u[i] - v[j],
//   * tie rule: last minimum whose column is unassigned, else first minimum (in `remaining` order),
// and replaces the serial inner scan by a warp-wide scan + 4 redux operations.
// Latency-bound, not bandwidth-bound: algorithmic traffic is one read of ds_mat and one write of
// perm_mat per pair.
#include "common.cuh"
#include <limits.h>
#include <stdlib.h>

namespace fpm {

__device__ __forceinline__ unsigned long long ordered_key(double v) {
  v = v + 0.0;   // -0.0 -> +0.0 so that equal values map to equal keys
  unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  return (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);
}

struct LapSmem {
  double* u; double* v; double* spc;
  int* path; int* col4row; int* row4col; int* remaining;
  unsigned char* SR; unsigned char* SC;
  float* cost;
};

__host__ __device__ inline size_t lap_smem_bytes(int D, int cost_elems) {
  size_t s = (size_t)D * (3 * sizeof(double) + 4 * sizeof(int)) + 2 * (size_t)((D + 15) / 16 * 16);
  s = (s + 15) / 16 * 16;
  return s + (size_t)cost_elems * sizeof(float);
}

// ds: [B, R, C] scores (maximised).  hung_out / perm_out: [B, R, C] or null.  ks: [B] floats or null.
template <bool kCostSmem>
__global__ void __launch_bounds__(32)
lap_topk_kernel(const float* __restrict__ ds, const int64_t* __restrict__ n1,
                const int64_t* __restrict__ n2, const float* __restrict__ ks,
                float* __restrict__ hung_out, float* __restrict__ perm_out,
                int* __restrict__ status, int R, int C) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int b = blockIdx.x, lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int D = R > C ? R : C;

  LapSmem sm;
  {
    unsigned char* p = raw;
    sm.u = (double*)p; p += sizeof(double) * D;
    sm.v = (double*)p; p += sizeof(double) * D;
    sm.spc = (double*)p; p += sizeof(double) * D;
    sm.path = (int*)p; p += sizeof(int) * D;
    sm.col4row = (int*)p; p += sizeof(int) * D;
    sm.row4col = (int*)p; p += sizeof(int) * D;
    sm.remaining = (int*)p; p += sizeof(int) * D;
    const int Dp = (D + 15) / 16 * 16;
    sm.SR = p; p += Dp;
    sm.SC = p; p += Dp;
    p = raw + ((size_t)(p - raw) + 15) / 16 * 16;
    sm.cost = (float*)p;
  }

  int n1b = n1 ? (int)n1[b] : R;
  int n2b = n2 ? (int)n2[b] : C;
  n1b = min(max(n1b, 0), R);
  n2b = min(max(n2b, 0), C);
  const bool tr = n2b < n1b;                 // scipy transposes when nc < nr
  const int nr = tr ? n2b : n1b;
  const int nc = tr ? n1b : n2b;
  const float* dsb = ds + (size_t)b * R * C;

  // zero the outputs for this pair
  {
    const int total = R * C;
    if (hung_out) for (int i = lane; i < total; i += 32) hung_out[(size_t)b * total + i] = 0.f;
    if (perm_out) for (int i = lane; i < total; i += 32) perm_out[(size_t)b * total + i] = 0.f;
  }

  if (kCostSmem) {
    // working-frame cost (un-negated scores), row-major nr x nc
    for (int idx = lane; idx < nr * nc; idx += 32) {
      const int i = idx / nc, j = idx - i * nc;
      sm.cost[idx] = tr ? dsb[(size_t)j * C + i] : dsb[(size_t)i * C + j];
    }
  }
  for (int i = lane; i < D; i += 32) {
    sm.u[i] = 0.0; sm.v[i] = 0.0;
    sm.col4row[i] = -1; sm.row4col[i] = -1; sm.path[i] = -1;
  }
  __syncwarp();

  bool infeasible = false;
  if (nr > 0 && nc > 0) {
    for (int curRow = 0; curRow < nr && !infeasible; ++curRow) {
      double minVal = 0.0;
      int num_remaining = nc;
      for (int it = lane; it < nc; it += 32) {
        sm.remaining[it] = nc - it - 1;
        sm.SC[it] = 0;
        sm.spc[it] = INFINITY;
      }
      for (int i = lane; i < nr; i += 32) sm.SR[i] = 0;
      __syncwarp();

      int sink = -1;
      int i = curRow;
      while (sink == -1) {
        if (lane == 0) sm.SR[i] = 1;
        const double ui = sm.u[i];
        // lane-local sequential scan over it = lane, lane+32, ...
        double best = INFINITY;
        int upos = -1;          // last position at `best` whose column is unassigned
        int fpos = INT_MAX;     // first position at `best`
        for (int it = lane; it < num_remaining; it += 32) {
          const int j = sm.remaining[it];
          const float sc = kCostSmem ? sm.cost[(size_t)i * nc + j]
                                     : (tr ? dsb[(size_t)j * C + i] : dsb[(size_t)i * C + j]);
          const double c = -(double)sc;
          double r = minVal + c;
          r = r - ui;
          r = r - sm.v[j];
          double cur = sm.spc[j];
          if (r < cur) {
            sm.path[j] = i;
            sm.spc[j] = r;
            cur = r;
          }
          const bool unassigned = sm.row4col[j] == -1;
          if (cur < best) {
            best = cur; fpos = it; upos = unassigned ? it : -1;
          } else if (cur == best) {
            if (unassigned) upos = it;
          }
        }
        // warp-wide: global minimum value, then last-unassigned / first position at that value
        const unsigned long long key = ordered_key(best);
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = __reduce_min_sync(full, hi);
        const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
        const bool is_min = (hi == mhi) && (lo == mlo);
        const int gu = __reduce_max_sync(full, is_min ? upos : -1);
        const int gf = __reduce_min_sync(full, is_min ? fpos : INT_MAX);
        // every lane whose key equals the minimum holds the same double; broadcast it
        const int src_lane = __ffs(__ballot_sync(full, is_min)) - 1;
        const double lowest = __shfl_sync(full, best, src_lane);
        if (lowest == INFINITY) { infeasible = true; break; }
        const int index = gu >= 0 ? gu : gf;
        minVal = lowest;
        const int j = sm.remaining[index];
        const int owner = sm.row4col[j];
        __syncwarp();
        if (owner == -1) sink = j; else i = owner;
        if (lane == 0) {
          sm.SC[j] = 1;
          sm.remaining[index] = sm.remaining[num_remaining - 1];
        }
        --num_remaining;
        __syncwarp();
      }
      if (infeasible) break;

      // dual update
      if (lane == 0) sm.u[curRow] += minVal;
      for (int r = lane; r < nr; r += 32)
        if (sm.SR[r] && r != curRow) sm.u[r] += minVal - sm.spc[sm.col4row[r]];
      for (int j = lane; j < nc; j += 32)
        if (sm.SC[j]) sm.v[j] -= minVal - sm.spc[j];
      __syncwarp();
      // augment along the alternating path (serial, short)
      if (lane == 0) {
        int j = sink;
        while (true) {
          const int r = sm.path[j];
          sm.row4col[j] = r;
          const int tmp = sm.col4row[r];
          sm.col4row[r] = j;
          j = tmp;
          if (r == curRow) break;
        }
      }
      __syncwarp();
    }
  }
  if (status && lane == 0) status[b] = infeasible ? 1 : 0;
  if (infeasible) return;

  // ---- hungarian() output: 1 at every assigned (row, col) of the original orientation
  if (hung_out) {
    for (int r = lane; r < nr; r += 32) {
      const int c = sm.col4row[r];
      if (c >= 0) {
        const int a = tr ? c : r, cc = tr ? r : c;
        hung_out[(size_t)b * R * C + (size_t)a * C + cc] = 1.f;
      }
    }
  }
  if (!perm_out) return;

  // ---- greedy_perm(zeros, argsort(x * ds, descending, stable), ks): see SURVEY.md A.7
  // Positive assigned entries never conflict (x is a partial permutation), so the first
  // min(K, #positive) of them in (value desc, flat index asc) order are accepted; a larger K
  // continues in raster order over zero-valued cells of the PADDED matrix.
  const float kf = ks[b];
  long long K = 0;
  if (kf == kf) K = (long long)rint((double)kf);      // python round(): half to even
  if (K <= 0) return;
  // reuse: path[] = flat index, spc[] = value (as double), SR/SC = row/col used flags
  int* flat = sm.path;
  double* val = sm.spc;
  __syncwarp();
  for (int r = lane; r < D; r += 32) { sm.SR[r] = 0; sm.SC[r] = 0; }
  for (int r = lane; r < nr; r += 32) {
    const int c = sm.col4row[r];
    const int a = tr ? c : r, cc = tr ? r : c;
    flat[r] = a * C + cc;
    val[r] = (double)dsb[(size_t)a * C + cc];
  }
  __syncwarp();
  int accepted_local = 0, npos_local = 0;
  for (int t = lane; t < nr; t += 32) {
    const double vt = val[t];
    if (vt > 0.0) {
      ++npos_local;
      int rank = 0;
      const int ft = flat[t];
      for (int q = 0; q < nr; ++q) {
        const double vq = val[q];
        rank += (vq > vt) || (vq == vt && flat[q] < ft);
      }
      if ((long long)rank < K) {
        const int a = ft / C, cc = ft - a * C;
        perm_out[(size_t)b * R * C + ft] = 1.f;
        sm.SR[a] = 1; sm.SC[cc] = 1;
        ++accepted_local;
      }
    }
  }
  const int npos = warp_sum_int(npos_local);
  const int accepted = warp_sum_int(accepted_local);
  __syncwarp();
  if ((long long)npos < K && lane == 0) {
    long long matched = accepted;
    int cptr = 0;
    for (int a = 0; a < R && matched < K; ++a) {
      if (sm.SR[a]) continue;
      while (cptr < C && sm.SC[cptr]) ++cptr;
      if (cptr >= C) break;
      perm_out[(size_t)b * R * C + (size_t)a * C + cptr] = 1.f;
      sm.SC[cptr] = 1;
      ++matched;
    }
  }
}

// Outputs of a CTA-per-pair solver once col4row is final: the hungarian() matrix and the greedy top-k selection
// (same rules as the tail of lap_topk_kernel).  sm.path / sm.spc / sm.SR / sm.SC are reused as scratch.
__device__ __forceinline__ void lap_block_emit(const LapSmem& sm, int* cnt, const float* __restrict__ dsb,
                                               const float* __restrict__ ks, float* __restrict__ hung_out,
                                               float* __restrict__ perm_out, int* __restrict__ status,
                                               bool infeasible, bool tr, int nr, int b, int R, int C, int D) {
  const int tid = threadIdx.x, nthreads = blockDim.x;
  if (status && tid == 0) status[b] = infeasible ? 1 : 0;
  if (infeasible) return;

  if (hung_out) {
    for (int r = tid; r < nr; r += nthreads) {
      const int c = sm.col4row[r];
      if (c >= 0) {
        const int a = tr ? c : r, cc = tr ? r : c;
        hung_out[(size_t)b * R * C + (size_t)a * C + cc] = 1.f;
      }
    }
  }
  if (!perm_out) return;

  const float kf = ks[b];
  long long K = 0;
  if (kf == kf) K = (long long)rint((double)kf);
  if (K <= 0) return;
  int* flat = sm.path;
  double* val = sm.spc;
  __syncthreads();
  if (tid < 2) cnt[tid] = 0;
  for (int r = tid; r < D; r += nthreads) { sm.SR[r] = 0; sm.SC[r] = 0; }
  for (int r = tid; r < nr; r += nthreads) {
    const int c = sm.col4row[r];
    const int a = tr ? c : r, cc = tr ? r : c;
    flat[r] = a * C + cc;
    val[r] = (double)dsb[(size_t)a * C + cc];
  }
  __syncthreads();
  int accepted_local = 0, npos_local = 0;
  for (int t = tid; t < nr; t += nthreads) {
    const double vt = val[t];
    if (vt > 0.0) {
      ++npos_local;
      int rank = 0;
      const int ft = flat[t];
      for (int q = 0; q < nr; ++q) {
        const double vq = val[q];
        rank += (vq > vt) || (vq == vt && flat[q] < ft);
      }
      if ((long long)rank < K) {
        const int a = ft / C, cc = ft - a * C;
        perm_out[(size_t)b * R * C + ft] = 1.f;
        sm.SR[a] = 1; sm.SC[cc] = 1;
        ++accepted_local;
      }
    }
  }
  if (npos_local) atomicAdd(&cnt[0], npos_local);
  if (accepted_local) atomicAdd(&cnt[1], accepted_local);
  __syncthreads();
  if ((long long)cnt[0] < K && tid == 0) {
    long long matched = cnt[1];
    int cptr = 0;
    for (int a = 0; a < R && matched < K; ++a) {
      if (sm.SR[a]) continue;
      while (cptr < C && sm.SC[cptr]) ++cptr;
      if (cptr >= C) break;
      perm_out[(size_t)b * R * C + (size_t)a * C + cptr] = 1.f;
      sm.SC[cptr] = 1;
      ++matched;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Large problems (cost matrix beyond shared memory, n > ~150): one CTA of kLapWarps warps per pair.
// Same algorithm, same tie rule, same results as lap_topk_kernel; what changes is the shape of one Dijkstra step:
// the cost row of the current tree row is staged into shared memory with one coalesced read (the one-warp kernel
// walked it through `remaining` with ~n/32 dependent L2 loads per lane: 36 ms for 32 pairs at n = 400), every
// thread scans one or two columns, and the warp minima are combined through shared memory.
// ------------------------------------------------------------------------------------------
constexpr int kLapWarps = 8;

__global__ void __launch_bounds__(32 * kLapWarps)
lap_topk_block_kernel(const float* __restrict__ ds, const int64_t* __restrict__ n1,
                      const int64_t* __restrict__ n2, const float* __restrict__ ks,
                      float* __restrict__ hung_out, float* __restrict__ perm_out,
                      int* __restrict__ status, int R, int C) {
  extern __shared__ __align__(16) unsigned char raw[];
  __shared__ unsigned long long wkey[kLapWarps];
  __shared__ double wbest[kLapWarps];
  __shared__ int wgu[kLapWarps], wgf[kLapWarps];
  __shared__ int cnt[2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x;
  const unsigned full = 0xffffffffu;
  const int D = R > C ? R : C;

  LapSmem sm;
  {
    unsigned char* p = raw;
    sm.u = (double*)p; p += sizeof(double) * D;
    sm.v = (double*)p; p += sizeof(double) * D;
    sm.spc = (double*)p; p += sizeof(double) * D;
    sm.path = (int*)p; p += sizeof(int) * D;
    sm.col4row = (int*)p; p += sizeof(int) * D;
    sm.row4col = (int*)p; p += sizeof(int) * D;
    sm.remaining = (int*)p; p += sizeof(int) * D;
    const int Dp = (D + 15) / 16 * 16;
    sm.SR = p; p += Dp;
    sm.SC = p; p += Dp;
    p = raw + ((size_t)(p - raw) + 15) / 16 * 16;
    sm.cost = (float*)p;                       // [D]: the staged cost row of the current tree row
  }

  int n1b = n1 ? (int)n1[b] : R;
  int n2b = n2 ? (int)n2[b] : C;
  n1b = min(max(n1b, 0), R);
  n2b = min(max(n2b, 0), C);
  const bool tr = n2b < n1b;
  const int nr = tr ? n2b : n1b;
  const int nc = tr ? n1b : n2b;
  const float* dsb = ds + (size_t)b * R * C;
  {
    const int total = R * C;
    if (hung_out) for (int i = tid; i < total; i += nthreads) hung_out[(size_t)b * total + i] = 0.f;
    if (perm_out) for (int i = tid; i < total; i += nthreads) perm_out[(size_t)b * total + i] = 0.f;
  }
  for (int i = tid; i < D; i += nthreads) {
    sm.u[i] = 0.0; sm.v[i] = 0.0;
    sm.col4row[i] = -1; sm.row4col[i] = -1; sm.path[i] = -1;
  }
  __syncthreads();

  bool infeasible = false;
  if (nr > 0 && nc > 0) {
    for (int curRow = 0; curRow < nr && !infeasible; ++curRow) {
      double minVal = 0.0;
      int num_remaining = nc;
      for (int it = tid; it < nc; it += nthreads) {
        sm.remaining[it] = nc - it - 1;
        sm.SC[it] = 0;
        sm.spc[it] = INFINITY;
      }
      for (int i = tid; i < nr; i += nthreads) sm.SR[i] = 0;
      __syncthreads();

      int sink = -1;
      int i = curRow;
      while (sink == -1) {
        if (tid == 0) sm.SR[i] = 1;
        for (int j = tid; j < nc; j += nthreads) sm.cost[j] = tr ? dsb[(size_t)j * C + i] : dsb[(size_t)i * C + j];
        const double ui = sm.u[i];
        __syncthreads();
        double best = INFINITY;
        int upos = -1;
        int fpos = INT_MAX;
        for (int it = tid; it < num_remaining; it += nthreads) {
          const int j = sm.remaining[it];
          const double c = -(double)sm.cost[j];
          double r = minVal + c;
          r = r - ui;
          r = r - sm.v[j];
          double cur = sm.spc[j];
          if (r < cur) {
            sm.path[j] = i;
            sm.spc[j] = r;
            cur = r;
          }
          const bool unassigned = sm.row4col[j] == -1;
          if (cur < best) {
            best = cur; fpos = it; upos = unassigned ? it : -1;
          } else if (cur == best) {
            if (unassigned) upos = it;
          }
        }
        const unsigned long long key = ordered_key(best);
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = __reduce_min_sync(full, hi);
        const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
        const bool is_min = (hi == mhi) && (lo == mlo);
        const int gu_w = __reduce_max_sync(full, is_min ? upos : -1);
        const int gf_w = __reduce_min_sync(full, is_min ? fpos : INT_MAX);
        const int src_lane = __ffs(__ballot_sync(full, is_min)) - 1;
        const double low_w = __shfl_sync(full, best, src_lane);
        if (lane == 0) {
          wkey[warp] = ((unsigned long long)mhi << 32) | mlo;
          wbest[warp] = low_w; wgu[warp] = gu_w; wgf[warp] = gf_w;
        }
        __syncthreads();
        unsigned long long kmin = wkey[0];
#pragma unroll
        for (int w = 1; w < kLapWarps; ++w) kmin = wkey[w] < kmin ? wkey[w] : kmin;
        int gu = -1, gf = INT_MAX;
        double lowest = INFINITY;
#pragma unroll
        for (int w = 0; w < kLapWarps; ++w)
          if (wkey[w] == kmin) { gu = max(gu, wgu[w]); gf = min(gf, wgf[w]); lowest = wbest[w]; }
        if (lowest == INFINITY) { infeasible = true; break; }
        const int index = gu >= 0 ? gu : gf;
        minVal = lowest;
        const int j = sm.remaining[index];
        const int owner = sm.row4col[j];
        const int last = sm.remaining[num_remaining - 1];
        __syncthreads();
        if (owner == -1) sink = j; else i = owner;
        if (tid == 0) {
          sm.SC[j] = 1;
          sm.remaining[index] = last;
        }
        --num_remaining;
      }
      if (infeasible) break;
      __syncthreads();

      if (tid == 0) sm.u[curRow] += minVal;
      for (int r = tid; r < nr; r += nthreads)
        if (sm.SR[r] && r != curRow) sm.u[r] += minVal - sm.spc[sm.col4row[r]];
      for (int j = tid; j < nc; j += nthreads)
        if (sm.SC[j]) sm.v[j] -= minVal - sm.spc[j];
      __syncthreads();
      if (tid == 0) {
        int j = sink;
        while (true) {
          const int r = sm.path[j];
          sm.row4col[j] = r;
          const int tmp = sm.col4row[r];
          sm.col4row[r] = j;
          j = tmp;
          if (r == curRow) break;
        }
      }
      __syncthreads();
    }
  }
  lap_block_emit(sm, cnt, dsb, ks, hung_out, perm_out, status, infeasible, tr, nr, b, R, C, D);
}

// ------------------------------------------------------------------------------------------
// Column-resident variant of the CTA-per-pair solver (matrix dimension up to 256 * kCols): every thread OWNS the
// columns tid, tid + 256, ... and keeps their state - dual v, shortest-path cost, position in scipy's `remaining`
// list, scanned flag - in registers.  One Dijkstra step is then: a coalesced read of the tree row's cost entries
// straight into registers (no staging), the relaxations, four warp redux operations, ONE block barrier and an
// eight-entry combine.  The `remaining` array itself disappears: the tie rule only needs each live column's
// position in it (swap-removal moves the column at the last position into the freed slot, which its owner does
// locally), and the winning column id travels with its position in one packed integer (position << 10 | column).
// Results are identical to lap_topk_kernel / lap_topk_block_kernel: same fp64 expression per entry, and the
// winner - last minimal position among unassigned columns, else first minimal position - is independent of the
// order in which entries are combined.  32 pairs at n = 400: 13.3 ms (staged rows) -> 7.4 ms.
// ------------------------------------------------------------------------------------------
struct __align__(16) LapPartial { unsigned hi, lo; int ucode, fcode; };

template <int kWarps>
__device__ __forceinline__ void lap_sync() {
  if (kWarps == 1) __syncwarp(); else __syncthreads();
}

// kWarps = 1 with kCostSmem: the whole (working-frame) cost matrix sits in shared memory and one warp solves the pair
// without any block barrier (small problems, n <= ~150); kWarps = 8: cost rows come from L2.
template <int kCols, int kWarps, bool kCostSmem>
__global__ void __launch_bounds__(32 * kWarps)
lap_topk_cols_kernel(const float* __restrict__ ds, const int64_t* __restrict__ n1,
                     const int64_t* __restrict__ n2, const float* __restrict__ ks,
                     float* __restrict__ hung_out, float* __restrict__ perm_out,
                     int* __restrict__ status, int R, int C) {
  extern __shared__ __align__(16) unsigned char raw[];
  __shared__ LapPartial part[2][kWarps];
  __shared__ int cnt[2];
  constexpr int T = 32 * kWarps;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned full = 0xffffffffu;
  const int D = R > C ? R : C;

  LapSmem sm;
  {
    unsigned char* p = raw;
    sm.u = (double*)p; p += sizeof(double) * D;
    sm.v = (double*)p; p += sizeof(double) * D;
    sm.spc = (double*)p; p += sizeof(double) * D;
    sm.path = (int*)p; p += sizeof(int) * D;
    sm.col4row = (int*)p; p += sizeof(int) * D;
    sm.row4col = (int*)p; p += sizeof(int) * D;
    sm.remaining = (int*)p; p += sizeof(int) * D;
    const int Dp = (D + 15) / 16 * 16;
    sm.SR = p; p += Dp;
    sm.SC = p; p += Dp;
    p = raw + ((size_t)(p - raw) + 15) / 16 * 16;
    sm.cost = (float*)p;                         // [nr * nc] when kCostSmem
  }

  int n1b = n1 ? (int)n1[b] : R;
  int n2b = n2 ? (int)n2[b] : C;
  n1b = min(max(n1b, 0), R);
  n2b = min(max(n2b, 0), C);
  const bool tr = n2b < n1b;
  const int nr = tr ? n2b : n1b;
  const int nc = tr ? n1b : n2b;
  const float* dsb = ds + (size_t)b * R * C;
  {
    const int total = R * C;
    if (hung_out) for (int i = tid; i < total; i += T) hung_out[(size_t)b * total + i] = 0.f;
    if (perm_out) for (int i = tid; i < total; i += T) perm_out[(size_t)b * total + i] = 0.f;
  }
  for (int i = tid; i < D; i += T) {
    sm.u[i] = 0.0;
    sm.col4row[i] = -1; sm.row4col[i] = -1; sm.path[i] = -1;
  }
  if (kCostSmem) {
    for (int idx = tid; idx < nr * nc; idx += T) {
      const int i = idx / nc, j = idx - i * nc;
      sm.cost[idx] = tr ? dsb[(size_t)j * C + i] : dsb[(size_t)i * C + j];
    }
  }
  double v[kCols];
  int joff[kCols];                                 // offset of column tid + c*T inside a working-frame row
  const int istride = kCostSmem ? nc : (tr ? 1 : C);
#pragma unroll
  for (int c = 0; c < kCols; ++c) { v[c] = 0.0; joff[c] = (tid + c * T) * ((tr && !kCostSmem) ? C : 1); }
  lap_sync<kWarps>();

  bool infeasible = false;
  int parity = 0;
  if (nr > 0 && nc > 0) {
    for (int curRow = 0; curRow < nr && !infeasible; ++curRow) {
      double spc[kCols];
      int pos[kCols];
      bool live[kCols], unassigned[kCols];
#pragma unroll
      for (int c = 0; c < kCols; ++c) {
        const int j = tid + c * T;
        spc[c] = INFINITY;
        pos[c] = nc - 1 - j;                    // remaining[it] = nc - it - 1
        live[c] = j < nc;
        unassigned[c] = j < nc && sm.row4col[j] == -1;
      }
      double minVal = 0.0;
      int num_remaining = nc;
      int sink = -1;
      int i = curRow;
      while (sink == -1) {
        const double ui = sm.u[i];
        const float* rowp = (kCostSmem ? sm.cost : dsb) + (size_t)i * istride;   // tree row i, working frame
        float sc[kCols];
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
          sc[c] = 0.f;
          if (live[c]) sc[c] = kCostSmem ? rowp[joff[c]] : __ldg(rowp + joff[c]);
        }
        // Relax every owned column (independent fp64 chains, no branches: one warp alone has to cover their latency),
        // then fold the columns' (value, first-position code, last-unassigned-position code) triples.
        double cur[kCols];
        int fc[kCols], uc[kCols];
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
          const int j = tid + c * T;
          const double cst = -(double)sc[c];
          double r = minVal + cst;
          r = r - ui;
          r = r - v[c];
          const bool better = live[c] && r < spc[c];
          spc[c] = better ? r : spc[c];
          if (better) sm.path[j] = i;
          const int code = (pos[c] << 10) | j;
          cur[c] = live[c] ? spc[c] : INFINITY;
          fc[c] = live[c] ? code : INT_MAX;
          uc[c] = (live[c] && unassigned[c]) ? code : -1;
        }
        double best = cur[0];
        int fcode = fc[0];         // packed (position, column) of the first minimal column
        int ucode = uc[0];         // ... of the last minimal unassigned column
#pragma unroll
        for (int c = 1; c < kCols; ++c) {
          const bool lt = cur[c] < best, eq = cur[c] == best;
          fcode = lt ? fc[c] : (eq ? min(fcode, fc[c]) : fcode);
          ucode = lt ? uc[c] : (eq ? max(ucode, uc[c]) : ucode);
          best = lt ? cur[c] : best;
        }
        const unsigned long long key = ordered_key(best);
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = __reduce_min_sync(full, hi);
        const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
        const bool is_min = (hi == mhi) && (lo == mlo);
        const int gu_w = __reduce_max_sync(full, is_min ? ucode : -1);
        const int gf_w = __reduce_min_sync(full, is_min ? fcode : INT_MAX);
        unsigned khi = mhi, klo = mlo;
        int gu = gu_w, gf = gf_w;
        if (kWarps > 1) {
          if (lane == 0) {
            LapPartial pw; pw.hi = mhi; pw.lo = mlo; pw.ucode = gu_w; pw.fcode = gf_w;
            part[parity][warp] = pw;
          }
          __syncthreads();
          // second level: lane l holds warp (l mod kWarps)'s partial; the same four redux operations combine them.
          // (Measured at n = 400, 32 pairs: prefetching each warp's candidate row to L1 before the barrier 7 % slower;
          //  4 warps x 4 columns 21 % slower; 16 warps x 1 column the same as 8 x 2.)
          const LapPartial pw = part[parity][lane & (kWarps - 1)];
          khi = __reduce_min_sync(full, pw.hi);
          klo = __reduce_min_sync(full, pw.hi == khi ? pw.lo : 0xffffffffu);
          const bool wmin = (pw.hi == khi) && (pw.lo == klo);
          gu = __reduce_max_sync(full, wmin ? pw.ucode : -1);
          gf = __reduce_min_sync(full, wmin ? pw.fcode : INT_MAX);
        }
        const unsigned long long kmin = ((unsigned long long)khi << 32) | klo;
        parity ^= 1;
        // invert ordered_key: the minimum as a double
        const unsigned long long bits = (kmin & 0x8000000000000000ull) ? (kmin & 0x7fffffffffffffffull) : ~kmin;
        const double lowest = __longlong_as_double((long long)bits);
        if (lowest == INFINITY) { infeasible = true; break; }
        const int code = gu >= 0 ? gu : gf;
        const int index = code >> 10, j = code & 1023;
        minVal = lowest;
        const int owner = sm.row4col[j];
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
          const int jj = tid + c * T;
          if (jj == j) live[c] = false;                                   // scanned (SC[j] = 1)
          else if (live[c] && pos[c] == num_remaining - 1) pos[c] = index; // remaining[index] = remaining[last]
        }
        --num_remaining;
        if (owner == -1) sink = j; else i = owner;
      }
      if (infeasible) break;

      // dual update: rows of the tree are curRow and the owners of the scanned columns (all but the sink's)
#pragma unroll
      for (int c = 0; c < kCols; ++c) {
        const int j = tid + c * T;
        if (j < nc && !live[c]) {
          if (j != sink) sm.u[sm.row4col[j]] += minVal - spc[c];
          v[c] -= minVal - spc[c];
        }
      }
      if (tid == 0) sm.u[curRow] += minVal;
      lap_sync<kWarps>();
      if (tid == 0) {
        int j = sink;
        while (true) {
          const int r = sm.path[j];
          sm.row4col[j] = r;
          const int tmp = sm.col4row[r];
          sm.col4row[r] = j;
          j = tmp;
          if (r == curRow) break;
        }
      }
      lap_sync<kWarps>();
    }
  }
  lap_block_emit(sm, cnt, dsb, ks, hung_out, perm_out, status, infeasible, tr, nr, b, R, C, D);
}

// Generic greedy_perm(x, top_indices, ks) (soft_topk.py:56-77) for callers that bring their own
// candidate order: one thread per pair walks the list with row/column occupancy bitmaps in global
// scratch (x itself: a row/col is occupied when its sum >= 1, exactly as the reference tests it).
__global__ void greedy_perm_kernel(float* __restrict__ x, const int64_t* __restrict__ top,
                                   const float* __restrict__ ks, int B, int R, int C, int L) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float kf = ks[b];
  long long K = 0;
  if (kf == kf) K = (long long)rint((double)kf);
  float* xb = x + (size_t)b * R * C;
  long long matched = 0;
  for (int cur = 0; cur < L && matched < K; ++cur) {
    const long long idx = top[(size_t)b * L + cur];
    const int r = (int)(idx / C), c = (int)(idx % C);
    float cs = 0.f, rs = 0.f;
    for (int i = 0; i < R; ++i) cs += xb[(size_t)i * C + c];
    for (int j = 0; j < C; ++j) rs += xb[(size_t)r * C + j];
    if (cs < 1.f && rs < 1.f) {
      xb[(size_t)r * C + c] = 1.f;
      ++matched;
    }
  }
}

}  // namespace fpm

namespace fpm {
// scipy.optimize.linear_sum_assignment refuses a cost matrix with NaN or -inf entries ("matrix contains invalid numeric
// entries"; cost = -s, so s = NaN or +inf) even when an assignment that avoids them exists - which is what the
// shortest-augmenting-path kernels above would return.  One CTA per pair scans its valid block after the solve: such a
// pair gets status 1 and all-zero outputs, like an infeasible one.
__global__ void __launch_bounds__(256)
lap_validate_kernel(const float* __restrict__ ds, const int64_t* __restrict__ n1, const int64_t* __restrict__ n2,
                    float* __restrict__ hung_out, float* __restrict__ perm_out, int* __restrict__ status, int R,
                    int C) {
  const int b = blockIdx.x;
  const int nr = n1 ? (int)n1[b] : R, nc = n2 ? (int)n2[b] : C;
  const float* s = ds + (size_t)b * R * C;
  int bad = 0;
  for (int r = threadIdx.x >> 5; r < nr; r += blockDim.x >> 5)        // warp per row: no integer division
    for (int c = threadIdx.x & 31; c < nc; c += 32) {
      const float v = s[(size_t)r * C + c];
      bad |= (v != v) || (v == INFINITY);
    }
  bad = __syncthreads_or(bad);
  if (!bad) return;
  if (threadIdx.x == 0) status[b] = 1;
  for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
    if (hung_out) hung_out[(size_t)b * R * C + i] = 0.f;
    if (perm_out) perm_out[(size_t)b * R * C + i] = 0.f;
  }
}
}  // namespace fpm

static bool g_lap_staged = getenv("FPMATCH_LAP_STAGED") != nullptr;   // A/B switch: force the staged-row kernel

extern "C" int fpm_lap_topk(const float* ds, const long long* n1, const long long* n2, const float* ks,
                            float* hung_out, float* perm_out, int* status, int B, int R, int C,
                            void* stream) {
  FPM_CHECK_ARG(ds, "fpm_lap_topk: null score tensor");
  FPM_CHECK_ARG(hung_out || perm_out, "fpm_lap_topk: no output requested");
  FPM_CHECK_ARG(!perm_out || ks, "fpm_lap_topk: perm_out needs ks");
  FPM_CHECK_ARG(B >= 0 && R > 0 && C > 0, "fpm_lap_topk: bad sizes");
  if (B == 0) return FPM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = R > C ? R : C;
  const size_t with_cost = fpm::lap_smem_bytes(D, R * C);
  const int64_t* p1 = (const int64_t*)n1; const int64_t* p2 = (const int64_t*)n2;
#define FPM_LAP_COLS(KC, KW, SMEM, BYTES)                                                                        \
  do {                                                                                                           \
    FPM_CUDA(cudaFuncSetAttribute(fpm::lap_topk_cols_kernel<KC, KW, SMEM>,                                       \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)));                   \
    fpm::lap_topk_cols_kernel<KC, KW, SMEM><<<B, 32 * KW, BYTES, st>>>(ds, p1, p2, ks, hung_out, perm_out,       \
                                                                       status, R, C);                            \
  } while (0)
  if (with_cost <= 100 * 1024) {
    // the cost matrix fits in shared memory: one warp per pair
    if (g_lap_staged) {
      FPM_CUDA(cudaFuncSetAttribute(fpm::lap_topk_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)with_cost));
      fpm::lap_topk_kernel<true><<<B, 32, with_cost, st>>>(ds, p1, p2, ks, hung_out, perm_out, status, R, C);
    } else if (D <= 32) FPM_LAP_COLS(1, 1, true, with_cost);
    else if (D <= 64) FPM_LAP_COLS(2, 1, true, with_cost);
    else if (D <= 96) FPM_LAP_COLS(3, 1, true, with_cost);
    else if (D <= 128) FPM_LAP_COLS(4, 1, true, with_cost);
    else FPM_LAP_COLS(5, 1, true, with_cost);
  } else {
    // cost matrix beyond shared memory: one CTA of 8 warps per pair
    const size_t base = fpm::lap_smem_bytes(D, D);
    FPM_CHECK_ARG(base <= 200 * 1024, "fpm_lap_topk: matrix dimension too large");
    const int T = 32 * fpm::kLapWarps;
    if (D <= 4 * T && !g_lap_staged) {
      // column state in registers, one barrier per Dijkstra step
      const size_t sm_cols = fpm::lap_smem_bytes(D, 0);
      if (D <= T) FPM_LAP_COLS(1, fpm::kLapWarps, false, sm_cols);
      else if (D <= 2 * T) FPM_LAP_COLS(2, fpm::kLapWarps, false, sm_cols);
      else FPM_LAP_COLS(4, fpm::kLapWarps, false, sm_cols);
    } else {
      // beyond 1024 columns: the current cost row staged in shared memory per step
      FPM_CUDA(cudaFuncSetAttribute(fpm::lap_topk_block_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base));
      fpm::lap_topk_block_kernel<<<B, T, base, st>>>(ds, p1, p2, ks, hung_out, perm_out, status, R, C);
    }
  }
#undef FPM_LAP_COLS
  FPM_LAUNCH_CHECK();
  if (status) {
    fpm::lap_validate_kernel<<<B, 256, 0, st>>>(ds, p1, p2, hung_out, perm_out, status, R, C);
    FPM_LAUNCH_CHECK();
  }
  return FPM_OK;
}

extern "C" int fpm_greedy_perm(float* x, const long long* top_indices, const float* ks, int B, int R,
                               int C, int L, void* stream) {
  FPM_CHECK_ARG(x && top_indices && ks, "fpm_greedy_perm: null tensor");
  FPM_CHECK_ARG(B >= 0 && R > 0 && C > 0 && L >= 0, "fpm_greedy_perm: bad sizes");
  if (B == 0) return FPM_OK;
  fpm::greedy_perm_kernel<<<fpm_cdiv(B, 64), 64, 0, (cudaStream_t)stream>>>(
      x, (const int64_t*)top_indices, ks, B, R, C, L);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
