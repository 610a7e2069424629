// Batched CSR / CSC kernels and the factorised-graph-matching affinity rebuild.
//
// Replaces the reference's JIT extension (/root/reference/src/extension/sparse_dot/*.cu|cpp,
// /root/reference/src/extension/bilinear_diag/*.cu|cpp: thread-per-output scalar loops on the legacy default
// stream) behind src.sparse_torch.{CSRMatrix3d,CSCMatrix3d}.dot/dotdiag, src.sparse.bilinear_diag_torch and
// utils.factorize_graph_matching.RebuildFGM (the dense NGM-v1 path that ngm.py:294-315 keeps commented out).
//
// Container layout (same as the reference, src/sparse_torch/csx_matrix.py:20-93): a batch of B sparse matrices of
// one shape [h, w]; `indptr` has B*h + 1 (CSR) or B*w + 1 (CSC) int64 entries holding GLOBAL offsets into
// `indices` / `data`; `indices` are local column (CSR) or row (CSC) ids, ascending inside a row / column.
//
// fpm_fgm_rebuild is not a port of the reference's route (CSR.diag -> CSR.CSC merge-join per output element,
// N^2 threads each walking two index lists): with the transposed factors at hand,
//     K = sum_t  G[:, t] v[t] H[:, t]^T + diag(kp)
// is a scatter over the E = e1*e2 Kronecker columns t (one thread per t; every column of G2 (x) G1 / H2 (x) H1
// holds a single one for incidence factors, so that is one atomic per t), and its gradient
//     dv[t] = sum_{r in G[:, t]} sum_{c in H[:, t]} g h dK[r, c]
// is the matching gather (fpm_bilinear_diag).
#include "common.cuh"

namespace fpm {

// out_data[p] = data[p] * diag[b, indices[p]]   (CSR . diag(v); sparse_dot/csr_dot_diag_cuda.cu:9-34)
__global__ void csr_dot_diag_kernel(const int64_t* __restrict__ indices, const int64_t* __restrict__ indptr,
                                    const float* __restrict__ data, const float* __restrict__ diag,
                                    float* __restrict__ out, long long rows_total, int h, int w) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows_total) return;
  const long long b = row / h;
  const float* d = diag + b * w;
  for (int64_t p = indptr[row]; p < indptr[row + 1]; ++p) out[p] = data[p] * d[indices[p]];
}

// out[b, i, j] = sum_k A[b, i, k] * Bm[b, k, j], A in CSR [h, k], Bm in CSC [k, w]: sorted-list intersection.
__global__ void csr_dot_csc_dense_kernel(const int64_t* __restrict__ ind1, const int64_t* __restrict__ ptr1,
                                         const float* __restrict__ dat1, const int64_t* __restrict__ ind2,
                                         const int64_t* __restrict__ ptr2, const float* __restrict__ dat2,
                                         float* __restrict__ out, int h, int w, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long b = idx / ((long long)h * w);
  const long long rem = idx - b * (long long)h * w;
  const int i = (int)(rem / w), j = (int)(rem - (long long)i * w);
  int64_t p = ptr1[b * h + i], pe = ptr1[b * h + i + 1];
  int64_t q = ptr2[b * w + j], qe = ptr2[b * w + j + 1];
  float acc = 0.f;
  while (p < pe && q < qe) {
    const int64_t a = ind1[p], c = ind2[q];
    if (a == c) { acc = fmaf(dat1[p], dat2[q], acc); ++p; ++q; }
    else if (a < c) ++p;
    else ++q;
  }
  out[idx] = acc;
}

// out[b, i, j] = sum_q D[b, i, ind2[q]] * dat2[q] over column j of the CSC matrix [k, w].
__global__ void dense_dot_csc_dense_kernel(const float* __restrict__ D, const int64_t* __restrict__ ind2,
                                           const int64_t* __restrict__ ptr2, const float* __restrict__ dat2,
                                           float* __restrict__ out, int h, int k, int w, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long b = idx / ((long long)h * w);
  const long long rem = idx - b * (long long)h * w;
  const int i = (int)(rem / w), j = (int)(rem - (long long)i * w);
  const float* drow = D + (b * h + i) * (long long)k;
  float acc = 0.f;
  for (int64_t q = ptr2[b * w + j]; q < ptr2[b * w + j + 1]; ++q) acc = fmaf(drow[ind2[q]], dat2[q], acc);
  out[idx] = acc;
}

// out[b, i] = sum_{p in row i of S1} sum_{q in col i of S3} S1[p] * T[b, idx1[p], idx3[q]] * S3[q]
// S1: CSR [x, f], T: dense [f, f], S3: CSC [f, x]   (= diag(S1 T S3); bilinear_diag_cuda.cu:7-44)
__global__ void bilinear_diag_kernel(const int64_t* __restrict__ ind1, const int64_t* __restrict__ ptr1,
                                     const float* __restrict__ dat1, const float* __restrict__ T,
                                     const int64_t* __restrict__ ind3, const int64_t* __restrict__ ptr3,
                                     const float* __restrict__ dat3, float* __restrict__ out, int x, int f,
                                     long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long b = idx / x;
  const float* Tb = T + b * (long long)f * f;
  float acc = 0.f;
  for (int64_t p = ptr1[idx]; p < ptr1[idx + 1]; ++p) {
    const float a = dat1[p];
    const float* trow = Tb + ind1[p] * (long long)f;
    for (int64_t q = ptr3[idx]; q < ptr3[idx + 1]; ++q) acc = fmaf(a * trow[ind3[q]], dat3[q], acc);
  }
  out[idx] = acc;
}

// K[b, r, c] += G[r, t] v[b, t] H[c, t] for every Kronecker column t; GT: CSR [E, N] (row t lists r),
// HT: CSC [N, E] (column t lists c).  K must be zero-filled.
__global__ void fgm_scatter_kernel(const int64_t* __restrict__ indg, const int64_t* __restrict__ ptrg,
                                   const float* __restrict__ datg, const int64_t* __restrict__ indh,
                                   const int64_t* __restrict__ ptrh, const float* __restrict__ dath,
                                   const float* __restrict__ v, float* __restrict__ K, int E, int N,
                                   long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // b * E + t
  if (idx >= total) return;
  const long long b = idx / E;
  const float vt = v[idx];
  if (vt == 0.f) return;
  float* Kb = K + b * (long long)N * N;
  for (int64_t p = ptrg[idx]; p < ptrg[idx + 1]; ++p) {
    const float gv = datg[p] * vt;
    float* krow = Kb + indg[p] * (long long)N;
    for (int64_t q = ptrh[idx]; q < ptrh[idx + 1]; ++q) atomicAdd(krow + indh[q], gv * dath[q]);
  }
}

__global__ void add_diag_kernel(float* __restrict__ K, const float* __restrict__ d, int N, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // b * N + i
  if (idx >= total) return;
  const long long b = idx / N, i = idx - b * N;
  K[(b * N + i) * (long long)N + i] += d[idx];
}

}  // namespace fpm

static inline unsigned blocks_for(long long total) { return (unsigned)((total + 255) / 256); }

extern "C" int fpm_csr_dot_diag(const long long* indices, const long long* indptr, const float* data,
                                const float* diag, float* out_data, int B, int h, int w, void* stream) {
  FPM_CHECK_ARG(indices && indptr && data && diag && out_data, "fpm_csr_dot_diag: null tensor");
  const long long rows = (long long)B * h;
  if (rows == 0) return FPM_OK;
  fpm::csr_dot_diag_kernel<<<blocks_for(rows), 256, 0, (cudaStream_t)stream>>>(
      (const int64_t*)indices, (const int64_t*)indptr, data, diag, out_data, rows, h, w);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_csr_dot_csc_dense(const long long* ind1, const long long* ptr1, const float* dat1,
                                     const long long* ind2, const long long* ptr2, const float* dat2, float* out,
                                     int B, int h, int w, void* stream) {
  FPM_CHECK_ARG(ind1 && ptr1 && dat1 && ind2 && ptr2 && dat2 && out, "fpm_csr_dot_csc_dense: null tensor");
  const long long total = (long long)B * h * w;
  if (total == 0) return FPM_OK;
  FPM_CHECK_ARG(total / 256 < 0x7fffffffLL, "fpm_csr_dot_csc_dense: output too large");
  fpm::csr_dot_csc_dense_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const int64_t*)ind1, (const int64_t*)ptr1, dat1, (const int64_t*)ind2, (const int64_t*)ptr2, dat2, out, h, w,
      total);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_dense_dot_csc_dense(const float* dense, const long long* ind2, const long long* ptr2,
                                       const float* dat2, float* out, int B, int h, int k, int w, void* stream) {
  FPM_CHECK_ARG(dense && ind2 && ptr2 && dat2 && out, "fpm_dense_dot_csc_dense: null tensor");
  const long long total = (long long)B * h * w;
  if (total == 0) return FPM_OK;
  FPM_CHECK_ARG(total / 256 < 0x7fffffffLL, "fpm_dense_dot_csc_dense: output too large");
  fpm::dense_dot_csc_dense_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      dense, (const int64_t*)ind2, (const int64_t*)ptr2, dat2, out, h, k, w, total);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_bilinear_diag(const long long* ind1, const long long* ptr1, const float* dat1, const float* T,
                                 const long long* ind3, const long long* ptr3, const float* dat3, float* out, int B,
                                 int x, int f, void* stream) {
  FPM_CHECK_ARG(ind1 && ptr1 && dat1 && T && ind3 && ptr3 && dat3 && out, "fpm_bilinear_diag: null tensor");
  const long long total = (long long)B * x;
  if (total == 0) return FPM_OK;
  fpm::bilinear_diag_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const int64_t*)ind1, (const int64_t*)ptr1, dat1, T, (const int64_t*)ind3, (const int64_t*)ptr3, dat3, out, x,
      f, total);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}

extern "C" int fpm_fgm_rebuild(const long long* indg, const long long* ptrg, const float* datg,
                               const long long* indh, const long long* ptrh, const float* dath, const float* ke_vec,
                               const float* kp_vec, float* K, int B, int E, int N, void* stream) {
  FPM_CHECK_ARG(indg && ptrg && datg && indh && ptrh && dath && ke_vec && kp_vec && K, "fpm_fgm_rebuild: null tensor");
  FPM_CHECK_ARG(B >= 0 && E >= 0 && N > 0, "fpm_fgm_rebuild: bad sizes");
  if (B == 0) return FPM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FPM_CUDA(cudaMemsetAsync(K, 0, (size_t)B * N * N * sizeof(float), st));
  const long long total = (long long)B * E;
  if (total > 0) {
    fpm::fgm_scatter_kernel<<<blocks_for(total), 256, 0, st>>>((const int64_t*)indg, (const int64_t*)ptrg, datg,
                                                               (const int64_t*)indh, (const int64_t*)ptrh, dath,
                                                               ke_vec, K, E, N, total);
    FPM_LAUNCH_CHECK();
  }
  fpm::add_diag_kernel<<<blocks_for((long long)B * N), 256, 0, st>>>(K, kp_vec, N, (long long)B * N);
  FPM_LAUNCH_CHECK();
  return FPM_OK;
}
