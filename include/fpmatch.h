/*
 * fpmatch.h - C ABI of libfpmatch_b200.so, the sm_100a matching-head kernels.
 *
 * This is the drop-in boundary for the reference's graph-matching hot path
 * (/root/reference/src/model/ngm.py:205-491 and the ops it calls).  The reference has no C ABI of its
 * own: its natives are pybind11 functions over at::Tensor (src/extension/sparse_dot/sparse_dot.cpp:322-331,
 * src/extension/bilinear_diag/bilinear_diag.cpp:324-326) and its hot ops are python functions.  Every entry
 * point below therefore names the PYTHON interface it replaces (file:line); INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add at each site.
 *
 * Conventions
 *   - plain pointers and sizes only; every tensor pointer is a DEVICE pointer to a contiguous fp32 /
 *     int64 / int32 array in the row-major layout given in the comment; `stream` is a cudaStream_t
 *     passed as void* (0 = legacy default stream).  Nothing synchronises the host.
 *   - keypoint counts n1/n2/ns and graph offsets ptr/eptr are int64 (`long long`), exactly the dtype the
 *     reference's data_dict carries (src/gmdataset.py:563-672), so no conversion launch is needed.
 *   - return value: 0 on success, a negative FPM_ERR_* for rejected arguments, a positive cudaError_t
 *     for launch failures; fpm_last_error() returns the message of the last failure on this thread.
 *     This mirrors the reference's behaviour of raising on bad shapes/devices
 *     (sparse_dot.cpp:42-45 CHECK_INPUT, utils/hungarian.py:29 ValueError).
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef FPMATCH_H_
#define FPMATCH_H_

#ifdef __cplusplus
extern "C" {
#endif

#define FPM_OK 0
#define FPM_ERR_ARG (-1)
#define FPM_ERR_UNSUPPORTED (-2)

/* ---- library ------------------------------------------------------------------------------------ */
int fpm_abi_version(void);
int fpm_device_ok(void);                 /* 1 iff the current device is compute capability 10.x */
const char* fpm_last_error(void);
void fpm_set_error(const char* msg);

/* ---- (1) feature_align ----------------------------------------------------------------------------
 * fpm_feature_align replaces utils.feature_align.feature_align (utils/feature_align.py:5-37):
 *   fmap [B,C,Hf,Wf], P [B,nmax,2] (x,y), ns [B] -> out [B,C,nmax], zero beyond ns[b]; bit-exact,
 *   including the (W,H)/(Hf,Wf) scaling mix-up and the post-fetch edge rule (:55-62, :98-118).
 *   feat_coords = 1: P already holds feature-space (x,y) = bilinear_interpolate(im, x, y) (:67-125).
 * fpm_fmap_prep + fpm_node_features are the fused form used by Net.forward (ngm.py:241-251):
 *   raw NCHW map -> channels-last map divided by its channel L2 norm; then one warp per keypoint writes
 *   X[ptr[b]+i, 0:C1+C2] = [align(nodes), align(edges)].
 * fpm_global_max replaces final_layers = AdaptiveMaxPool2d(1,1) (feature_extractor.py:54, ngm.py:238):
 *   out[b*out_stride + out_offset + c] = max_hw fmap[b,c,:].
 * fpm_affinity_coeff: coeff[b,:] = tanh(A * (g[b]/||g[b]||) + a)  (ngm.py:262-268, affinity_layer.py:13).
 */
int fpm_feature_align(const float* fmap, const float* P, const long long* ns, float* out, int B, int C,
                      int Hf, int Wf, int nmax, float ori_w, float ori_h, int feat_coords, void* stream);
int fpm_fmap_prep(const float* fmap_nchw, float* out_nhwc, int B, int C, int Hf, int Wf, void* stream);
int fpm_global_max(const float* fmap, float* out, int B, int C, int HW, int out_stride, int out_offset,
                   void* stream);
int fpm_node_features(const float* nodes_nhwc, const float* edges_nhwc, const float* P, const long long* ns,
                      const long long* ptr, float* X, int B, int nmax, int C1, int H1, int W1, int C2, int H2,
                      int W2, float ori_w, float ori_h, void* stream);
int fpm_affinity_coeff(const float* gcat, const float* W, const float* bias, float* coeff, int B, int IN,
                       int OUT, void* stream);

/* ---- dense contractions ------------------------------------------------------------------------------
 * C[M,N] = act(A[M,K] * Bt[N,K]^T + bias[N]), act 0 = none, 1 = relu.  fp32 CUDA cores (fpm_gemm_nt_f32)
 * or tcgen05 tensor cores with fp32-faithful 3xTF32 splitting (fpm_gemm_nt_tc, passes = 3) / plain TF32
 * (passes = 1).  Used for SplineConv's slab GEMM, the AFA-U projections and feed-forward.
 */
int fpm_gemm_nt_f32(const float* A, const float* Bt, const float* bias, float* C, int M, int N, int K,
                    int lda, int ldb, int ldc, int act, void* stream);
int fpm_tf32_split(const float* src, float* hi, float* lo, long long n, void* stream); /* src = hi + lo, tf32-exact */
int fpm_gemm_nt_tc(const float* A_hi, const float* A_lo, const float* Bt_hi, const float* Bt_lo,
                   const float* bias, float* C, int M, int N, int K, int lda, int ldb, int ldc, int act,
                   int passes, void* stream);   /* passes = 1: *_hi are the raw operands, *_lo ignored */
/* Error-compensated fp16 mode (same accuracy as 3xTF32, twice the MMA rate): fpm_f16_split_rows scales each row
 * by a power of two s and writes a*s = hi + 2^-11 lo (fp16 hi/lo, [rows,K]) and inv_scale[row] = 1/s. */
int fpm_f16_split_rows(const float* src, void* hi_f16, void* lo_f16, float* inv_scale, int rows, int K,
                       void* stream);
int fpm_gemm_nt_f16x3(const void* A_hi, const void* A_lo, const float* inv_a, const void* Bt_hi,
                      const void* Bt_lo, const float* inv_b, const float* bias, float* C, int M, int N, int K,
                      int lda, int ldb, int ldc, int act, void* stream);

/* Kernel choice for the error-compensated modes: 1 (default) = persistent CTA-pair kernel (tcgen05.mma.cta_group::2 on
 * 256x128 tiles, TMEM double-buffered so the epilogue of one tile overlaps the main loop of the next), 0 = one 128x256
 * tile per CTA.  Also settable with the environment variable FPMATCH_GEMM_PAIR=0/1 before the first call. */
int fpm_gemm_set_pair(int on);
int fpm_gemm_set_max_clusters(int clusters);   /* cap on the SM pairs the persistent GEMM occupies (0 = all) */

/* Tile-table form of fpm_gemm_nt_f16x3 (persistent CTA-pair kernel): computes only the *tab_count C blocks listed in
 * tab (device, 4 ints per tile: first A row of a 256-row block, first Bt row of a 128-row block, first C column, row-map
 * offset or -1 for C row = A row); rowmap[off + r] = C row of A row (a0 + r), -1 = padding.  M = rows of the A buffer,
 * m_ident = rows of C that identity tiles may write, max_tiles = host-side bound of *tab_count (sizes the grid). */
int fpm_gemm_nt_f16x3_tiles(const void* A_hi, const void* A_lo, const float* inv_a, const void* Bt_hi,
                            const void* Bt_lo, const float* inv_b, float* C, int M, int N, int K, int lda, int ldb,
                            int ldc, const int* tab, const int* tab_count, const int* rowmap, long long max_tiles,
                            int m_ident, void* stream);

/* debug aid: record {smid, t_entry, t_setup, t_mainloop_done, t_end} (ns) per CTA into buf[5*cap] (NULL = off) */
int fpm_gemm_set_trace(void* buf, int cap);

/* ---- (3a) SplineConv -----------------------------------------------------------------------------------
 * Replaces torch_geometric SplineConv(768,768,dim=2,kernel_size=5,aggr='max') as driven by
 * src/model/spline_conv.py:28-58.  Y [total_nodes, KS*KS+1, C] = x @ [W_0 .. W_24, root] comes from the GEMM
 * above; fpm_spline_gather_max blends the 4 active slabs per in-edge, takes the max over in-edges, adds the
 * root slab + bias, then relu (mode 0) or x + 0.1*out (mode 1).  fpm_csr_by_dst builds the in-edge lists
 * from edge_index[1] (global node ids of a PyG-style batch; ptr/eptr = node/edge offsets per graph).
 */
/* Slab planner (device side, no host round trip): marks the (node, weight slab) products the gather will read and emits
 * the tile table for fpm_gemm_nt_f16x3_tiles - slabs needed by >= 1/4 of the nodes are computed for all nodes, the
 * others only for their nodes, compacted behind the dense rows of the A buffer (fpm_spline_gather_rows copies the
 * fp16 halves there).  mask [T] u32 zero-filled, rowmap [rowmap_cap] i32 filled with -1, meta [2 + 3*(KS*KS+1) + 2 + 33] i32
 * zero-filled
 * (meta[0] = tile count, meta[1] = compact rows in use), tab [max_tiles] int4. */
int fpm_spline_plan(const long long* edge_src, const float* pseudo, unsigned* mask, int* meta, int* tab, int* rowmap,
                    int T, int E, int C, int kernel_size, int max_tiles, int rowmap_cap, void* stream);
int fpm_spline_gather_rows(const int* meta, const int* rowmap, void* a_hi, void* a_lo, float* inv_a, int T_pad, int K,
                           int rowmap_cap, void* stream);
int fpm_csr_by_dst(const long long* edge_dst, const long long* ptr, const long long* eptr, int* in_ptr,
                   int* in_eid, int B, int total_nodes, int max_edges_per_graph, void* stream);
int fpm_spline_gather_max(const float* Y, const float* xin, const long long* edge_src, const float* pseudo,
                          const int* in_ptr, const int* in_eid, const float* bias, float* out, int* argmax,
                          void* out_hi, void* out_lo, float* out_inv, int total_nodes, int C, int kernel_size,
                          int mode, void* stream);
                          /* argmax (optional, training): [total_nodes,C] int32 winning edge id per channel.
                           * out_hi / out_lo / out_inv (optional, all or none): the result rows also (or, with out =
                           * NULL, only) as the fp16 hi / lo halves + row scales of fpm_f16_split_rows, i.e. directly
                           * as the A operand of the next layer's fpm_gemm_nt_f16x3_tiles */

/* ---- (2) affinities --------------------------------------------------------------------------------------
 * Replaces InnerProductWithWeightsAffinity.forward (src/model/affinity_layer.py:11-22) for Kp (node mode:
 * eidx* = NULL, rows ptrA[b]..ptrA[b+1]) and Ke (edge mode: row r = X[eidx[0][eptr[b]+r]] - X[eidx[1][..]],
 * src/model/spline_conv.py:73-81), fused with the 0.5 factor of ngm.py:287 (`scale`) and the zero padding of
 * ngm.py:317-318.  out [B,Rmax,Cmax]; out_t (optional) the transpose [B,Cmax,Rmax] (= emb of ngm.py:321).
 */
int fpm_affinity(const float* XA, const float* XB, const float* coeff, const long long* ptrA,
                 const long long* ptrB, const long long* eptrA, const long long* eptrB, const long long* eidxA,
                 const long long* eidxB, int EA, int EB, float* out, float* out_t, int B, int Rmax, int Cmax,
                 int Kdim, float scale, int raw, void* stream);   /* raw = 1: plain dot products, no softplus */
/* Ke through linearity: with P = raw node products (X1 (.) c_edge) X2^T [B,Rn,Cn] (fpm_affinity, raw = 1),
 * out[b,k1,k2] = scale * (softplus(P[s1,s2] - P[s1,d2] - P[d1,s2] + P[d1,d2]) - 0.5); 33x fewer FLOPs than
 * the reference's e1 x 768 x e2 product (ngm.py:282-287). */
/* Tensor-core route of the same affinity (csrc/affinity_tc.cu): fpm_f16_split_rows_scaled = error-compensated fp16 split
 * of X rows times their pair's coefficient vector (ptr [B+1] = row offsets of the pairs); fpm_affinity_tiles = tile
 * table + row map for fpm_gemm_nt_f16x3_tiles (tab [B*tA*tB,4], tA = ceil(Rmax/256), tB = ceil(Cmax/128); rowmap
 * [B*tA*256]); the GEMM writes raw products into P [B*Rmax, ldp = 128*tB]; fpm_affinity_finish applies
 * scale * (softplus - 0.5) (raw = 0) and the zero padding -> out [B,Rmax,Cmax] (+ out_t [B,Cmax,Rmax] optional). */
int fpm_f16_split_rows_scaled(const float* src, const float* coeff, const long long* ptr, int B, void* hi, void* lo,
                              float* inv_scale, int rows, int K, void* stream);
int fpm_affinity_tiles(const long long* ptrA, const long long* ptrB, int B, int Rmax, int Cmax, int* tab,
                       int* tab_count, int* rowmap, void* stream);
int fpm_affinity_finish(const float* P, const long long* ptrA, const long long* ptrB, float* out, float* out_t, int B,
                        int Rmax, int Cmax, int ldp, float scale, int raw, void* stream);
int fpm_affinity_edges_factored(const float* P, const long long* eidxA, const long long* eptrA,
                                const long long* ptrA, const long long* eidxB, const long long* eptrB,
                                const long long* ptrB, int EA, int EB, float* out, int B, int Rn, int Cn,
                                int e1max, int e2max, float scale, void* stream);

/* ---- (2)/(3b) association-graph GNN -------------------------------------------------------------------------
 * Replaces construct_sparse_aff_mat + SparseTensor + PYGNNLayer.forward (utils/factorize_graph_matching.py:57-95,
 * src/model/ngm.py:326-348, src/model/gnn.py:207-218) with the Kronecker structure kept factorised.
 * edges: [B,2,emax] int32, per pair the (G-node, H-node) of every G/H column, -1 padded.
 * weights (9 device pointers): lin_l.weight, lin_l.bias, lin_r.weight, n_self_func.0.weight, .0.bias,
 * n_self_func.2.weight, .2.bias, classifier.weight, classifier.bias.
 * xprev [B,N,16] (cin = 17) or NULL (cin = 1); mprev_t [B,n2max,n1max]; xout [B,N,16]; score [B,n1max,n2max].
 * fpm_final_classifier: s[b,i1,i2] = classifier([x1, sinkhorn channel])  (ngm.py:368-369).
 *
 * fpm_assoc_effective: the structure the reference's index lists really describe (gmdataset.py:623-642 eliminates the
 * zero columns of kron(G2,G1) and kron(H2,H1) SEPARATELY, ngm.py:333-342 cuts [idx; diag] to the length of K_value =
 * e1_pyg*e2_pyg + n1*n2).  edges1/edges2 as above; eptr1/eptr2 [B+1] int64 edge offsets of the two PyG batches;
 * outputs: eff1/eff2 [B,2,emax] effective edge tables, ndiag [B] int64 surviving self-loop count (n1*n2 for complete
 * tables), part [B,4] int32 = (ps2, pd2, ccut, 0) the column block the cut ends in (ccut = 0: none), status [1] int32
 * (bit 0: G / H lists of different length, i.e. an asymmetric adjacency; accumulated with atomicOr, caller zeroes).
 * fpm_assoc_in_csr: in-neighbour lists of an edge table; in_col (optional) = column id of every list entry.
 * fpm_gnn_layer: in_*1 / in_*2 from fpm_assoc_in_csr on eff1 / eff2; part may be NULL (no cut-off block anywhere).
 */
int fpm_assoc_effective(const int* edges1, const int* edges2, const long long* eptr1, const long long* eptr2,
                        const long long* n1, const long long* n2, int* eff1, int* eff2, long long* ndiag, int* part,
                        int* status, int B, int e1max, int e2max, void* stream);
int fpm_assoc_in_csr(const int* edges, int* in_ptr, int* in_src, int* in_col, int B, int nmax, int emax,
                     void* stream);
int fpm_gnn_layer(const float* xprev, const float* mprev_t, const int* in_ptr1, const int* in_src1,
                  const int* in_col1, const int* in_ptr2, const int* in_src2, const long long* ndiag,
                  const int* part, const float* const* weights, float* xout, float* score, int B, int n1max,
                  int n2max, int e1max, int e2max, int cin, void* stream);
int fpm_final_classifier(const float* x1, const float* sk_t, const float* cw, const float* cb, float* s, int B,
                         int n1max, int n2max, void* stream);

/* ---- (4) Sinkhorn / soft-top-k ------------------------------------------------------------------------------
 * fpm_sinkhorn_log replaces Sinkhorn.forward -> pygmtools.sinkhorn(backend='pytorch', batched_operation=False)
 * (src/model/sinkhorn.py:58-87): s [B,R,C], n1/n2 [B] (NULL = full) -> out [B,R,C] (+ out_t [B,C,R] optional).
 * fpm_soft_topk replaces soft_topk(..., return_prob=True)[1] (src/model/soft_topk.py:8-53,166-255).
 * workspace: device scratch of fpm_*_workspace_bytes() bytes, needed only when a pair's matrix does not fit
 * in shared memory (0 bytes means none).
 */
long long fpm_sinkhorn_workspace_bytes(int B, int R, int C, int dummy_row);
int fpm_sinkhorn_log(const float* s, const long long* n1, const long long* n2, float* out, float* out_t,
                     void* workspace, int B, int R, int C, int max_iter, float tau, int dummy_row, void* stream);
long long fpm_soft_topk_workspace_bytes(int B, int R, int C);
int fpm_soft_topk(const float* scores, const float* ks, const long long* n1, const long long* n2, float* out,
                  void* workspace, int B, int R, int C, int max_iter, float tau, void* stream);

/* ---- (5) AFA-U ---------------------------------------------------------------------------------------------
 * fpm_afau_attention replaces CrossSet_MultiHeadAttention.forward (src/model/afau.py:231-300): q [B,nr,256],
 * k,v [B,nc,256], cost addressed cost[b*cs_b + i*cs_r + j*cs_c] -> out [B,nr,256].
 * fpm_add_instnorm replaces AddAndInstanceNormalization.forward (afau.py:154-176) (+ optional max over rows,
 * ngm.py:402-405; `out` may be NULL when only that maximum is wanted, for n <= 112 and E % 4 == 0).  fpm_onehot_proj: projection of the one-hot column embedding of ngm.py:396-399.
 * fpm_k_head replaces final_row/final_col + sigmoid (ngm.py:406-412); weights (8 pointers): final_row.0.weight,
 * .0.bias, .2.weight, .2.bias, final_col.0.weight, .0.bias, .2.weight, .2.bias.
 */
int fpm_afau_attention(const float* q, const float* k, const float* v, const float* cost, long long cs_b,
                       long long cs_r, long long cs_c, const float* mix1_w, const float* mix1_b,
                       const float* mix2_w, const float* mix2_b, float* out, int B, int nr, int nc, int q_zero,
                       void* stream);   /* q_zero = 1: caller guarantees q == 0 (row block of Net.forward): q.k is skipped */
int fpm_add_instnorm(const float* a, const float* other, int other_mode, const float* gamma, const float* beta,
                     float* out, float* rowmax, int B, int n, int E, float eps, void* stream);
int fpm_onehot_proj(const float* W, const long long* n, float* out, int B, int nmax, int OUT, int IN,
                    void* stream);
/* add_instnorm(onehot, vec) where onehot[b, r, c] = (c == r && r < hot[b]) is the column embedding of ngm.py:396-399,
 * never materialised (n <= 112, E % 4 == 0). */
int fpm_onehot_instnorm(const long long* hot, const float* vec, const float* gamma, const float* beta, float* out,
                        float* rowmax, int B, int n, int E, float eps, void* stream);
int fpm_k_head(const float* g_row, const float* g_col, const float* const* weights, const long long* n1,
               const long long* n2, float* ks, float* k_scaled, int B, int E, int Hd, int mean_k, void* stream);

/* ---- (6) linear assignment + greedy top-k ------------------------------------------------------------------
 * fpm_lap_topk replaces utils.hungarian.hungarian (utils/hungarian.py:8-65; scipy linear_sum_assignment,
 * reproduced exactly including tie-breaks) -> hung_out [B,R,C] (optional), and the argsort + greedy_perm tail
 * of ngm.py:445-449 -> perm_out [B,R,C] (optional; needs ks [B] = k * min(n1,n2), rounded half-to-even).
 * status [B] (optional): 1 where the cost matrix was infeasible (scipy would raise).
 * fpm_greedy_perm replaces src.model.soft_topk.greedy_perm (soft_topk.py:56-77) for caller-supplied orders.
 */
int fpm_lap_topk(const float* ds, const long long* n1, const long long* n2, const float* ks, float* hung_out,
                 float* perm_out, int* status, int B, int R, int C, void* stream);
int fpm_greedy_perm(float* x, const long long* top_indices, const float* ks, int B, int R, int C, int L,
                    void* stream);

/* ---- (N2) loss and metrics --------------------------------------------------------------------------------------
 * fpm_permutation_loss replaces the per-pair loop of PermutationLoss.forward (src/loss_func.py:26-59):
 *   pair_sum[b] = sum over the valid n1_b x n2_b block of BCE(pred, gt) (logs clamped at -100 as torch does);
 *   the loss is sum_b pair_sum[b] / sum_b n1_b.  fpm_permutation_loss_bwd: grad = gscale[0] * (p - y) / max(p(1-p), 1e-12)
 *   inside the valid blocks, 0 outside (gscale = upstream gradient / sum n1, a device scalar).
 * fpm_matching_stats replaces the loops of matching_recall / matching_precision (src/evaluation_metric.py:58-131):
 *   stats[b] = {sum(pred*gt), sum(gt), sum(pred)} over rows < ns[b].
 */
int fpm_permutation_loss(const float* pred, const float* gt, const long long* n1, const long long* n2,
                         float* pair_sum, int B, int R, int C, void* stream);
int fpm_permutation_loss_bwd(const float* pred, const float* gt, const long long* n1, const long long* n2,
                             const float* gscale, float* grad, int B, int R, int C, void* stream);
int fpm_matching_stats(const float* pred, const float* gt, const long long* ns, float* stats, int B, int R, int C,
                       void* stream);
/* The scalar tail of Net.forward in eval mode (ngm.py:456-469) in one launch: cls_prob [B] = sigmoid(logits);
 * scalars[0] = BCE-with-logits mean against label [B] (0 if label is NULL), scalars[1] = k_factor * mse(ks, gt_ks / min(n1, n2)),
 * scalars[2] = l1(ks * min(n1, n2), gt_ks) with gt_ks = sum(gt_perm [B,R,C]) (both 0 if ks is NULL).
 * workspace: 3*B floats + 1 int, the int zero before the first call (the kernel rearms it); one call at a time per workspace. */
int fpm_head_losses(const float* logits, const float* label, const float* ks, const float* gt_perm, const long long* n1,
                    const long long* n2, float k_factor, float* cls_prob, float* workspace, float* scalars, int B,
                    int R, int C, void* stream);

/* ---- (A14) batched CSR / CSC products, dense factorised-graph-matching affinity ------------------------------------
 * Replace the reference's JIT extension (src/extension/sparse_dot/sparse_dot.cpp:191-331, bilinear_diag.cpp:303-326)
 * behind src.sparse_torch.CSRMatrix3d.dot / dotdiag, src.sparse.bilinear_diag_torch and RebuildFGM
 * (utils/factorize_graph_matching.py:140-186).  Container layout as src/sparse_torch/csx_matrix.py:20-93: indices
 * int64 [nnz] (local, ascending per row / column), indptr int64 [B*h+1] (CSR) / [B*w+1] (CSC) with global offsets.
 *   fpm_csr_dot_diag        out_data[p] = data[p] * diag[b, indices[p]]                 (csr_dot_diag_to_csr)
 *   fpm_csr_dot_csc_dense   out [B,h,w] = CSR [B,h,k] . CSC [B,k,w]                      (csr_dot_csc_to_dense)
 *   fpm_dense_dot_csc_dense out [B,h,w] = dense [B,h,k] . CSC [B,k,w]                    (dense_dot_csc_to_dense)
 *   fpm_bilinear_diag       out [B,x] = diag(CSR [B,x,f] . T [B,f,f] . CSC [B,f,x])      (bilinear_diag)
 *   fpm_fgm_rebuild         K [B,N,N] = sum_t G[:,t] ke_vec[b,t] H[:,t]^T + diag(kp_vec), from GT = CSR [B,E,N] and
 *                           HT = CSC [B,N,E] (the transposed Kronecker factors): one scatter over the E columns.
 * The reference's sparse x sparse -> sparse product exists on the CPU only (sparse_dot.cpp:50-144, raises for CUDA
 * at :204) and is not rebuilt.
 */
int fpm_csr_dot_diag(const long long* indices, const long long* indptr, const float* data, const float* diag,
                     float* out_data, int B, int h, int w, void* stream);
int fpm_csr_dot_csc_dense(const long long* ind1, const long long* ptr1, const float* dat1, const long long* ind2,
                          const long long* ptr2, const float* dat2, float* out, int B, int h, int w, void* stream);
int fpm_dense_dot_csc_dense(const float* dense, const long long* ind2, const long long* ptr2, const float* dat2,
                            float* out, int B, int h, int k, int w, void* stream);
int fpm_bilinear_diag(const long long* ind1, const long long* ptr1, const float* dat1, const float* T,
                      const long long* ind3, const long long* ptr3, const float* dat3, float* out, int B, int x,
                      int f, void* stream);
int fpm_fgm_rebuild(const long long* indg, const long long* ptrg, const float* datg, const long long* indh,
                    const long long* ptrh, const float* dath, const float* ke_vec, const float* kp_vec, float* K,
                    int B, int E, int N, void* stream);

/* ---- training: hand-written backward of the differentiable ops -----------------------------------------------
 * The reference differentiates its forward with torch autograd (train.py / src/train/training_loop.py:33-64 call
 * loss.backward() on PermutationLoss(ds_mat)); these entry points are the vector-Jacobian products of the kernels
 * above, wired into torch.autograd.Function objects by fpmatch/autograd.py.  Row of SURVEY.md section 8(a) in brackets.
 *
 * [A1] fpm_node_features_bwd: dX [total,C1+C2] -> gradients of the two PREPARED (channels-last, normalised) maps;
 *      fpm_fmap_prep_bwd: through the channel L2 normalisation back to the raw NCHW map (ngm.py:65-67,241-251).
 * [A2] fpm_spline_scatter_bwd: G [total,C] (gradient of the conv output before relu / residual scaling) + the
 *      argmax edge ids -> dY [total, KS*KS+1, C]; out_ptr/out_eid = edge lists grouped by SOURCE node
 *      (fpm_csr_by_dst on edge_index[0]).  dX = dY W and dW = dY^T X are the dense GEMMs above; fpm_transpose_f32
 *      makes their K-major operands ([R,C] -> [C,ldo], zero padded).
 * [A4] fpm_bmm_ragged: Out[ptrO[b]+i,:] = coeff_out[b,:] (.) sum_j M[b,i,j] (X[ptrX[b]+j,:] (.) coeff_in[b,:]) (trans = 1:
 *      M[b,j,i]); fpm_segment_rowdot: out[b,:] = sum_{rows of pair b} X (.) Y.  Together: dX1, dX2, d coeff of
 *      InnerProductWithWeightsAffinity (affinity_layer.py:11-19).
 * [A7] fpm_gnn_layer_bwd: backward of fpm_gnn_layer (forward recomputed from its inputs); grads = flat fp32 buffer
 *      lin_l.weight[16*cin] lin_l.bias[16] lin_r.weight[16*cin] n_self_func.0.weight[16*cin] .0.bias[16]
 *      n_self_func.2.weight[256] .2.bias[16] classifier.weight[16] classifier.bias[1], accumulated (caller zeroes);
 *      out_ptr/out_dst/out_col: out-neighbour lists (fpm_assoc_in_csr on the EFFECTIVE edge table with its two rows
 *      swapped); ndiag / part as in fpm_gnn_layer;
 *      gagg: scratch [B,N,4] (cin 1) / [B,N,20] (cin 17); dxprev [B,N,16], dm [B,n1max,n2max] are overwritten.
 * [A8] fpm_sinkhorn_log_bwd: gout [B,R,C] -> gs [B,R,C] through the unrolled iterations (forward replayed on chip).
 * [A10] fpm_soft_topk_bwd: gout [B,R,C] -> gscores [B,R,C]; anchors and k carry no gradient (soft_topk.py:27).
 */
int fpm_node_features_bwd(const float* dX, const float* P, const long long* ns, const long long* ptr,
                          float* dnodes_nhwc, float* dedges_nhwc, int B, int nmax, int C1, int H1, int W1, int C2,
                          int H2, int W2, float ori_w, float ori_h, void* stream);
int fpm_fmap_prep_bwd(const float* fmap_nchw, const float* dy_nhwc, float* dx_nchw, int B, int C, int Hf, int Wf,
                      void* stream);
int fpm_spline_scatter_bwd(const float* G, const int* argmax, const long long* edge_dst, const float* pseudo,
                           const int* out_ptr, const int* out_eid, float* dY, int total_nodes, int C,
                           int kernel_size, void* stream);
/* Column-compacted scatter: colmap [KS*KS+1] int32 (>= 0: column block of dYd [total,nD,C]; -(g+1): column block g of
 * dYs [nR,nS,C], written at row rowpos[node] when that is >= 0; INT_MIN: slab unread, skipped). */
int fpm_spline_scatter_bwd_compact(const float* G, const int* argmax, const long long* edge_dst, const float* pseudo,
                                   const int* out_ptr, const int* out_eid, const int* colmap, const int* rowpos,
                                   float* dYd, float* dYs, int total_nodes, int C, int kernel_size, int nD, int nS,
                                   void* stream);
int fpm_transpose_f32(const float* src, float* dst, int R, int C, int ldo, void* stream);
int fpm_bmm_ragged(const float* Mat, int B, int Rmax, int Cmax, int trans, const float* X, const long long* ptrX,
                   const long long* ptrO, const float* coeff_in, const float* coeff_out, float* Out, int D,
                   void* stream);
int fpm_segment_rowdot(const float* X, const float* Y, const long long* ptr, float* out, int B, int D, void* stream);
int fpm_gnn_layer_bwd(const float* xprev, const float* mprev_t, const int* in_ptr1, const int* in_src1,
                      const int* in_col1, const int* in_ptr2, const int* in_src2, const int* out_ptr1,
                      const int* out_dst1, const int* out_col1, const int* out_ptr2, const int* out_dst2,
                      const long long* ndiag, const int* part, const float* const* weights, const float* dxout,
                      const float* dscore, float* dxprev, float* dm, float* gagg, float* grads, int B, int n1max,
                      int n2max, int e1max, int e2max, int cin, void* stream);
/* [A9] fpm_afau_attention_bwd: backward of fpm_afau_attention (out = its forward output); dq is overwritten, dk / dv and
 *      dmix [16 heads x 65: mix1_weight[h,0,:], mix1_weight[h,1,:], mix1_bias[h,:], mix2_weight[h,:], mix2_bias[h]] are
 *      accumulated (caller zero-fills).  fpm_add_instnorm_bwd: dy [B,n,E] and / or drowmax [B,E] -> dx [B,n,E] (gradient of
 *      `a` and of a tensor `other`); dgamma, dbeta, dvec (row-vector `other`, mode 2) [E] are accumulated. */
int fpm_afau_attention_bwd(const float* q, const float* k, const float* v, const float* cost, long long cs_b,
                           long long cs_r, long long cs_c, const float* mix1_w, const float* mix1_b,
                           const float* mix2_w, const float* mix2_b, const float* out, const float* dout, float* dq,
                           float* dk, float* dv, float* dmix, int B, int nr, int nc, void* stream);
int fpm_add_instnorm_bwd(const float* a, const float* other, int other_mode, const float* gamma, const float* dy,
                         const float* drowmax, float* dx, float* dgamma, float* dbeta, float* dvec, int B, int n, int E,
                         float eps, void* stream);
long long fpm_sinkhorn_bwd_workspace_bytes(int B, int R, int C, int max_iter);
int fpm_sinkhorn_log_bwd(const float* s, const long long* n1, const long long* n2, const float* gout, float* gs,
                         void* workspace, int B, int R, int C, int max_iter, float tau, int dummy_row, void* stream);
long long fpm_soft_topk_bwd_workspace_bytes(int B, int R, int C);
int fpm_soft_topk_bwd(const float* scores, const float* ks, const long long* n1, const long long* n2,
                      const float* gout, float* gscores, void* workspace, int B, int R, int C, int max_iter,
                      float tau, void* stream);

/* ---- [A13] genuine / imposter classifier, inference (eval-mode BatchNorm) ------------------------------------------
 * Replaces MatchClassifier.forward (/root/reference/src/model/ngm.py:75-106) on `s * perm_mat` (ngm.py:451-455):
 * two blocks of Conv3x3(pad 1) -> ReLU -> BatchNorm2d(running statistics) -> MaxPool2 with 1 -> 16 -> 32 channels,
 * global average pool, Linear(32 -> 1).  s, perm [B,H,W] (perm nullable: classify s itself); w1 [16,1,3,3], b1 [16],
 * w2 [32,16,3,3], b2 [32]; bn1 / bn2 = arrays of 4 device pointers {weight, bias, running_mean, running_var};
 * fcw [32], fcb [1]; workspace = fpm_match_classifier_workspace_floats(B,H,W) floats; logits [B].  H, W >= 4.
 */
long long fpm_match_classifier_workspace_floats(int B, int H, int W);
int fpm_match_classifier(const float* s, const float* perm, const float* w1, const float* b1,
                         const float* const* bn1, const float* w2, const float* b2, const float* const* bn2,
                         const float* fcw, const float* fcb, float eps, float* workspace, float* logits, int B, int H,
                         int W, void* stream);

/* ---- dense NGM-v1 message passing --------------------------------------------------
 * Replaces the aggregation of GNNLayer.forward (/root/reference/src/model/gnn.py:54-68): A [B,N,N] 0/1 adjacency,
 * W [B,N,N,fe] edge tensor (fe = 1 or F), X [B,N,F] (F in {1,2,4,8,16,32}).
 *   trans = 0: out[b,i,:] = sum_j A[i,j] s_i W[i,j,:] X[j,:],  s_i = 1 / max(sum_j |A[i,j]|, 1e-12) when norm (written to
 *              inv [B,N] when non-null), else 1;
 *   trans = 1: out[b,j,:] = sum_i A[i,j] inv[i] W[i,j,:] X[i,:]  (the backward's dx1; inv nullable = 1).
 * fpm_fgm_aggregate_dw: dW[b,i,j,c'] = A[i,j] inv[i] (fe == 1 ? sum_c dx2[i,c] x1[j,c] : dx2[i,c'] x1[j,c']).
 */
int fpm_fgm_aggregate(const float* A, const float* W, const float* X, float* inv, float* out, int B, int N, int F,
                      int fe, int norm, int trans, void* stream);
int fpm_fgm_aggregate_dw(const float* A, const float* inv, const float* dx2, const float* x1, float* dW, int B, int N,
                         int F, int fe, void* stream);

/* ---- keypoint-graph construction ---------------------------------------------------
 * Replaces the per-image host code of /root/reference/utils/build_graphs.py:12-119 (build_graphs,
 * delaunay_triangulate, fully_connect) and /root/reference/src/gmdataset.py:169-189 (to_pyg_graph), :345-352
 * (graph 2 of a genuine pair = perm^T graph 1) and :623-642 (Kronecker index lists) for a whole padded batch.
 *   P [B,nmax,2] fp64 keypoints (x, y), ns [B] valid counts.
 *   fpm_graph_adjacency: A [B,nmax,nmax] 0/1 fp32, zero padded; stg 0 = 'fc', 1 = 'tri' (Delaunay), 2 = 'near' (thre).
 *   fpm_graph_row_counts: rowcnt [B*nmax] = nonzeros of each row (columns >= row when upper_only: sym = False).
 *   fpm_graph_edges: rowoff [B*nmax+1] = exclusive prefix sum of rowcnt; E = total; writes (each nullable)
 *      edge_index [2,E] int64 (node ids offset by ptr[b]), edge_attr [E,2] = clip(0.5 (P_i-P_j)/rescale + 0.5, 0, 1),
 *      x [sum(ns),2] = P/rescale, edge_list [B,2,emax] int32 pair-local (src, dst) (caller pre-fills -1).
 *   fpm_graph_permute: map [B,n1max] int32 (node of graph 2 matched to node i of graph 1, -1 = none):
 *      A2[map[i],map[j]] = A1[i,j] on a zeroed A2 [B,n2max,n2max]; elist2 = elist1 mapped column by column.
 *   fpm_graph_incidence: edge_list -> dense one-hot G, H [B,npad,epad] (caller zero-fills).
 *   fpm_graph_kron_index: idxG/idxH flat int64, pair b at koff[b] (koff = prefix sum of es1*es2), entry
 *      k2*e1+k1 = i2*n1max + i1 (source nodes for G, target nodes for H).
 */
int fpm_graph_adjacency(const double* P, const long long* ns, float* A, int B, int nmax, int stg, double thre,
                        void* stream);
int fpm_graph_row_counts(const float* A, const long long* ns, int* rowcnt, int B, int nmax, int upper_only,
                         void* stream);
int fpm_graph_edges(const float* A, const double* P, const long long* ns, const long long* ptr,
                    const long long* rowoff, long long* edge_index, float* edge_attr, float* x, int* edge_list,
                    int B, int nmax, long long E, int emax, int upper_only, double rescale, void* stream);
int fpm_graph_permute(const float* A1, const int* map, const int* elist1, float* A2, int* elist2, int B, int n1max,
                      int n2max, int emax, void* stream);
int fpm_graph_incidence(const int* edge_list, float* G, float* H, int B, int emax, int npad, int epad, void* stream);
int fpm_graph_kron_index(const int* elist1, const int* elist2, const long long* es1, const long long* es2,
                         const long long* koff, long long* idxG, long long* idxH, int B, int e1max, int e2max,
                         int n1max, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FPMATCH_H_ */
