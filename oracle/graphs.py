"""CPU oracle for the keypoint-graph construction.

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and nothing on the product path).

What the reference does on the host for every image of every pair
(``/root/reference/utils/build_graphs.py:12-74,78-100,103-119`` and
``/root/reference/src/gmdataset.py:169-189``):

1. ``A`` = 0/1 adjacency of the Delaunay triangulation of the keypoints (scipy/Qhull), fully connected when
   ``n < 3`` or Qhull raises, or the fully connected / distance-thresholded graph for ``stg = 'fc' / 'near'``;
2. ``G, H`` = one-hot incidence factors with one column per nonzero of ``A`` in row-major order;
3. the PyG graph: ``edge_index = nonzero(A)``, ``edge_attr = clip(0.5 (P_i - P_j) / 320 + 0.5, 0, 1)``,
   ``x = P / 320``, all computed in fp64 and rounded to fp32 at the end.

Two statements of step 1 live here:

* ``delaunay_adjacency_ref`` calls scipy exactly like the reference (the pin: scipy is installed here);
* ``delaunay_adjacency`` restates the triangulation through the empty-circle property, which is what the
  CUDA kernel evaluates: for points in general position the pair (i, j) is a Delaunay edge iff the largest
  angle ``i-k-j`` over the points k left of ``i->j`` plus the largest over the points right of it is below pi
  (either side may be empty: hull edge).  With ``cot`` of those angles ``= dot / |cross|`` the test is
  ``dot_L * cross_R + dot_R * cross_L > 0``; the largest angle per side is tracked with the division-free
  comparison ``dot * cross_best < dot_best * cross``.  Everything is fp64 with one rounding per operation
  (no fused multiply-add), the same operation order as the kernel.

Degenerate inputs (where Qhull's output depends on its internal merge order and cannot be restated):
  - a point collinear with and strictly between i and j removes the edge (true for any triangulation);
  - exactly co-circular empty quadrilaterals (integer grids): the diagonal that contains the smallest vertex
    index is kept (a fan) - a valid Delaunay triangulation, but Qhull may pick the other diagonal;
  - a point equal to a lower-indexed point is left isolated (matches Qhull's ``coplanar`` handling as probed
    with scipy 1.18.1: the higher index is dropped);
  - all points collinear or identical -> fully connected (the reference's QhullError fallback).
"""
from __future__ import annotations

import numpy as np

RESCALE = 320.0  # max(RESCALE) of /root/reference/src/gmdataset.py:36-48,171


def fully_connect(P: np.ndarray, thre=None) -> np.ndarray:
    """build_graphs.py:103-119."""
    n = P.shape[0]
    A = np.ones((n, n)) - np.eye(n)
    if thre is not None:
        d = np.sqrt(((P[:, None, :].astype(np.float64) - P[None, :, :].astype(np.float64)) ** 2).sum(-1))
        A[d > thre] = 0
    return A


def delaunay_adjacency_ref(P: np.ndarray) -> np.ndarray:
    """The reference's own route (build_graphs.py:78-100): scipy Delaunay, every simplex fully connected."""
    from scipy.spatial import Delaunay
    n = P.shape[0]
    if n < 3:
        return fully_connect(P)
    try:
        d = Delaunay(P)
    except Exception:
        return fully_connect(P)
    A = np.zeros((n, n))
    s = d.simplices
    for a, b in ((0, 1), (0, 2), (1, 2)):
        A[s[:, a], s[:, b]] = 1
        A[s[:, b], s[:, a]] = 1
    return A


def delaunay_adjacency(P: np.ndarray) -> np.ndarray:
    """Empty-circle restatement (the algorithm of csrc/graph_build.cu), fp64."""
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    if n < 3:
        return fully_connect(P)
    # all collinear / identical -> Qhull error -> fully connected
    d0 = P - P[0]
    nz = np.nonzero((d0 != 0).any(1))[0]
    if nz.size == 0:
        return fully_connect(P)
    q = d0[nz[0]]
    if np.all(q[0] * d0[:, 1] - q[1] * d0[:, 0] == 0):
        return fully_connect(P)
    same = (P[:, None, :] == P[None, :, :]).all(-1)
    dropped = np.array([same[i, :i].any() for i in range(n)])
    A = np.zeros((n, n))
    for i in range(n):
        if dropped[i]:
            continue
        js = np.arange(i + 1, n)
        js = js[~dropped[js]]
        J = js.size
        if J == 0:
            continue
        # running best (largest angle = smallest cot = dot / |cross|) on either side of i -> j, scanned in k order
        # with the division-free comparison the kernel uses:  dt / cr < dL / cL  <=>  dt * cL < dL * cr
        kL = np.full(J, -1); kR = np.full(J, -1)
        dL = np.zeros(J); cL = np.ones(J); dR = np.zeros(J); cR = np.ones(J)
        blocked = np.zeros(J, bool)
        for k in range(n):
            ax = P[i, 0] - P[k, 0]; ay = P[i, 1] - P[k, 1]                      # k -> i
            bx = P[js, 0] - P[k, 0]; by = P[js, 1] - P[k, 1]                    # k -> j
            cr = ax * by - ay * bx
            dt = ax * bx + ay * by
            valid = (js != k) & (k != i)
            left = valid & (cr > 0); right = valid & (cr < 0)
            blocked |= valid & (cr == 0) & (dt < 0)
            upL = left & ((kL < 0) | (dt * cL < dL * cr))
            upR = right & ((kR < 0) | (dt * cR < dR * -cr))
            kL = np.where(upL, k, kL); dL = np.where(upL, dt, dL); cL = np.where(upL, cr, cL)
            kR = np.where(upR, k, kR); dR = np.where(upR, dt, dR); cR = np.where(upR, -cr, cR)
        s = dL * cR + dR * cL
        tie_keep = np.minimum(i, js) < np.minimum(kL, kR)
        ok = np.where((kL >= 0) & (kR >= 0), (s > 0) | ((s == 0) & tie_keep), True) & ~blocked
        A[i, js[ok]] = 1
        A[js[ok], i] = 1
    return A


def adjacency(P: np.ndarray, stg: str = "tri", thre=0) -> np.ndarray:
    assert stg in ("fc", "tri", "near"), "No strategy named {} found.".format(stg)
    if stg == "tri":
        return delaunay_adjacency(P)
    if stg == "near":
        return fully_connect(P, thre=thre)
    return fully_connect(P)


def build_graphs(P: np.ndarray, n: int, n_pad=None, edge_pad=None, stg="fc", sym=True, thre=0, ref=False):
    """build_graphs.py:12-74: (A, G, H, edge_num)."""
    A = (delaunay_adjacency_ref(P[:n]) if (ref and stg == "tri") else adjacency(P[:n], stg, thre))
    edge_num = int(A.sum())
    assert n > 0 and edge_num > 0
    n_pad = n if n_pad is None else n_pad
    edge_pad = edge_num if edge_pad is None else edge_pad
    G = np.zeros((n_pad, edge_pad), dtype=np.float32)
    H = np.zeros((n_pad, edge_pad), dtype=np.float32)
    src, dst = np.nonzero(A if sym else np.triu(A))
    G[src, np.arange(src.size)] = 1
    H[dst, np.arange(src.size)] = 1
    return A, G, H, edge_num


def pyg_graph(A: np.ndarray, P: np.ndarray):
    """gmdataset.py:169-189 without the O(n^3) hyperedge list: (x, edge_index, edge_attr)."""
    P = np.asarray(P, dtype=np.float64)
    edge_feat = 0.5 * (P[:, None, :] - P[None, :, :]) / RESCALE + 0.5
    src, dst = np.nonzero(A)
    edge_attr = np.clip(edge_feat[src, dst], 0, 1)
    return (P / RESCALE).astype(np.float32), np.stack([src, dst]).astype(np.int64), edge_attr.astype(np.float32)


def permute_adjacency(A1: np.ndarray, perm: np.ndarray):
    """gmdataset.py:345-352: G2 = perm^T G1, H2 = perm^T H1, A2 = G2 H2^T for a (partial) permutation."""
    _, G1, H1, _ = _gh(A1)
    G2 = perm.T.dot(G1)
    H2 = perm.T.dot(H1)
    return G2.dot(H2.T), G2, H2


def _gh(A):
    n = A.shape[0]
    src, dst = np.nonzero(A)
    G = np.zeros((n, src.size), dtype=np.float32); H = np.zeros((n, src.size), dtype=np.float32)
    G[src, np.arange(src.size)] = 1; H[dst, np.arange(src.size)] = 1
    return A, G, H, src.size
