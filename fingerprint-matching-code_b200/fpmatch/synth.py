"""Seeded synthetic fingerprint-pair batches in the ``data_dict`` format ``Net.forward`` consumes.

No dataset ships with the reference (``dataset/Synthetic/R1`` is absent), so every test and
benchmark uses this generator.  It reproduces what the reference's CPU data pipeline emits:

* keypoints in the 320x240 frame, Delaunay adjacency and the incidence factors ``G``/``H`` in the
  row-major edge order of ``/root/reference/utils/build_graphs.py:60-72``;
* the PyG-style graph with pseudo-coordinates ``clip(0.5*(P_i-P_j)/320+0.5, 0, 1)`` in
  ``np.nonzero(A)`` order (``/root/reference/src/gmdataset.py:169-189``);
* genuine pairs: jittered copy, ``gt_perm = I``, graph 2 = ``perm^T G1`` (``gmdataset.py:345-352``);
  imposter pairs: an independent point set with its own triangulation and ``gt_perm = 0``;
* optionally the Kronecker index lists ``KGHs_sparse`` (``gmdataset.py:623-642``), which only the CPU
  oracle needs - the CUDA path works from the per-graph edge lists.

The backbone is out of scope, so batches carry seeded feature maps (``fmaps``) shaped like the
ResNet-18 layer3/layer4 outputs for a 240x320 image instead of images.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .graph import GraphBatch, GraphData

RESCALE = (320, 240)  # (W, H), /root/reference/src/gmdataset.py:36-48


def delaunay_adjacency(P: np.ndarray) -> np.ndarray:
    """Symmetric 0/1 adjacency of the Delaunay triangulation (fully connected for n < 3)."""
    n = P.shape[0]
    A = np.zeros((n, n), dtype=np.float64)
    if n < 3:
        return np.ones((n, n)) - np.eye(n)
    from scipy.spatial import Delaunay
    try:
        tri = Delaunay(P)
    except Exception:  # QhullError -> same fallback as the reference (build_graphs.py:96-99)
        return np.ones((n, n)) - np.eye(n)
    s = tri.simplices
    for a, b in ((0, 1), (0, 2), (1, 2)):
        A[s[:, a], s[:, b]] = 1
        A[s[:, b], s[:, a]] = 1
    return A


def incidence_from_adjacency(A: np.ndarray):
    """Edge list (src, dst) in row-major ``A`` order plus the dense one-hot factors G, H."""
    src, dst = np.nonzero(A)
    n, e = A.shape[0], src.shape[0]
    G = np.zeros((n, e), dtype=np.float32)
    H = np.zeros((n, e), dtype=np.float32)
    G[src, np.arange(e)] = 1
    H[dst, np.arange(e)] = 1
    return src.astype(np.int64), dst.astype(np.int64), G, H


def pyg_like_graph(src: np.ndarray, dst: np.ndarray, P: np.ndarray) -> GraphData:
    """A is given through its row-major nonzeros (src, dst); mirrors ``to_pyg_graph``."""
    rescale = max(RESCALE)
    edge_attr = 0.5 * (P[src] - P[dst]) / rescale + 0.5
    edge_attr = np.clip(edge_attr, 0, 1)
    return GraphData(
        x=torch.tensor(P / rescale).to(torch.float32),
        edge_index=torch.tensor(np.stack([src, dst]), dtype=torch.long),
        edge_attr=torch.tensor(edge_attr).to(torch.float32),
    )


def _pad_stack(arrs, dtype=torch.float32):
    shape = [max(a.shape[i] for a in arrs) for i in range(arrs[0].ndim)]
    out = torch.zeros([len(arrs)] + shape, dtype=dtype)
    for b, a in enumerate(arrs):
        sl = (b,) + tuple(slice(0, s) for s in a.shape)
        out[sl] = torch.as_tensor(a, dtype=dtype)
    return out


def make_batch(batch_size: int, n: int, seed: int = 1234, imposter_every: int = 2,
               ragged: bool = False, n_min: Optional[int] = None, with_kron: bool = False,
               with_dense_gh: bool = True, with_fmaps: bool = True, jitter: float = 1.5,
               fmap_seed: Optional[int] = None, fmap_noise: Optional[float] = None, partial: int = 0) -> dict:
    """Build one batch.

    ``partial = d`` turns every genuine pair into a PARTIAL match, the normal case of the reference's real data
    (``get_pair``: keypoints of one print without a counterpart in the other): ``d`` keypoints of image 1 have no
    match, image 2 carries the matched ones in shuffled order plus ``d`` keypoints of its own, ``gt_perm`` is a partial
    permutation and graph 2 = ``perm^T G1`` / ``perm^T H1`` keeps only what the permutation carries over
    (gmdataset.py:345-352) - so G2 and H2 have DIFFERENT all-zero columns and the Kronecker index lists of
    gmdataset.py:623-642 lose their pairing (ngm.py:333-342 truncates them to a common length).

    ``imposter_every = k`` makes every k-th pair (b % k == k-1) an imposter; 0 = all genuine.
    ``ragged`` draws n1_b, n2_b uniformly from [n_min, n] (imposters get independent sizes).
    ``fmap_noise = s`` makes the second image's feature maps a noisy copy of the first's (maps2 = maps1 + s*N(0,1)),
    so that matching keypoints of a genuine pair have related features and training has something to learn;
    the default draws the two images' maps independently.
    """
    rng = np.random.RandomState(seed)
    n_min = n_min if n_min is not None else max(4, n // 2)
    P1s, P2s, g1s, g2s, G1s, H1s, G2s, H2s, perms, labels = [], [], [], [], [], [], [], [], [], []
    e1s, e2s, edges1, edges2, kgh = [], [], [], [], []
    for b in range(batch_size):
        genuine = not (imposter_every and b % imposter_every == imposter_every - 1)
        n1 = int(rng.randint(n_min, n + 1)) if ragged else n
        P1 = np.stack([rng.uniform(0, RESCALE[0], n1), rng.uniform(0, RESCALE[1], n1)], 1)
        A1 = delaunay_adjacency(P1)
        s1, d1, G1, H1 = incidence_from_adjacency(A1)
        if genuine and partial:
            d = min(int(partial), max(n1 - 3, 0))
            keep = np.sort(rng.permutation(n1)[:n1 - d])                # matched keypoints of image 1
            n2 = n1
            pos = rng.permutation(n2)                                   # their positions in image 2
            P2 = np.stack([rng.uniform(0, RESCALE[0], n2), rng.uniform(0, RESCALE[1], n2)], 1)
            P2[pos[:len(keep)]] = P1[keep] + rng.normal(0, jitter, (len(keep), 2))
            P2[:, 0] = np.clip(P2[:, 0], 0, RESCALE[0] - 1e-3)
            P2[:, 1] = np.clip(P2[:, 1], 0, RESCALE[1] - 1e-3)
            perm = np.zeros((n1, n2), dtype=np.float32)
            perm[keep, pos[:len(keep)]] = 1
            G2, H2 = perm.T @ G1, perm.T @ H1           # gmdataset.py:349-350
            A2 = G2 @ H2.T
            colnode = lambda M: np.where(M.sum(0) > 0, M.argmax(0), -1)
            s2g, d2g = colnode(G2).astype(np.int64), colnode(H2).astype(np.int64)   # -1: all-zero column
        elif genuine:
            n2 = n1
            P2 = P1 + rng.normal(0, jitter, P1.shape)
            P2[:, 0] = np.clip(P2[:, 0], 0, RESCALE[0] - 1e-3)
            P2[:, 1] = np.clip(P2[:, 1], 0, RESCALE[1] - 1e-3)
            perm = np.eye(n1, dtype=np.float32)
            G2, H2 = perm.T @ G1, perm.T @ H1           # gmdataset.py:349-350
            A2 = G2 @ H2.T
            s2g, d2g = s1.copy(), d1.copy()             # column order of G2/H2 == graph 1's
        else:
            n2 = int(rng.randint(n_min, n + 1)) if ragged else n
            P2 = np.stack([rng.uniform(0, RESCALE[0], n2), rng.uniform(0, RESCALE[1], n2)], 1)
            A2 = delaunay_adjacency(P2)
            s2g, d2g, G2, H2 = incidence_from_adjacency(A2)
            perm = np.zeros((n1, n2), dtype=np.float32)
        s2, d2 = np.nonzero(A2)                          # PyG edge order = nonzero(A2)
        P1s.append(P1.astype(np.float32)); P2s.append(P2.astype(np.float32))
        g1s.append(pyg_like_graph(s1, d1, P1)); g2s.append(pyg_like_graph(s2, d2, P2))
        G1s.append(G1); H1s.append(H1); G2s.append(G2); H2s.append(H2)
        perms.append(perm); labels.append(1.0 if genuine else 0.0)
        e1s.append(len(s1)); e2s.append(len(s2g))
        edges1.append((s1, d1)); edges2.append((s2g, d2g))

    ns1 = torch.tensor([p.shape[0] for p in P1s], dtype=torch.long)
    ns2 = torch.tensor([p.shape[0] for p in P2s], dtype=torch.long)
    n1max, n2max = int(ns1.max()), int(ns2.max())
    e1max, e2max = max(e1s), max(e2s)

    def edge_table(edges, emax):
        t = torch.full((batch_size, 2, emax), -1, dtype=torch.int32)
        for b, (s, d) in enumerate(edges):
            t[b, 0, :len(s)] = torch.from_numpy(s.astype(np.int32))
            t[b, 1, :len(d)] = torch.from_numpy(d.astype(np.int32))
        return t

    data = {
        "Ps": [_pad_stack(P1s), _pad_stack(P2s)],
        "ns": [ns1, ns2],
        "es": [torch.tensor(e1s), torch.tensor(e2s)],
        "gt_perm_mat": _pad_stack(perms),
        "pyg_graphs": [GraphBatch.from_data_list(g1s), GraphBatch.from_data_list(g2s)],
        "label": torch.tensor(labels, dtype=torch.float32),
        "batch_size": batch_size,
        "num_graphs": 2,
        # compact form of Gs/Hs: per pair the (src, dst) node of every G/H column, -1 padded.
        "edge_lists": [edge_table(edges1, e1max), edge_table(edges2, e2max)],
    }
    if with_dense_gh:
        data["Gs"] = [_pad_stack(G1s), _pad_stack(G2s)]
        data["Hs"] = [_pad_stack(H1s), _pad_stack(H2s)]
        data["As"] = [_pad_stack([g @ h.T for g, h in zip(G1s, H1s)]),
                      _pad_stack([g @ h.T for g, h in zip(G2s, H2s)])]
    else:
        data["As"] = [None, None]
    if with_kron:
        add_kron(data)
    if with_fmaps:
        g = torch.Generator().manual_seed(seed + 77 if fmap_seed is None else fmap_seed)
        data["fmaps"] = [
            (torch.randn(batch_size, 256, 15, 20, generator=g),
             torch.randn(batch_size, 512, 8, 10, generator=g))
            for _ in range(2)
        ]
        if fmap_noise is not None:
            a, b = data["fmaps"][0]
            data["fmaps"][1] = (a + fmap_noise * data["fmaps"][1][0], b + fmap_noise * data["fmaps"][1][1])
    return data


def add_kron(data: dict) -> dict:
    """Attach ``KGHs_sparse`` (the CSC ``.indices`` of kron(G2,G1) / kron(H2,H1), gmdataset.py:623-642) computed from the
    batch's edge tables - for batches that were generated or sharded without it (only the CPU oracle reads it)."""
    t1, t2 = data["edge_lists"]
    n1max = data["Ps"][0].shape[1]
    kgh = []
    for b in range(t1.shape[0]):
        # CSC .indices of kron(G2,G1): one entry per NON-ZERO column t = k2*e1max + k1 (G2 column k2 and G1 column k1
        # both non-zero), in column order, value = row i2*n1max + i1; kron(H2,H1) likewise with ITS OWN non-zero
        # columns (csx_matrix.py:39-41 eliminates zeros per matrix) - the two lists pair up only if the same
        # columns survive in G and H
        s1, d1 = t1[b, 0][t1[b, 0] >= 0].long(), t1[b, 1][t1[b, 1] >= 0].long()
        s2, d2 = t2[b, 0][t2[b, 0] >= 0].long(), t2[b, 1][t2[b, 1] >= 0].long()
        kgh.append(((s2[:, None] * n1max + s1[None, :]).reshape(-1), (d2[:, None] * n1max + d1[None, :]).reshape(-1)))
    data["KGHs_sparse"] = kgh
    return data


def clone_batch(data: dict) -> dict:
    """Deep copy of the tensors/graphs so one batch can feed two implementations."""
    def cp(v):
        if isinstance(v, torch.Tensor):
            return v.clone()
        if isinstance(v, GraphBatch):
            return GraphBatch(v.x.clone(), v.edge_index.clone(), v.edge_attr.clone(),
                              v.ptr.clone(), v.eptr.clone())
        if isinstance(v, (list, tuple)):
            return type(v)(cp(x) for x in v)
        return v
    return {k: cp(v) for k, v in data.items()}


def batch_to(data: dict, device, non_blocking: bool = False) -> dict:
    """The reference's ``data_to_cuda`` (``utils/data_to_cuda.py:5-33``) for this dict layout."""
    def mv(v):
        if isinstance(v, torch.Tensor):
            return v.to(device, non_blocking=non_blocking)
        if isinstance(v, GraphData):
            return v.to(device, non_blocking)
        if isinstance(v, (list, tuple)):
            return type(v)(mv(x) for x in v)
        return v
    return {k: mv(v) for k, v in data.items()}
