"""Batched keypoint-graph construction on the GPU.

The reference builds the inputs of the matching head on the host, image by image
(``/root/reference/utils/build_graphs.py:12-119``, ``/root/reference/src/gmdataset.py:169-189,345-352,563-672``):
scipy Delaunay, python loops for ``G``/``H``, numpy for the PyG edge list, an O(n^3) hyper-edge list and scipy
``kron`` index lists (92 MB per pair at 400 keypoints).  ``build_graph_batch`` produces the same adjacency, edge
order and pseudo-coordinates for a whole padded batch from the keypoint tensor already on the device, and
``collate_pairs`` assembles the ``data_dict`` ``Net.forward`` consumes - the head then runs from
``(feature maps, keypoints, counts)`` alone.  Kernels: ``csrc/graph_build.cu``.  There is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from .graph import GraphBatch
from .ops import _chk, _count, _i64, _stream, on_tensor_device

RESCALE = 320.0                       # max(RESCALE), /root/reference/src/gmdataset.py:36-48,171
_STG = {"fc": 0, "tri": 1, "near": 2}


def _points64(P: Tensor) -> Tensor:
    if not isinstance(P, Tensor) or not P.is_cuda:
        raise RuntimeError("fpmatch: keypoints must be a CUDA tensor (no CPU fallback exists)")
    if P.dim() != 3 or P.shape[-1] != 2:
        raise ValueError("keypoints must be [B, nmax, 2]")
    return P.to(torch.float64).contiguous()


@on_tensor_device
def graph_adjacency(P: Tensor, ns: Tensor, stg: str = "tri", thre: float = 0.0) -> Tensor:
    """``A [B,nmax,nmax]`` fp32 0/1: ``delaunay_triangulate`` / ``fully_connect`` (build_graphs.py:78-119)."""
    assert stg in _STG, "No strategy named {} found.".format(stg)      # build_graphs.py:43
    P = _points64(P)
    ns = _i64(ns)
    B, nmax = P.shape[0], P.shape[1]
    A = torch.empty(B, nmax, nmax, device=P.device, dtype=torch.float32)
    rc = _lib.lib().fpm_graph_adjacency(P.data_ptr(), _chk(ns, "ns", torch.int64), A.data_ptr(), B, nmax,
                                        _STG[stg], float(thre), _stream())
    _lib.check(rc, "fpm_graph_adjacency")
    _count()
    return A


@dataclass
class EdgeSet:
    """Edges of a batch of adjacency matrices in ``np.nonzero`` (row-major) order."""
    edge_index: Tensor      # [2, E] int64, node ids offset by ptr[b]
    edge_attr: Tensor       # [E, 2] fp32 pseudo-coordinates
    x: Tensor               # [sum n, 2] fp32 = P / 320
    ptr: Tensor             # [B+1] int64 node offsets
    eptr: Tensor            # [B+1] int64 edge offsets
    edge_list: Tensor       # [B, 2, emax] int32 pair-local (src, dst), -1 padded
    es: List[int]           # edges per graph (host)


@on_tensor_device
def graph_edges(A: Tensor, P: Tensor, ns: Tensor, upper_only: bool = False, emax: Optional[int] = None) -> EdgeSet:
    """Edge list, pseudo-coordinates and node coordinates of ``to_pyg_graph`` (gmdataset.py:169-189) plus the
    (src, dst) table that stands for the one-hot ``G``/``H`` columns (build_graphs.py:60-72; ``upper_only`` =
    ``sym=False``).  One host read (the per-graph edge counts) sizes the outputs."""
    P = _points64(P)
    ns = _i64(ns)
    B, nmax = P.shape[0], P.shape[1]
    L = _lib.lib()
    rowcnt = torch.empty(B * nmax, device=P.device, dtype=torch.int32)
    rc = L.fpm_graph_row_counts(_chk(A, "A"), ns.data_ptr(), rowcnt.data_ptr(), B, nmax, int(upper_only), _stream())
    _lib.check(rc, "fpm_graph_row_counts")
    rowoff = torch.zeros(B * nmax + 1, device=P.device, dtype=torch.int64)
    torch.cumsum(rowcnt, 0, out=rowoff[1:])
    ptr = torch.zeros(B + 1, device=P.device, dtype=torch.int64)
    torch.cumsum(ns.clamp(max=nmax), 0, out=ptr[1:])
    eptr = rowoff[::nmax].contiguous()
    host = torch.cat([eptr, ptr[-1:]]).tolist()                   # the one synchronising read
    eoffs, total_nodes = host[:-1], host[-1]
    es = [eoffs[b + 1] - eoffs[b] for b in range(B)]
    E = eoffs[-1]
    emax = max(es + [0]) if emax is None else emax
    assert emax >= max(es + [0])                                   # build_graphs.py:57
    edge_index = torch.empty(2, E, device=P.device, dtype=torch.int64)
    edge_attr = torch.empty(E, 2, device=P.device, dtype=torch.float32)
    x = torch.empty(total_nodes, 2, device=P.device, dtype=torch.float32)
    edge_list = torch.full((B, 2, emax), -1, device=P.device, dtype=torch.int32)
    rc = L.fpm_graph_edges(A.data_ptr(), P.data_ptr(), ns.data_ptr(), ptr.data_ptr(), rowoff.data_ptr(),
                           edge_index.data_ptr(), edge_attr.data_ptr(), x.data_ptr(), edge_list.data_ptr(), B, nmax,
                           E, emax, int(upper_only), RESCALE, _stream())
    _lib.check(rc, "fpm_graph_edges")
    _count(2)
    return EdgeSet(edge_index, edge_attr, x, ptr, eptr, edge_list, es)


@on_tensor_device
def permute_graph(A1: Tensor, perm: Tensor, edge_list1: Optional[Tensor], n2max: Optional[int] = None):
    """Graph 2 of a genuine pair: ``G2 = perm^T G1``, ``H2 = perm^T H1``, ``A2 = G2 H2^T`` (gmdataset.py:345-352).
    ``perm [B,n1max,n2max]`` is a (partial) permutation; returns ``(A2, edge_list2)`` with unmatched ends = -1."""
    B, n1max = A1.shape[0], A1.shape[1]
    n2max = perm.shape[2] if n2max is None else n2max
    has, arg = perm.max(dim=2)
    mp = torch.where(has > 0, arg, torch.full_like(arg, -1)).to(torch.int32).contiguous()
    A2 = torch.zeros(B, n2max, n2max, device=A1.device, dtype=torch.float32)
    el2 = torch.empty_like(edge_list1) if edge_list1 is not None else None
    emax = edge_list1.shape[2] if edge_list1 is not None else 0
    rc = _lib.lib().fpm_graph_permute(_chk(A1, "A1"), mp.data_ptr(), _chk(edge_list1, "edge_list1", torch.int32),
                                      A2.data_ptr(), el2.data_ptr() if el2 is not None else None, B, n1max, n2max,
                                      emax, _stream())
    _lib.check(rc, "fpm_graph_permute")
    _count()
    return A2, el2


@on_tensor_device
def incidence_dense(edge_list: Tensor, n_pad: int, edge_pad: Optional[int] = None):
    """Dense one-hot ``G, H [B, n_pad, edge_pad]`` from the (src, dst) table (build_graphs.py:60-72)."""
    B, emax = edge_list.shape[0], edge_list.shape[2]
    edge_pad = emax if edge_pad is None else edge_pad
    assert edge_pad >= emax                                         # build_graphs.py:57
    G = torch.zeros(B, n_pad, edge_pad, device=edge_list.device, dtype=torch.float32)
    H = torch.zeros_like(G)
    rc = _lib.lib().fpm_graph_incidence(_chk(edge_list, "edge_list", torch.int32), G.data_ptr(), H.data_ptr(), B,
                                        emax, n_pad, edge_pad, _stream())
    _lib.check(rc, "fpm_graph_incidence")
    _count()
    return G, H


@on_tensor_device
def kron_index_lists(edge_list1: Tensor, edge_list2: Tensor, es1: Sequence[int], es2: Sequence[int], n1max: int):
    """``KGHs_sparse``: per pair ``(idxG, idxH)``, the row of the single one in every column of
    ``kron(G2, G1)`` / ``kron(H2, H1)`` (gmdataset.py:623-642).  Views into two flat int64 buffers.

    Each Kronecker product drops its own all-zero columns (csx_matrix.py:39-41), so with a partial permutation
    (``-1`` ends in ``edge_list2``) the sources and the targets are compacted independently, exactly as the
    reference's lists are; ``es1`` / ``es2`` are then ignored in favour of the surviving column counts."""
    B = edge_list1.shape[0]
    dev = edge_list1.device
    if bool(((edge_list1[:, 0] < 0) != (edge_list1[:, 1] < 0)).any()) or \
            bool(((edge_list2[:, 0] < 0) != (edge_list2[:, 1] < 0)).any()):
        def compact(t):
            order = torch.argsort((t < 0).to(torch.int8), dim=-1, stable=True)
            return torch.gather(t, -1, order)
        edge_list1, edge_list2 = compact(edge_list1).contiguous(), compact(edge_list2).contiguous()
        c1, c2 = (edge_list1 >= 0).sum(-1), (edge_list2 >= 0).sum(-1)              # [B, 2] surviving G / H columns
        if not (torch.equal(c1[:, 0], c1[:, 1]) and torch.equal(c2[:, 0], c2[:, 1])):
            raise NotImplementedError("kron_index_lists: G and H keep different numbers of columns (asymmetric "
                                      "adjacency); the reference's two lists then have different lengths")
        es1, es2 = c1[:, 0].tolist(), c2[:, 0].tolist()
    sizes = [a * b for a, b in zip(es1, es2)]
    koff_host = [0]
    for s in sizes:
        koff_host.append(koff_host[-1] + s)
    koff = torch.tensor(koff_host, dtype=torch.int64).to(dev)
    e1 = torch.tensor(list(es1), dtype=torch.int64).to(dev)
    e2 = torch.tensor(list(es2), dtype=torch.int64).to(dev)
    idxG = torch.empty(koff_host[-1], device=dev, dtype=torch.int64)
    idxH = torch.empty_like(idxG)
    rc = _lib.lib().fpm_graph_kron_index(_chk(edge_list1, "edge_list1", torch.int32),
                                         _chk(edge_list2, "edge_list2", torch.int32), e1.data_ptr(), e2.data_ptr(),
                                         koff.data_ptr(), idxG.data_ptr(), idxH.data_ptr(), B, edge_list1.shape[2],
                                         edge_list2.shape[2], n1max, _stream())
    _lib.check(rc, "fpm_graph_kron_index")
    _count()
    return [(idxG[koff_host[b]:koff_host[b + 1]], idxH[koff_host[b]:koff_host[b + 1]]) for b in range(B)]


@dataclass
class BuiltGraphs:
    A: Tensor               # [B, nmax, nmax]
    graph: GraphBatch       # PyG-style batch (x = P/320, edge_index, edge_attr) in nonzero(A) order
    edge_list: Tensor       # [B, 2, emax] int32 columns of G / H
    es: List[int]


@on_tensor_device
def build_graph_batch(P: Tensor, ns: Tensor, stg: str = "tri", sym: bool = True, thre: float = 0.0) -> BuiltGraphs:
    """``build_graphs`` + ``to_pyg_graph`` for a padded batch ``P [B,nmax,2]``, ``ns [B]``."""
    A = graph_adjacency(P, ns, stg, thre)
    e = graph_edges(A, P, ns)
    gb = GraphBatch(e.x, e.edge_index, e.edge_attr, e.ptr, e.eptr)
    if sym:
        return BuiltGraphs(A, gb, e.edge_list, e.es)
    half = graph_edges(A, P, ns, upper_only=True)
    return BuiltGraphs(A, gb, half.edge_list, half.es)


@on_tensor_device
def collate_pairs(P1: Tensor, P2: Tensor, ns1: Tensor, ns2: Tensor, gt_perm_mat: Optional[Tensor] = None,
                  label: Optional[Tensor] = None, fmaps=None, images=None, stg: str = "tri",
                  tgt_stg: str = "same", with_dense_gh: bool = False, with_kron: bool = False) -> dict:
    """Device-side ``get_pair_classify`` + ``collate_fn`` (gmdataset.py:304-372,563-672) from padded keypoints.

    ``tgt_stg = 'same'`` follows the reference default: a pair whose ``gt_perm_mat`` is non-zero gets graph 2 by
    carrying graph 1's topology through the permutation; a pair with an all-zero ``gt_perm_mat`` (imposter) gets
    its own triangulation.  Any other value triangulates graph 2 with that strategy for every pair.
    """
    dev = P1.device
    B = P1.shape[0]
    n1max, n2max = P1.shape[1], P2.shape[1]
    if gt_perm_mat is None:
        gt_perm_mat = torch.zeros(B, n1max, n2max, device=dev)
    g1 = build_graph_batch(P1, ns1, stg)
    genuine = None
    if tgt_stg == "same":
        genuine = gt_perm_mat.flatten(1).sum(1) > 0                            # gmdataset.py:346
        gl = genuine.tolist()
    if genuine is not None and all(gl):
        A2, el2 = permute_graph(g1.A, gt_perm_mat, g1.edge_list, n2max)
        es2 = list(g1.es)
    else:
        own = build_graph_batch(P2, ns2, stg if tgt_stg == "same" else tgt_stg)
        A2, el2, es2 = own.A, own.edge_list, own.es
        if genuine is not None and any(gl):
            A2p, el2p = permute_graph(g1.A, gt_perm_mat, g1.edge_list, n2max)
            A2 = torch.where(genuine[:, None, None], A2p, A2)
            es2 = [g1.es[b] if gl[b] else es2[b] for b in range(B)]
            emax = max(el2.shape[2], el2p.shape[2])
            pad = lambda t: torch.nn.functional.pad(t, (0, emax - t.shape[2]), value=-1)
            el2 = torch.where(genuine[:, None, None], pad(el2p), pad(el2))[:, :, :max(es2)].contiguous()
    e2 = graph_edges(A2, P2, ns2)                                          # PyG order = nonzero(A2), gmdataset.py:244
    graph2 = GraphBatch(e2.x, e2.edge_index, e2.edge_attr, e2.ptr, e2.eptr)
    data = {
        "Ps": [P1.float(), P2.float()],
        "ns": [_i64(ns1), _i64(ns2)],
        "es": [torch.tensor(g1.es), torch.tensor(es2)],
        "gt_perm_mat": gt_perm_mat,
        "As": [g1.A, A2],
        "pyg_graphs": [g1.graph, graph2],
        "edge_lists": [g1.edge_list, el2],
        "batch_size": B,
        "num_graphs": 2,
    }
    if label is not None:
        data["label"] = label
    if fmaps is not None:
        data["fmaps"] = fmaps
    if images is not None:
        data["images"] = images
    if with_dense_gh:
        G1, H1 = incidence_dense(g1.edge_list, n1max)
        G2, H2 = incidence_dense(el2, n2max)
        data["Gs"], data["Hs"] = [G1, G2], [H1, H2]
    if with_kron:
        data["KGHs_sparse"] = kron_index_lists(g1.edge_list, el2, g1.es, es2, n1max)
    return data
