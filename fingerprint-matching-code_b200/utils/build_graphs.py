"""Drop-in for ``/root/reference/utils/build_graphs.py`` - same names and return values, computed on the GPU.

The reference triangulates with scipy/Qhull and fills ``G``/``H`` in a python double loop
(``build_graphs.py:60-72,78-100``); these functions keep the numpy-in / numpy-out signatures so the dataset
code can call them unchanged, but run the batched kernels of ``csrc/graph_build.cu`` (one graph = a batch of
one).  Pipelines that already hold the keypoints on the device should call ``fpmatch.graph_build`` directly and
skip the host round trip.  Degenerate point sets (exactly co-circular quadrilaterals, duplicates) follow the
rules documented in ``oracle/graphs.py``; Qhull's choice there is an artefact of its merge order.
"""
from typing import Tuple

import numpy as np
import torch
from torch import Tensor

from fpmatch import graph_build as _gb


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("fpmatch: build_graphs needs a CUDA device (no CPU fallback exists)")
    return torch.device("cuda", torch.cuda.current_device())


def _adjacency(P: np.ndarray, stg: str, thre=None) -> Tensor:
    dev = _device()
    Pt = torch.as_tensor(np.ascontiguousarray(P, dtype=np.float64)).to(dev)[None]
    ns = torch.tensor([P.shape[0]], dtype=torch.int64, device=dev)
    if stg == "near" and thre is None:
        stg = "fc"
    return _gb.graph_adjacency(Pt, ns, stg, 0.0 if thre is None else float(thre)), Pt, ns


def build_graphs(P_np: np.ndarray, n: int, n_pad: int = None, edge_pad: int = None, stg: str = 'fc',
                 sym: bool = True, thre: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    r"""
    Build graph matrix :math:`\mathbf G, \mathbf H` from point set :math:`\mathbf P`
    (:math:`\mathbf A = \mathbf G \cdot \mathbf H^\top`).

    :param P_np: :math:`(n\times 2)` point set containing point coordinates
    :param n: number of exact points in the point set
    :param n_pad: padded node length
    :param edge_pad: padded edge length
    :param stg: ``fc`` (fully connected), ``near`` (edges longer than ``thre`` removed) or ``tri`` (Delaunay)
    :param sym: True for a symmetric adjacency, False for half adjacency (G/H list the upper half only)
    :param thre: threshold of the ``near`` strategy
    :return: :math:`A`, :math:`G`, :math:`H`, edge_num
    """
    assert stg in ('fc', 'tri', 'near'), 'No strategy named {} found.'.format(stg)
    A, Pt, ns = _adjacency(P_np[0:n, :], stg, thre if stg == 'near' else None)
    edge_num = int(A.sum().item())
    assert n > 0 and edge_num > 0, 'Error in n = {} and edge_num = {}'.format(n, edge_num)
    if n_pad is None:
        n_pad = n
    if edge_pad is None:
        edge_pad = edge_num
    assert n_pad >= n
    assert edge_pad >= edge_num
    edges = _gb.graph_edges(A, Pt, ns, upper_only=not sym)
    G, H = _gb.incidence_dense(edges.edge_list, n_pad, edge_pad)
    return A[0].double().cpu().numpy(), G[0].cpu().numpy(), H[0].cpu().numpy(), edge_num


def delaunay_triangulate(P: np.ndarray) -> np.ndarray:
    r"""Adjacency matrix of the Delaunay triangulation of ``P`` (fully connected below 3 points or when the
    points are collinear - the reference's QhullError fallback)."""
    return _adjacency(P, 'tri')[0][0].double().cpu().numpy()


def fully_connect(P: np.ndarray, thre=None) -> np.ndarray:
    r"""Adjacency matrix of the fully connected graph; edges longer than ``thre`` are removed."""
    return _adjacency(P, 'near' if thre is not None else 'fc', thre)[0][0].double().cpu().numpy()


def make_grids(start, stop, num) -> np.ndarray:
    r"""Cell-centre grid points (host helper, pure index arithmetic; build_graphs.py:122-143)."""
    assert len(start) == len(stop) == len(num)
    length = int(np.prod(num))
    P = np.zeros((length, len(num)), dtype=np.float32)
    for axis, (begin, end, cnt) in enumerate(zip(start, stop, num)):
        edges = np.linspace(begin, end, cnt + 1)
        centres = edges[1:] - (edges[1] - edges[0]) / 2           # same roundings as the reference's shift-then-slice
        # the reference's np.repeat([g], length / n, axis=i): axis 0 cycles through the values, axis 1 holds each
        P[:, axis] = np.tile(centres, length // cnt) if axis == 0 else np.repeat(centres, length // cnt)
    return P


def reshape_edge_feature(F: Tensor, G: Tensor, H: Tensor, device=None) -> Tensor:
    r"""
    Edge feature matrix :math:`\mathbf{X}_{e_{ij}} = concat(\mathbf{F}_i, \mathbf{F}_j)` arranged by the
    columns of :math:`G`, :math:`H`.  ``F`` is :math:`(b\times d \times n)`, ``G``/``H`` :math:`(b\times n \times e)`;
    returns :math:`(b \times 2d \times e)`.
    """
    if device is None:
        device = F.device
    return torch.cat((torch.matmul(F, G), torch.matmul(F, H)), dim=1).to(device)
