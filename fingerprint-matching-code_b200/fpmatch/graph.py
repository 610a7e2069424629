"""Minimal stand-ins for the two torch_geometric containers the matching head touches.

The reference hands ``Net.forward`` a ``torch_geometric.data.Batch`` per image
(``/root/reference/src/gmdataset.py:183-188,606-607``) and reads ``x``, ``edge_index``,
``edge_attr`` and ``to_data_list()`` from it (``src/model/spline_conv.py:28-41,66-81``).
torch_geometric is not part of this image, so the synthetic generator builds these
light objects instead; a real PyG ``Batch`` is accepted wherever a ``GraphBatch`` is
(the head only uses the attributes below plus ``batch``/``ptr``).
"""
from __future__ import annotations

from typing import List, Optional

import torch


class GraphData:
    """One keypoint graph: ``x [n, C]``, ``edge_index [2, e]`` (src, dst), ``edge_attr [e, 2]``."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, hyperedge_index=None):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.hyperedge_index = hyperedge_index

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    def to(self, device, non_blocking: bool = False):
        """Returns a NEW graph on ``device`` (the source object is left untouched)."""
        mv = lambda v: v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
        return GraphData(mv(self.x), mv(self.edge_index), mv(self.edge_attr), mv(self.hyperedge_index))


class GraphBatch(GraphData):
    """Disjoint union of graphs, PyG style: node rows concatenated, edge indices offset.

    ``ptr [B+1]`` holds node offsets, ``eptr [B+1]`` edge offsets, ``batch [sum n]`` the graph
    id of every node (same meaning as PyG's ``Batch.batch``).
    """

    def __init__(self, x, edge_index, edge_attr, ptr, eptr):
        super().__init__(x, edge_index, edge_attr)
        self.ptr = ptr
        self.eptr = eptr
        self.num_graphs = int(ptr.numel() - 1)

    @property
    def batch(self) -> torch.Tensor:
        counts = self.ptr[1:] - self.ptr[:-1]
        return torch.repeat_interleave(
            torch.arange(self.num_graphs, device=self.ptr.device), counts)

    @staticmethod
    def from_data_list(graphs: List[GraphData]) -> "GraphBatch":
        ns = [g.num_nodes for g in graphs]
        es = [int(g.edge_index.shape[1]) for g in graphs]
        ptr = torch.zeros(len(graphs) + 1, dtype=torch.long)
        eptr = torch.zeros(len(graphs) + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.tensor(ns, dtype=torch.long), 0)
        eptr[1:] = torch.cumsum(torch.tensor(es, dtype=torch.long), 0)
        x = torch.cat([g.x for g in graphs], 0)
        ei = torch.cat([g.edge_index + int(ptr[i]) for i, g in enumerate(graphs)], 1)
        ea = torch.cat([g.edge_attr for g in graphs], 0)
        return GraphBatch(x, ei, ea, ptr, eptr)

    def to_data_list(self) -> List[GraphData]:
        out = []
        for b in range(self.num_graphs):
            n0, n1 = int(self.ptr[b]), int(self.ptr[b + 1])
            e0, e1 = int(self.eptr[b]), int(self.eptr[b + 1])
            out.append(GraphData(self.x[n0:n1], self.edge_index[:, e0:e1] - n0,
                                 self.edge_attr[e0:e1]))
        return out

    def to(self, device, non_blocking: bool = False):
        mv = lambda v: v.to(device, non_blocking=non_blocking)
        return GraphBatch(mv(self.x), mv(self.edge_index), mv(self.edge_attr), mv(self.ptr), mv(self.eptr))


def graph_offsets(g, device=None):
    """Return ``(ptr, eptr)`` (int64, [B+1]) for a GraphBatch or a duck-typed PyG Batch."""
    if hasattr(g, "ptr") and getattr(g, "eptr", None) is not None:
        return g.ptr, g.eptr
    batch = g.batch
    nb = int(batch.max().item()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=nb)
    ptr = torch.zeros(nb + 1, dtype=torch.long, device=batch.device)
    ptr[1:] = torch.cumsum(counts, 0)
    ebatch = batch[g.edge_index[0]]
    ecounts = torch.bincount(ebatch, minlength=nb)
    eptr = torch.zeros(nb + 1, dtype=torch.long, device=batch.device)
    eptr[1:] = torch.cumsum(ecounts, 0)
    return ptr, eptr
