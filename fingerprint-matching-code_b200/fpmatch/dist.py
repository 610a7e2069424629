"""Data-parallel plumbing: fingerprint pairs are independent, so a batch shards across ranks by contiguous
slices of the pair dimension with NO collective on the data path.  The only exchange
is the gather of per-pair results for metrics (what evaluate_binary_classifier.py accumulates at
``evaluate_binary_classifier.py:98-103``) - NCCL on GPUs, gloo in the CPU tests.  Training adds the gradient
all-reduce: ``wrap_ddp`` puts the model under ``torch.nn.parallel.DistributedDataParallel`` (bucketed NCCL
all-reduce overlapped with the hand-written backward kernels), replacing the reference's disabled single-process
``nn.DataParallel`` subclass (``/root/reference/src/parallel/data_parallel.py:6-17``, ``train.py:148``).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist

from .graph import GraphBatch


def shard_bounds(batch_size: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of the pair dimension owned by ``rank``."""
    base, rem = divmod(batch_size, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(data: dict, rank: int, world: int) -> dict:
    """Slice every per-pair entry of a ``data_dict`` (tensors, the two graph batches, per-pair lists)."""
    B = data["gt_perm_mat"].shape[0]
    lo, hi = shard_bounds(B, rank, world)

    def cut(v):
        if isinstance(v, torch.Tensor):
            return v[lo:hi] if v.dim() > 0 and v.shape[0] == B else v
        if isinstance(v, GraphBatch):
            return GraphBatch.from_data_list(v.to_data_list()[lo:hi])
        if isinstance(v, (list, tuple)):              # per-graph containers (Ps, ns, fmaps, pyg_graphs, ...)
            return type(v)(cut(x) for x in v)
        return v

    PER_PAIR_LISTS = ("KGHs_sparse", "cls", "id_list")
    out = {k: (v[lo:hi] if k in PER_PAIR_LISTS and isinstance(v, list) else cut(v)) for k, v in data.items()}
    out["batch_size"] = hi - lo
    # padded widths may shrink inside a shard; keep the global padding so shapes agree across ranks
    return out


def gather_pairs(local: Dict[str, torch.Tensor], batch_size: int) -> Dict[str, torch.Tensor]:
    """All-gather per-pair result tensors (dim 0 = pairs) back into global pair order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(batch_size, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    out = {}
    for k, t in local.items():
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        out[k] = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
    return out


def wrap_ddp(net: torch.nn.Module, device: torch.device, static_graph: bool = True,
             bucket_cap_mb: int = 64) -> torch.nn.Module:
    """One process per GPU: gradients are averaged over ranks by NCCL all-reduce during ``backward``.

    ``find_unused_parameters`` is on because the reference model carries parameters its forward never touches (the
    GCNConv of every PYGNNLayer, gnn.py:198; the k-branch in stage 1; edge_affinity, whose output SAGEConv drops)."""
    from torch.nn.parallel import DistributedDataParallel
    # static_graph: the set of unused parameters is the same every step, so DDP learns it in the first iteration instead
    # of walking the autograd graph after every forward (the 8-pair step is bound by host launch time: r2s2 measured
    # 1.9 ms of exposed DDP time per step against 0.27 ms for the bare 126 MB all-reduce).  gradient_as_bucket_view:
    # gradients live in the communication buckets, no copy in / out around the all-reduce.
    return DistributedDataParallel(net, device_ids=[device.index] if device.type == "cuda" else None,
                                   find_unused_parameters=True, broadcast_buffers=False, static_graph=static_graph,
                                   gradient_as_bucket_view=True, bucket_cap_mb=bucket_cap_mb)
