"""Drop-in for ``/root/reference/utils/data_to_cuda.py`` (``data_to_cuda``): recursive host->device move of a
collated batch (lists, tuples, dicts, tensors, CSR/CSC containers, graph batches).

Beyond the reference's signature, ``mover`` replaces the per-tensor ``.cuda()`` call - the prefetcher
(``fpmatch.prefetch.CudaPrefetcher``) passes one that copies asynchronously into persistent device buffers."""
import torch

from fpmatch.graph import GraphBatch, GraphData
from src.sparse_torch.csx_matrix import CSCMatrix3d, CSRMatrix3d


def data_to_cuda(inputs, device="cuda", mover=None):
    """Move every tensor-like element of ``inputs`` to the GPU.  Containers are updated in place and returned, as in
    the reference; tuples come back as lists (data_to_cuda.py:14-17)."""
    mv = mover if mover is not None else (lambda t: t.to(device))
    if type(inputs) is list:
        for i, x in enumerate(inputs):
            inputs[i] = data_to_cuda(x, device, mover)
    elif type(inputs) is tuple:
        inputs = [data_to_cuda(x, device, mover) for x in inputs]
    elif type(inputs) is dict:
        for key in inputs:
            inputs[key] = data_to_cuda(inputs[key], device, mover)
    elif inputs is None or type(inputs) in (str, int, float, bool):
        pass
    elif isinstance(inputs, torch.Tensor):
        inputs = mv(inputs)
    elif isinstance(inputs, (CSRMatrix3d, CSCMatrix3d)):
        inputs = inputs.__class__([mv(inputs.indices), mv(inputs.indptr), mv(inputs.data)], shape=inputs.shape)
    elif isinstance(inputs, GraphBatch):
        inputs = GraphBatch(mv(inputs.x), mv(inputs.edge_index), mv(inputs.edge_attr), mv(inputs.ptr), mv(inputs.eptr))
    elif isinstance(inputs, GraphData):
        inputs = GraphData(mv(inputs.x), mv(inputs.edge_index), mv(inputs.edge_attr))
    elif hasattr(inputs, "edge_index") and hasattr(inputs, "to"):        # a torch_geometric Data / Batch
        inputs = inputs.to(device)
    else:
        raise TypeError('Unknown type of inputs: {}'.format(type(inputs)))
    return inputs
