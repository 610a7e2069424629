"""Drop-in for ``/root/reference/src/model/spline_conv.py`` (``SConv``, ``SiameseSConvOnNodes``,
``SiameseNodeFeaturesToEdgeFeatures``) with its own ``SplineConv`` (torch_geometric is not required).

``SplineConv`` keeps torch_geometric 1.6.3's parameters (``weight [25, in, out]``, ``root [in, out]``,
``bias [out]``) and initialisation, so reference checkpoints load; a PyG 2.x checkpoint's ``lin.weight``
is accepted as ``root``.  Forward = one dense GEMM against all 26 weight slabs + one gather/max kernel
(``csrc/spline.cu``) instead of a per-edge weighting kernel and a scatter.
"""
import math

import torch
import torch.nn
import torch.nn.functional as F

from fpmatch import ops
from fpmatch.graph import GraphData, graph_offsets


def slab_plan_enabled() -> bool:
    """The slab-sparse GEMM runs on the persistent fp16x3 kernel; other GEMM modes keep the dense slab product."""
    return ops.gemm_mode() == "3xf16" and ops.gemm_pair_enabled() and ops.slab_plan_enabled()


class SplineConv(torch.nn.Module):
    """SplineConv(in, out, dim=2, kernel_size=5, is_open_spline=True, degree=1, aggr='max',
    root_weight=True, bias=True): the configuration of ``spline_conv.py:17``; others are rejected."""

    def __init__(self, in_channels, out_channels, dim=2, kernel_size=5, is_open_spline=True, degree=1,
                 aggr="max", root_weight=True, bias=True):
        super().__init__()
        if dim != 2 or degree != 1 or not is_open_spline or aggr != "max" or not root_weight or not bias:
            raise NotImplementedError("only the configuration used by the matching head is implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.dim, self.degree, self.kernel_size = dim, degree, kernel_size
        K = kernel_size ** dim
        self.weight = torch.nn.Parameter(torch.empty(K, in_channels, out_channels))
        self.root = torch.nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = torch.nn.Parameter(torch.empty(out_channels))
        self._packed = None      # (key, [K+1)*out, in] slab matrix)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.in_channels * self.weight.size(0))
        torch.nn.init.uniform_(self.weight, -bound, bound)
        torch.nn.init.uniform_(self.root, -bound, bound)
        torch.nn.init.zeros_(self.bias)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        k2 = prefix + "lin.weight"                       # torch_geometric 2.x name of the root weight
        if k2 in state_dict and prefix + "root" not in state_dict:
            state_dict[prefix + "root"] = state_dict.pop(k2).t()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def packed_weight(self) -> torch.Tensor:
        """[(K+1)*out, in]: row k*out + o holds W[k][:, o]; slab K is the root weight."""
        key = (self.weight.data_ptr(), self.weight._version, self.root.data_ptr(), self.root._version,
               self.weight.device)
        if self._packed is None or self._packed[0] != key:
            w = torch.cat([self.weight.detach(), self.root.detach().unsqueeze(0)], 0)
            self._packed = (key, w.permute(0, 2, 1).reshape(-1, self.in_channels).contiguous())
        return self._packed[1]

    def forward(self, x, edge_index, pseudo, csr=None, ptr=None, eptr=None, mode=None, residual=None, plan=None,
                presplit=None, split_out=None):
        """``csr`` = in-edge lists from ``ops.csr_by_dst``; built on the fly when missing.  ``mode``:
        None / 2 = plain conv output, 0 = relu(out), 1 = residual + 0.1 * out.  ``plan`` = ``ops.SlabPlan`` of the
        graph (shared by the layers of an SConv): only the (node, slab) products some edge reads are computed.
        ``split_out`` (with a plan): operand buffers of the NEXT layer's slab GEMM; the result is written there as its
        fp16 split and not as an fp32 tensor (returns None).  ``presplit``: such buffers holding THIS layer's input
        (``x`` is then unused)."""
        total = x.shape[0] if x is not None else plan.T
        if csr is None:
            if ptr is None:
                ptr = torch.tensor([0, total], dtype=torch.int64, device=edge_index.device)
                eptr = torch.tensor([0, edge_index.shape[1]], dtype=torch.int64, device=edge_index.device)
            max_e = int((eptr[1:] - eptr[:-1]).max())
            csr = ops.csr_by_dst(edge_index.contiguous(), ptr, eptr, total, max_e)
        if plan is None and slab_plan_enabled() and self.out_channels % 128 == 0 and self.in_channels % 8 == 0:
            plan = ops.SlabPlan(edge_index.contiguous(), pseudo.contiguous(), total, self.out_channels, self.kernel_size)
        if plan is not None:
            Y = ops.spline_slab_gemm(None if presplit is not None else x.detach().contiguous(), self.packed_weight(),
                                     plan, presplit=presplit)
        else:
            assert presplit is None and split_out is None
            Y = ops.gemm_nt(x.detach().contiguous(), self.packed_weight(), weight_operand=True)
        bias = self.bias.detach().contiguous()
        mode = 2 if mode is None else mode
        return ops.spline_gather_max(Y, residual, edge_index, pseudo.contiguous(), csr[0], csr[1], bias, mode,
                                     self.kernel_size, split_out=split_out, want_out=split_out is None)


class SConv(torch.nn.Module):
    def __init__(self, input_features, output_features):
        super(SConv, self).__init__()
        self.in_channels = input_features
        self.num_layers = 2
        self.convs = torch.nn.ModuleList()
        for _ in range(self.num_layers):
            self.convs.append(SplineConv(input_features, output_features, dim=2, kernel_size=5, aggr="max"))
            input_features = output_features
        self.out_channels = input_features
        self.reset_parameters()

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def prepare_weights(self):
        """Pack the slab matrices and (in the error-compensated fp16 mode) split them into their cached fp16 hi / lo
        operands on the CURRENT stream.  ``Net.matching_head`` calls this before it forks the two images' chains onto
        two streams: both chains read the same cached operands, which must exist before either starts."""
        for conv in self.convs:
            packed = conv.packed_weight()
            if slab_plan_enabled() or ops.gemm_mode() == "3xf16":
                ops.f16_split_rows(packed, cache=True)
            elif ops.gemm_mode() == "3xtf32":
                ops.tf32_split(packed, cache=True)

    def forward(self, data, residual_scale_input=None):
        """relu(conv0(x)) -> conv1; with ``residual_scale_input`` the x + 0.1 * result of
        SiameseSConvOnNodes is fused into the second gather kernel."""
        x = data.x
        edge_index = data.edge_index.to(x.device).contiguous()
        edge_attr = data.edge_attr.to(x.device, torch.float32).contiguous()
        ptr, eptr = graph_offsets(data)
        # in-edge lists: built once per forward and shared by both conv layers.  The per-graph edge bound
        # comes from the host-side shape when the caller provides it (no device sync), else from eptr.
        max_e = getattr(data, "max_edges_per_graph", None)
        if max_e is None:
            max_e = int((eptr[1:] - eptr[:-1]).max()) if eptr.numel() > 1 else 0
        csr = ops.csr_by_dst(edge_index, ptr.to(x.device).contiguous(), eptr.to(x.device).contiguous(),
                             x.shape[0], int(max_e))
        plan = None
        c0 = self.convs[0]
        if slab_plan_enabled() and c0.out_channels % 128 == 0 and c0.in_channels % 8 == 0:
            plan = ops.SlabPlan(edge_index, edge_attr, x.shape[0], c0.out_channels, c0.kernel_size)
        # with a plan the hidden layer only ever exists as the fp16 operand of the second layer's slab GEMM: the first
        # layer's gather writes it in that form (no fp32 tensor, no separate split pass)
        mid = None
        if plan is not None and ops.gather_split_enabled() and c0.out_channels == self.convs[1].in_channels:
            mid = ops.slab_operand_buffers(plan, c0.out_channels, x.device)
        h = self.convs[0](x, edge_index, edge_attr, csr=csr, mode=0, plan=plan, split_out=mid)
        if residual_scale_input is not None:
            return self.convs[1](h, edge_index, edge_attr, csr=csr, mode=1, residual=residual_scale_input, plan=plan,
                                 presplit=mid)
        return self.convs[1](h, edge_index, edge_attr, csr=csr, mode=2, plan=plan, presplit=mid)


class SiameseSConvOnNodes(torch.nn.Module):
    def __init__(self, input_node_dim):
        super(SiameseSConvOnNodes, self).__init__()
        self.num_node_features = input_node_dim
        self.mp_network = SConv(input_features=self.num_node_features, output_features=self.num_node_features)

    def forward(self, graph):
        old_features = graph.x.detach().contiguous()
        graph.x = self.mp_network(graph, residual_scale_input=old_features)
        return graph


class SiameseNodeFeaturesToEdgeFeatures(torch.nn.Module):
    def __init__(self, total_num_nodes):
        super(SiameseNodeFeaturesToEdgeFeatures, self).__init__()
        self.num_edge_features = total_num_nodes

    def forward(self, graph, hyperedge=False):
        if hyperedge:
            raise NotImplementedError("hyperedge attributes are not on the matching head's path")
        orig_graphs = graph.to_data_list()
        return [self.vertex_attr_to_edge_attr(g) for g in orig_graphs]

    def vertex_attr_to_edge_attr(self, graph):
        """Assigns the difference of node features to each edge (spline_conv.py:73-81).  Net.forward does
        not call this: the affinity kernel forms x[src] - x[dst] on the fly."""
        graph.edge_attr = graph.x[graph.edge_index[0]] - graph.x[graph.edge_index[1]]
        return graph
