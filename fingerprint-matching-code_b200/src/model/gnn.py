"""Drop-in for ``PYGNNLayer`` of ``/root/reference/src/model/gnn.py:171-226`` (the live NGM-v2 layer).

Parameters keep the reference's names (``conv2.lin_l``, ``conv2.lin_r``, ``n_self_func``, ``classifier``,
and the unused ``conv`` GCN weights) so checkpoints load unchanged.

``GNNLayer`` (gnn.py:11-87, the dense NGM-v1 layer that ``construct_aff_mat`` feeds; dormant in the reference's
``Net.forward``) is mirrored too: same parameters (``e_func``, ``n_func``, ``n_self_func``, ``classifier``), dense
``A [b,N,N]`` / ``W [b,N,N,fe]`` inputs, aggregation by ``csrc/fgm_gnn.cu`` with a hand-written backward.

Two forward paths of ``PYGNNLayer``:
* ``forward_factorised`` - what ``Net.forward`` uses: the association graph is never materialised; one
  launch per layer works from the two keypoint graphs' in-neighbour lists (``csrc/gnn.cu``).
* ``forward(adj, x, n1, n2, idx)`` - the reference signature for a caller that brings an explicit sparse
  adjacency (anything exposing ``coo()`` like torch_sparse.SparseTensor, or a ``(row, col)`` pair).  It
  evaluates the same layer with the generic mean aggregation below; it is a compatibility path, not the
  hot path.
"""
import torch
import torch.nn as nn

from fpmatch import ops
from src.model.sinkhorn import Sinkhorn


class _SAGEParams(nn.Module):
    """Parameter holder with torch_geometric SAGEConv's names: lin_l (bias), lin_r (no bias)."""
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)


class _GCNParams(nn.Module):
    """The reference constructs a GCNConv it never calls (gnn.py:198); its parameters exist in
    checkpoints, so they exist here (torch_geometric 1.6.3 layout: weight [in, out], bias [out])."""
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        nn.init.xavier_uniform_(self.weight)


class PYGNNLayer(nn.Module):
    def __init__(self, in_node_features, in_edge_features, out_node_features, out_edge_features,
                 sk_channel=0, sk_iter=20, sk_tau=0.05, edge_emb=False):
        super(PYGNNLayer, self).__init__()
        self.in_nfeat = in_node_features
        self.in_efeat = in_edge_features
        self.out_efeat = out_edge_features
        self.sk_channel = sk_channel
        assert out_node_features == out_edge_features + self.sk_channel
        if self.sk_channel > 0:
            self.out_nfeat = out_node_features - self.sk_channel
            self.sk = Sinkhorn(sk_iter, sk_tau)
            self.classifier = nn.Linear(self.out_nfeat, self.sk_channel)
        else:
            self.out_nfeat = out_node_features
            self.sk = self.classifier = None
        if edge_emb:
            # gnn.py:188-196: the reference creates this edge MLP (so its parameters exist in checkpoints) and its
            # PYGNNLayer.forward never calls it (gnn.py:207-226) - same here
            self.e_func = nn.Sequential(
                nn.Linear(self.in_efeat + self.in_nfeat, self.out_efeat),
                nn.ReLU(),
                nn.Linear(self.out_efeat, self.out_efeat),
                nn.ReLU()
            )
        else:
            self.e_func = None
        self.conv = _GCNParams(self.in_nfeat, self.out_nfeat)
        self.conv2 = _SAGEParams(self.in_nfeat, self.out_nfeat)
        self.n_self_func = nn.Sequential(
            nn.Linear(self.in_nfeat, self.out_nfeat),
            nn.ReLU(),
            nn.Linear(self.out_nfeat, self.out_nfeat),
            nn.ReLU()
        )

    def kernel_weights(self):
        """The 9 tensors fpm_gnn_layer takes, in header order."""
        d = lambda t: t.detach().contiguous()
        return [d(self.conv2.lin_l.weight), d(self.conv2.lin_l.bias), d(self.conv2.lin_r.weight),
                d(self.n_self_func[0].weight), d(self.n_self_func[0].bias),
                d(self.n_self_func[2].weight), d(self.n_self_func[2].bias),
                d(self.classifier.weight).reshape(-1), d(self.classifier.bias)]

    def forward_factorised(self, xprev, mprev_t, assoc, n1, n2):
        """One NGM layer for the whole batch on the factorised association graph ``assoc``
        (``fpmatch.ops.AssocStructure``).  Returns (x1 [B,N,16], sinkhorn [B,n1max,n2max],
        sinkhorn^T [B,n2max,n1max]); the layer's 17-channel output is (x1, vec(sinkhorn^T))."""
        assert self.sk_channel == 1 and self.out_nfeat == 16
        x1, score = ops.gnn_layer(xprev, mprev_t, assoc, self.kernel_weights())
        sk, sk_t = ops.sinkhorn_log(score, n1, n2, self.sk.max_iter, self.sk.tau, True, want_t=True)
        return x1, sk, sk_t

    def forward(self, adj_sparse, x, n1=None, n2=None, idx=None):
        if hasattr(adj_sparse, "coo"):
            row, col, _ = adj_sparse.coo()
        else:
            row, col = adj_sparse
        N = x.shape[1]
        xs = x[0]
        # SAGEConv(x, adj.t()) with values dropped: mean over entries (row_t -> col_t)
        agg = torch.zeros_like(xs).index_add_(0, col, xs[row])
        cnt = torch.bincount(col, minlength=N).clamp(min=1).to(xs.dtype)
        agg = agg / cnt[:, None]
        x1 = self.conv2.lin_l(agg) + self.conv2.lin_r(xs)
        x1 = (x1 + self.n_self_func(xs)).unsqueeze(0)
        if self.classifier is not None:
            assert n1.max() * n2.max() == x.shape[1]
            x2 = self.classifier(x1)
            n1_rep = torch.repeat_interleave(n1[idx].unsqueeze(0), self.sk_channel, dim=0)
            n2_rep = torch.repeat_interleave(n2[idx].unsqueeze(0), self.sk_channel, dim=0)
            x3 = x2.permute(0, 2, 1).reshape(x.shape[0] * self.sk_channel, int(n2.max()), int(n1.max())).transpose(1, 2)
            x4 = self.sk(x3.contiguous(), n1_rep, n2_rep, dummy_row=True).transpose(2, 1).contiguous()
            x5 = x4.reshape(x.shape[0], self.sk_channel, int(n1.max() * n2.max())).permute(0, 2, 1)
            return torch.cat((x1, x5), dim=-1)
        return x1


class GNNLayer(nn.Module):
    """Dense NGM-v1 message-passing layer (``/root/reference/src/model/gnn.py:11-87``).

    ``forward(A, W, x, n1, n2, norm)``: ``A [b,N,N]`` 0/1 adjacency of the association graph, ``W [b,N,N,fe]`` edge
    tensor (``K.unsqueeze(-1)`` in NGM), ``x [b,N,fn]`` node features; returns ``(W_new, x_new)``.  The N x N x fe
    product the reference materialises for ``torch.matmul`` is never formed: one kernel evaluates
    ``x2[i] = sum_j normalize(A)[i,j] W_new[i,j] n_func(x)[j]`` row by row (``fpm_fgm_aggregate``), with its own
    backward (``fpmatch.autograd.FgmAggregateFn``).  The small per-node MLPs and the optional edge MLP are stock
    ``nn.Linear`` layers, as in the reference.
    """

    def __init__(self, in_node_features, in_edge_features, out_node_features, out_edge_features,
                 sk_channel=0, sk_iter=20, sk_tau=0.05, edge_emb=False):
        super(GNNLayer, self).__init__()
        self.in_nfeat = in_node_features
        self.in_efeat = in_edge_features
        self.out_efeat = out_edge_features
        self.sk_channel = sk_channel
        assert out_node_features == out_edge_features + self.sk_channel
        if self.sk_channel > 0:
            self.out_nfeat = out_node_features - self.sk_channel
            self.sk = Sinkhorn(sk_iter, sk_tau)
            self.classifier = nn.Linear(self.out_nfeat, self.sk_channel)
        else:
            self.out_nfeat = out_node_features
            self.sk = self.classifier = None
        if edge_emb:
            self.e_func = nn.Sequential(
                nn.Linear(self.in_efeat + self.in_nfeat, self.out_efeat), nn.ReLU(),
                nn.Linear(self.out_efeat, self.out_efeat), nn.ReLU())
        else:
            self.e_func = None
        self.n_func = nn.Sequential(
            nn.Linear(self.in_nfeat, self.out_nfeat), nn.ReLU(),
            nn.Linear(self.out_nfeat, self.out_nfeat), nn.ReLU())
        self.n_self_func = nn.Sequential(
            nn.Linear(self.in_nfeat, self.out_nfeat), nn.ReLU(),
            nn.Linear(self.out_nfeat, self.out_nfeat), nn.ReLU())

    def forward(self, A, W, x, n1=None, n2=None, norm=True):
        """
        :param A: adjacent matrix in 0/1 (b x n x n)
        :param W: edge feature tensor (b x n x n x feat_dim)
        :param x: node feature tensor (b x n x feat_dim)
        """
        from fpmatch import autograd as fa
        if self.e_func is not None:
            W1 = torch.mul(A.unsqueeze(-1), x.unsqueeze(1))
            W_new = self.e_func(torch.cat((W, W1), dim=-1))
        else:
            W_new = W
        x1 = self.n_func(x)
        fe = W_new.shape[-1]
        assert fe == 1 or fe == x1.shape[-1], (fe, x1.shape[-1])        # the reference's matmul broadcast rule
        x2 = fa.FgmAggregateFn.apply(A.to(torch.float32), W_new.to(torch.float32), x1, bool(norm))
        x2 = x2 + self.n_self_func(x)

        if self.classifier is not None:
            assert n1.max() * n2.max() == x.shape[1]
            x3 = self.classifier(x2)
            n1_rep = torch.repeat_interleave(n1, self.sk_channel, dim=0)
            n2_rep = torch.repeat_interleave(n2, self.sk_channel, dim=0)
            n1m, n2m = int(n1.max()), int(n2.max())
            x4 = x3.permute(0, 2, 1).reshape(x.shape[0] * self.sk_channel, n2m, n1m).transpose(1, 2).contiguous()
            if x4.requires_grad and torch.is_grad_enabled():
                x5 = fa.SinkhornFn.apply(x4, n1_rep.to(x4.device), n2_rep.to(x4.device), self.sk.max_iter,
                                         self.sk.tau, True)
            else:
                x5 = self.sk(x4, n1_rep, n2_rep, dummy_row=True)
            x5 = x5.transpose(2, 1).contiguous()
            x6 = x5.reshape(x.shape[0], self.sk_channel, n1m * n2m).permute(0, 2, 1)
            x_new = torch.cat((x2, x6), dim=-1)
        else:
            x_new = x2
        return W_new, x_new
