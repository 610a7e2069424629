"""torch.autograd.Function wrappers: the differentiable ops of the matching head for TRAINING
(BASELINE.json config 3: stage-1 step, ``/root/reference/train.py`` + ``src/train/training_loop.py:33-64``).

The reference lets torch autograd differentiate its python forward.  Here every stage of the head is one
Function whose forward launches the inference kernels (``fpmatch.ops``) and whose backward launches the
hand-written vector-Jacobian kernels declared under "training" in ``include/fpmatch.h``.  PyTorch only routes
tensors between them and differentiates the handful of ``[B, 1024]``-sized expressions left in python.

Nothing here has a CPU path: the ops raise for CPU tensors.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.autograd import Function

from . import ops

Tensor = torch.Tensor


class NodeFeaturesFn(Function):
    """normalize_over_channels + feature_align + concat (ngm.py:241-248): raw NCHW maps -> X [sum n, C1+C2]."""

    @staticmethod
    def forward(ctx, nodes: Tensor, edges: Tensor, P: Tensor, ns: Tensor, ptr: Tensor, total: int, ori_size):
        nodes = nodes.contiguous(); edges = edges.contiguous()
        ncl, ecl = ops.fmap_prep(nodes), ops.fmap_prep(edges)
        X = ops.node_features(ncl, ecl, nodes.shape[2:], edges.shape[2:], P, ns, ptr, total, ori_size)
        ctx.save_for_backward(nodes, edges, P, ns, ptr)
        ctx.ori_size = ori_size
        return X

    @staticmethod
    def backward(ctx, dX: Tensor):
        nodes, edges, P, ns, ptr = ctx.saved_tensors
        d1, d2 = ops.node_features_bwd(dX.contiguous(), P, ns, ptr, tuple(nodes.shape[1:]), tuple(edges.shape[1:]),
                                       ctx.ori_size)
        return ops.fmap_prep_bwd(nodes, d1), ops.fmap_prep_bwd(edges, d2), None, None, None, None, None


class GraphCtx:
    """Per-graph-batch structure shared by both SplineConv layers: edge lists grouped by destination (forward
    gather) and by source (backward scatter)."""

    def __init__(self, edge_index: Tensor, pseudo: Tensor, ptr: Tensor, eptr: Tensor, total: int, max_edges: int):
        self.edge_index = edge_index.contiguous()
        self.pseudo = pseudo.contiguous()
        self.in_csr = ops.csr_by_dst(self.edge_index, ptr, eptr, total, max_edges)
        swapped = torch.stack((self.edge_index[1], self.edge_index[0]), 0).contiguous()
        self.out_csr = ops.csr_by_dst(swapped, ptr, eptr, total, max_edges)
        self.total = total
        self._plans = {}

    def plan(self, channels: int, kernel_size: int):
        """Slab plan of this graph (built once, shared by both conv layers) or None when the dense product is used."""
        from src.model.spline_conv import slab_plan_enabled
        if not slab_plan_enabled() or channels % 128:
            return None
        key = (channels, kernel_size)
        if key not in self._plans:
            self._plans[key] = ops.SlabPlan(self.edge_index, self.pseudo, self.total, channels, kernel_size)
        return self._plans[key]

    # Below this many nodes the compacted backward loses: its two host reads drain the launch queue of a step that is
    # launch-bound anyway (measured: 8 pairs x 100 keypoints 8.1 -> 9.4 ms, 64 pairs 17.4 -> 15.6 ms).
    COMPACT_BACKWARD_MIN_NODES = 4096

    def groups(self, channels: int, kernel_size: int):
        """Wide / narrow slab groups for the compacted backward (None when the dense product is used)."""
        plan = self.plan(channels, kernel_size)
        if plan is None or self.total < self.COMPACT_BACKWARD_MIN_NODES:
            return None
        key = ("groups", channels, kernel_size)
        if key not in self._plans:
            self._plans[key] = ops.SlabGroups(plan)
        return self._plans[key]


class SplineConvFn(Function):
    """SplineConv(768,768,dim=2,kernel_size=5,aggr='max') + relu (mode 0) / xin + 0.1*out (mode 1) / nothing (2)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, root: Tensor, bias: Tensor, xin: Optional[Tensor], packed: Tensor,
                g: GraphCtx, mode: int, kernel_size: int):
        x = x.contiguous()
        plan = g.plan(weight.shape[2], kernel_size)
        Y = ops.spline_slab_gemm(x, packed, plan) if plan is not None else ops.gemm_nt(x, packed, weight_operand=True)
        out, arg = ops.spline_gather_max(Y, xin, g.edge_index, g.pseudo, g.in_csr[0], g.in_csr[1],
                                         bias.detach().contiguous(), mode, kernel_size, want_argmax=True)
        del Y
        ctx.save_for_backward(x, packed, out if mode == 0 else None, arg)
        ctx.g, ctx.mode, ctx.ks = g, mode, kernel_size
        ctx.wshape = tuple(weight.shape)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        x, packed, out, arg = ctx.saved_tensors
        g, mode = ctx.g, ctx.mode
        K, cin, cout = ctx.wshape
        gout = gout.contiguous()
        dxin = None
        if mode == 0:
            G = gout * (out > 0).to(gout.dtype)
        elif mode == 1:
            G = gout * 0.1
            dxin = gout
        else:
            G = gout
        dbias = G.sum(0)
        grp = g.groups(cout, ctx.ks)
        if grp is None:
            dY = ops.spline_scatter_bwd(G, arg, g.edge_index, g.pseudo, g.out_csr[0], g.out_csr[1], ctx.ks)
            # dX = dY W : [total, (K+1)*out] x [(K+1)*out, in]
            dx = ops.gemm_nt(dY, ops.transpose_pad(packed))
            # dW = dY^T X : both operands K-major along the node dimension
            dW = ops.gemm_nt(ops.transpose_pad(dY), ops.transpose_pad(x))          # [(K+1)*out, in]
            del dY
        else:
            # Same products with the all-zero blocks of dY left out: the wide slabs over all nodes, the narrow slabs
            # over the few nodes that read them (ops.SlabGroups).
            dYd, dYs = ops.spline_scatter_bwd_compact(G, arg, g.edge_index, g.pseudo, g.out_csr[0], g.out_csr[1],
                                                      grp, ctx.ks)
            Wv = packed.view(K + 1, cout, cin)
            xt = ops.transpose_pad(x)
            dW = torch.zeros((K + 1, cout, cin), dtype=torch.float32, device=x.device)
            wide = torch.tensor(grp.wide, device=x.device)
            Wd = Wv[wide].reshape(-1, cin)                                          # [nD*out, in]
            dx = ops.gemm_nt(dYd, ops.transpose_pad(Wd))
            dW[wide] = ops.gemm_nt(ops.transpose_pad(dYd), xt).view(len(grp.wide), cout, cin)
            del dYd
            if dYs is not None:
                narrow = torch.tensor(grp.narrow, device=x.device)
                Ws = Wv[narrow].reshape(-1, cin)
                dx.index_add_(0, grp.rows, ops.gemm_nt(dYs, ops.transpose_pad(Ws)))  # rows are unique: deterministic
                xr = ops.transpose_pad(x[grp.rows].contiguous())
                dW[narrow] = ops.gemm_nt(ops.transpose_pad(dYs), xr).view(len(grp.narrow), cout, cin)
                del dYs
            dW = dW.view((K + 1) * cout, cin)
        dW = dW.view(K + 1, cout, cin)
        dweight = dW[:K].permute(0, 2, 1)
        droot = dW[K].t()
        return dx, dweight, droot, dbias, dxin, None, None, None, None


class FgmAggregateFn(Function):
    """x2[b,i,:] = sum_j normalize(A)[i,j] W[i,j,:] x1[j,:] of the dense NGM-v1 layer (gnn.py:54-68).  Gradients reach
    the edge tensor W and x1; the 0/1 adjacency A carries none."""

    @staticmethod
    def forward(ctx, A: Tensor, W: Tensor, x1: Tensor, norm: bool):
        A = A.contiguous(); W = W.contiguous(); x1 = x1.contiguous()
        out, inv = ops.fgm_aggregate(A, W, x1, norm)
        ctx.save_for_backward(A, W, x1, inv)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        A, W, x1, inv = ctx.saved_tensors
        gout = gout.contiguous()
        dW = ops.fgm_aggregate_dw(A, inv, gout, x1, W.shape[-1]) if ctx.needs_input_grad[1] else None
        dx1 = ops.fgm_aggregate(A, W, gout, trans=True, inv=inv) if ctx.needs_input_grad[2] else None
        return None, dW, dx1, None


class AffinityFn(Function):
    """Kp = softplus((X1 (.) c) X2^T) - 0.5, padded, with its transpose (affinity_layer.py:11-19, ngm.py:317-321)."""

    @staticmethod
    def forward(ctx, X1: Tensor, X2: Tensor, coeff: Tensor, ptr1: Tensor, ptr2: Tensor, n1max: int, n2max: int):
        X1 = X1.contiguous(); X2 = X2.contiguous(); coeff = coeff.contiguous()
        Kp, Kp_t = ops.affinity_nodes(X1, X2, coeff, ptr1, ptr2, n1max, n2max)
        ctx.save_for_backward(X1, X2, coeff, ptr1, ptr2, Kp)
        return Kp, Kp_t

    @staticmethod
    def backward(ctx, dKp: Optional[Tensor], dKp_t: Optional[Tensor]):
        X1, X2, coeff, ptr1, ptr2, Kp = ctx.saved_tensors
        dK = None
        if dKp is not None:
            dK = dKp
        if dKp_t is not None:
            dK = dKp_t.transpose(1, 2) if dK is None else dK + dKp_t.transpose(1, 2)
        # softplus'(p) = sigmoid(p) = 1 - exp(-softplus(p)),  softplus(p) = Kp + 0.5
        dP = (dK * (1.0 - torch.exp(-(Kp + 0.5)))).contiguous()
        T = ops.bmm_ragged(dP, False, X2, ptr2, ptr1, X1.shape[0])             # dP X2
        dcoeff = ops.segment_rowdot(X1, T, ptr1)
        n1 = ptr1[1:] - ptr1[:-1]
        bidx = torch.repeat_interleave(torch.arange(n1.numel(), device=X1.device), n1, output_size=X1.shape[0])
        dX1 = T * coeff[bidx]
        dX2 = ops.bmm_ragged(dP, True, X1, ptr1, ptr2, X2.shape[0], coeff_in=coeff)   # dP^T (c (.) X1)
        return dX1, dX2, dcoeff, None, None, None, None


class SinkhornFn(Function):
    @staticmethod
    def forward(ctx, s: Tensor, n1: Tensor, n2: Tensor, max_iter: int, tau: float, dummy_row: bool):
        s = s.contiguous()
        out = ops.sinkhorn_log(s, n1, n2, max_iter, tau, dummy_row)
        ctx.save_for_backward(s, n1, n2)
        ctx.cfg = (max_iter, tau, dummy_row)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        s, n1, n2 = ctx.saved_tensors
        return ops.sinkhorn_log_bwd(s, n1, n2, gout.contiguous(), *ctx.cfg), None, None, None, None, None


class SoftTopkFn(Function):
    @staticmethod
    def forward(ctx, ss: Tensor, ks: Tensor, n1: Tensor, n2: Tensor, max_iter: int, tau: float):
        ss = ss.contiguous()
        out = ops.soft_topk(ss, ks, n1, n2, max_iter, tau)
        ctx.save_for_backward(ss, ks, n1, n2)
        ctx.cfg = (max_iter, tau)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        ss, ks, n1, n2 = ctx.saved_tensors
        return ops.soft_topk_bwd(ss, ks, n1, n2, gout.contiguous(), *ctx.cfg), None, None, None, None, None


class NgmSolverFn(Function):
    """The three PYGNNLayers (factorised SAGE aggregation + linears + classifier + Sinkhorn) and the final
    classifier (ngm.py:326-369): Kp^T -> s.  ``params`` = 9 tensors per layer in ``PYGNNLayer.kernel_weights``
    order, then classifier.weight, classifier.bias."""

    @staticmethod
    def forward(ctx, Kp_t: Tensor, meta: dict, *params: Tensor):
        st, n1, n2 = meta["assoc"], meta["n1"], meta["n2"]
        n1max, n2max = meta["n1max"], meta["n2max"]
        nl = meta["layers"]
        det = lambda t: t.detach().contiguous()
        lw = [[det(params[9 * i + j]).reshape(-1) if j == 7 else det(params[9 * i + j]) for j in range(9)]
              for i in range(nl)]
        cw, cb = det(params[9 * nl]).reshape(-1), det(params[9 * nl + 1])
        xprev, m_t = None, Kp_t.contiguous()
        saved = []
        for i in range(nl):
            x1, score = ops.gnn_layer(xprev, m_t, st, lw[i])
            _, sk_t = ops.sinkhorn_log(score, n1, n2, meta["sk_iter"], meta["sk_tau"], True, want_t=True)
            saved.append((xprev, m_t, score))
            xprev, m_t = x1, sk_t
        s = ops.final_classifier(xprev, m_t, cw, cb, n1max, n2max)
        ctx.meta, ctx.lw, ctx.cw = meta, lw, cw
        ctx.saved = saved + [(xprev, m_t, None)]
        ctx.pshapes = [tuple(p.shape) for p in params]
        return s

    @staticmethod
    def backward(ctx, ds: Tensor):
        meta = ctx.meta
        st, n1, n2 = meta["assoc"], meta["n1"], meta["n2"]
        n1max, n2max = meta["n1max"], meta["n2max"]
        nl = meta["layers"]
        B = ds.shape[0]
        N = n1max * n2max
        x_last, skt_last, _ = ctx.saved[nl]
        ds = ds.contiguous()
        ds_t = ds.transpose(1, 2).reshape(B, N)                      # association-node order p = i2*n1max + i1
        cw = ctx.cw
        dcw = torch.empty_like(cw)
        dcw[:16] = torch.einsum("bp,bpc->c", ds_t, x_last)
        dcw[16] = (ds_t * skt_last.reshape(B, N)).sum()
        dcb = ds.sum().reshape(1)
        dx1 = (ds_t.unsqueeze(-1) * cw[:16]).contiguous()
        dsk = (ds * cw[16]).contiguous()
        grads: List[Optional[Tensor]] = [None] * len(ctx.pshapes)
        for i in range(nl - 1, -1, -1):
            xprev, m_t, score = ctx.saved[i]
            dscore = ops.sinkhorn_log_bwd(score, n1, n2, dsk, meta["sk_iter"], meta["sk_tau"], True)
            dxprev, dm, wg = ops.gnn_layer_bwd(xprev, m_t, st, ctx.lw[i], dx1, dscore)
            for j in range(9):
                grads[9 * i + j] = wg[j].reshape(ctx.pshapes[9 * i + j])
            dx1, dsk = dxprev, dm
        grads[9 * nl] = dcw.reshape(ctx.pshapes[9 * nl])
        grads[9 * nl + 1] = dcb.reshape(ctx.pshapes[9 * nl + 1])
        dKp_t = dsk.transpose(1, 2).contiguous()
        ctx.saved = None
        return (dKp_t, None, *grads)


# ---------------------------------------------------------------------------------------------------------------------
# AFA-U k-branch (afau.py:22-300, ngm.py:386-412): trained in stages 2-5 of train.py
# ---------------------------------------------------------------------------------------------------------------------
class LinearFn(Function):
    """act(x W^T + b) on the tensor-core GEMM; x [..., K], W [N, K].  act: 0 none, 1 relu.

    Runs in the 3xTF32 mode: the fp16 split keeps 22 bits of each ROW's maximum, the tf32 split 21 bits of each
    ELEMENT.  The k-branch feeds these outputs to an InstanceNorm over nearly identical rows (the row embedding is
    all zeros), which amplifies absolute errors of small elements by ~1e5; with the fp16 split the FFN gradients
    were 7x further from an fp64 evaluation than the fp32 oracle is, with tf32 they are on par.  The GEMMs are
    small (K = 256 / 600), so the halved MMA rate is invisible."""
    MODE = "3xtf32"

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], act: int):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        w = weight.detach().contiguous()
        out = ops.gemm_nt(x2, w, None if bias is None else bias.detach().contiguous(), act, mode=LinearFn.MODE)
        ctx.save_for_backward(x2, w, out if act == 1 else None)
        ctx.meta = (shape, act, bias is not None)
        return out.view(*shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g: Tensor):
        x2, w, out = ctx.saved_tensors
        shape, act, has_bias = ctx.meta
        g2 = g.reshape(-1, w.shape[0]).contiguous()
        if act == 1:
            g2 = g2 * (out > 0).to(g2.dtype)
        dx = ops.gemm_nt(g2, ops.transpose_pad(w), mode=LinearFn.MODE)                         # g2 [M, N] . W [N, K]
        dw = ops.gemm_nt(ops.transpose_pad(g2), ops.transpose_pad(x2), mode=LinearFn.MODE)     # g2^T . x
        return dx.view(shape), dw, (g2.sum(0) if has_bias else None), None


class OnehotProjFn(Function):
    """Projection of the one-hot column embedding: out[b, j, :] = W[:, j] for j < n[b], else 0 (ngm.py:396-399)."""

    @staticmethod
    def forward(ctx, W: Tensor, n: Tensor, nmax: int):
        ctx.save_for_backward(n)
        ctx.meta = (tuple(W.shape), nmax)
        return ops.onehot_proj(W.detach().contiguous(), n, nmax)

    @staticmethod
    def backward(ctx, g: Tensor):
        (n,) = ctx.saved_tensors
        (OUT, IN), nmax = ctx.meta
        mask = (torch.arange(nmax, device=g.device)[None, :] < n[:, None]).to(g.dtype)       # [B, nmax]
        dW = torch.zeros((OUT, IN), dtype=g.dtype, device=g.device)
        m = min(nmax, IN)
        dW[:, :m] = (g * mask[:, :, None]).sum(0).t()[:, :m]
        return dW, None, None


class AfauAttentionFn(Function):
    """CrossSet_MultiHeadAttention.forward (afau.py:231-300) with the heads concatenated: q [B,nr,256], k/v [B,nc,256]."""

    @staticmethod
    def forward(ctx, q, k, v, cost, transposed_cost, mix1_w, mix1_b, mix2_w, mix2_b):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        cost = cost.detach().contiguous()
        d = lambda t: t.detach().contiguous()
        out = ops.afau_attention(q, k, v, cost, transposed_cost, d(mix1_w), d(mix1_b), d(mix2_w), d(mix2_b))
        ctx.save_for_backward(q, k, v, cost, d(mix1_w), d(mix1_b), d(mix2_w), d(mix2_b), out)
        ctx.transposed_cost = transposed_cost
        return out

    @staticmethod
    def backward(ctx, g):
        q, k, v, cost, m1w, m1b, m2w, m2b, out = ctx.saved_tensors
        dq, dk, dv, dm1w, dm1b, dm2w, dm2b = ops.afau_attention_bwd(q, k, v, cost, ctx.transposed_cost, m1w, m1b, m2w,
                                                                    m2b, out, g.contiguous())
        return dq, dk, dv, None, None, dm1w, dm1b, dm2w, dm2b


class AddInstNormFn(Function):
    """AddAndInstanceNormalization.forward (afau.py:154-176): InstanceNorm1d(affine) over the rows of a + other;
    ``want_rowmax`` additionally returns the per-channel maximum over rows (the padded MaxPool1d of ngm.py:402-405)."""

    @staticmethod
    def forward(ctx, a, other, gamma, beta, eps, want_rowmax):
        a = a.contiguous()
        other_c = None if other is None else other.detach().contiguous()
        g, b = gamma.detach().contiguous(), beta.detach().contiguous()
        res = ops.add_instnorm(a, other_c, g, b, want_rowmax=want_rowmax, eps=eps)
        ctx.save_for_backward(a, other_c, g)
        ctx.meta = (eps, want_rowmax, other is not None and other.dim() == 1)
        if want_rowmax:
            return res[0], res[1]
        return res

    @staticmethod
    def backward(ctx, dy, drowmax=None):
        a, other, g = ctx.saved_tensors
        eps, want_rowmax, vec = ctx.meta
        dy_c = None if dy is None else dy.contiguous()
        dr_c = None if (drowmax is None or not want_rowmax) else drowmax.contiguous()
        if dy_c is None and dr_c is None:
            return None, None, None, None, None, None
        dx, dgamma, dbeta, dvec = ops.add_instnorm_bwd(a, other, g, dy_c, dr_c, eps)
        dother = None if other is None else (dvec if vec else dx)
        return dx, dother, dgamma, dbeta, None, None
