"""Op-level CPU restatements (torch CPU, fp32 by default).  Test infrastructure - see package docstring.

Every function follows the reference's arithmetic order where the reference owns the arithmetic
and the published algorithm where a third-party package owns it.  ``dtype`` is threaded through so
the same code can run in fp64 as a "truth" when a test needs to bound fp32 reordering noise.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# A1  feature_align  (/root/reference/utils/feature_align.py:5-125)
# --------------------------------------------------------------------------------------------
def bilinear_interpolate(im: Tensor, x: Tensor, y: Tensor) -> Tensor:
    """One point; follows feature_align.py:67-125 statement by statement."""
    x = x.to(torch.float32)
    y = y.to(torch.float32)
    x0 = torch.floor(x); x1 = x0 + 1
    y0 = torch.floor(y); y1 = y0 + 1
    x0 = torch.clamp(x0, 0, im.shape[2] - 1); x1 = torch.clamp(x1, 0, im.shape[2] - 1)
    y0 = torch.clamp(y0, 0, im.shape[1] - 1); y1 = torch.clamp(y1, 0, im.shape[1] - 1)
    x0 = x0.to(torch.int32); x1 = x1.to(torch.int32)
    y0 = y0.to(torch.int32); y1 = y1.to(torch.int32)
    Ia = im[:, y0, x0]; Ib = im[:, y1, x0]; Ic = im[:, y0, x1]; Id = im[:, y1, x1]
    if x0 == x1:                       # taps are fetched BEFORE this adjustment (:98-113)
        if x0 == 0:
            x0 = x0 - 1
        else:
            x1 = x1 + 1
    if y0 == y1:
        if y0 == 0:
            y0 = y0 - 1
        else:
            y1 = y1 + 1
    x0 = x0.to(torch.float32); x1 = x1.to(torch.float32)
    y0 = y0.to(torch.float32); y1 = y1.to(torch.float32)
    wa = (x1 - x) * (y1 - y); wb = (x1 - x) * (y - y0)
    wc = (x - x0) * (y1 - y); wd = (x - x0) * (y - y0)
    return Ia * wa + Ib * wb + Ic * wc + Id * wd


def feature_align_loop(raw_feature: Tensor, P: Tensor, ns: Tensor, ori_size) -> Tensor:
    """Reference loop structure (per image, per point): feature_align.py:24-64."""
    B, C, n_max = raw_feature.shape[0], raw_feature.shape[1], P.shape[1]
    ori = torch.tensor(ori_size, dtype=torch.float32)
    out = torch.zeros(B, C, n_max, dtype=torch.float32)
    for b in range(B):
        feat = raw_feature[b]
        feat_size = torch.as_tensor(feat.shape[1:3], dtype=torch.float32)   # (Hf, Wf): the W/H quirk
        step = ori / feat_size
        for i in range(int(ns[b])):
            p = (P[b, i] - step / 2) / ori * feat_size
            out[b, :, i] = bilinear_interpolate(feat, p[0], p[1])
    return out


def feature_align(raw_feature: Tensor, P: Tensor, ns: Tensor, ori_size) -> Tensor:
    """Vectorised twin of ``feature_align_loop``: same fp32 op sequence per element, so bit-equal."""
    B, C, Hf, Wf = raw_feature.shape
    n_max = P.shape[1]
    ori = torch.tensor(ori_size, dtype=torch.float32)
    feat_size = torch.tensor([Hf, Wf], dtype=torch.float32)
    step = ori / feat_size
    p = (P.to(torch.float32) - step / 2) / ori * feat_size            # [B, n, 2]
    x, y = p[..., 0], p[..., 1]
    x0 = torch.floor(x); x1 = x0 + 1
    y0 = torch.floor(y); y1 = y0 + 1
    x0 = torch.clamp(x0, 0, Wf - 1); x1 = torch.clamp(x1, 0, Wf - 1)
    y0 = torch.clamp(y0, 0, Hf - 1); y1 = torch.clamp(y1, 0, Hf - 1)
    xi0, xi1, yi0, yi1 = x0.long(), x1.long(), y0.long(), y1.long()
    flat = raw_feature.reshape(B, C, Hf * Wf)

    def tap(yi, xi):
        idx = (yi * Wf + xi)[:, None, :].expand(B, C, n_max)
        return torch.gather(flat, 2, idx)

    Ia, Ib, Ic, Id = tap(yi0, xi0), tap(yi1, xi0), tap(yi0, xi1), tap(yi1, xi1)
    eqx, eqy = xi0 == xi1, yi0 == yi1
    x0 = torch.where(eqx & (xi0 == 0), x0 - 1, x0); x1 = torch.where(eqx & (xi0 != 0), x1 + 1, x1)
    y0 = torch.where(eqy & (yi0 == 0), y0 - 1, y0); y1 = torch.where(eqy & (yi0 != 0), y1 + 1, y1)
    wa = ((x1 - x) * (y1 - y))[:, None, :]; wb = ((x1 - x) * (y - y0))[:, None, :]
    wc = ((x - x0) * (y1 - y))[:, None, :]; wd = ((x - x0) * (y - y0))[:, None, :]
    out = Ia * wa + Ib * wb + Ic * wc + Id * wd
    valid = (torch.arange(n_max)[None, :] < ns.view(-1, 1))[:, None, :]
    return torch.where(valid, out, torch.zeros((), dtype=out.dtype))


def normalize_over_channels(x: Tensor) -> Tensor:
    """ngm.py:65-67."""
    return x / torch.norm(x, dim=1, keepdim=True)


def concat_features(embeddings: Tensor, num_vertices: Tensor) -> Tensor:
    """ngm.py:70-72: [B, C, nmax] -> [sum n, C]."""
    res = torch.cat([e[:, :int(nv)] for e, nv in zip(embeddings, num_vertices)], dim=-1)
    return res.transpose(0, 1)


# --------------------------------------------------------------------------------------------
# A2  SplineConv  (torch_geometric 1.6.3 SplineConv + torch_spline_conv 1.2.0)
# --------------------------------------------------------------------------------------------
def spline_basis(pseudo: Tensor, kernel_size: int = 5, degree: int = 1) -> Tuple[Tensor, Tensor]:
    """Open B-spline basis, degree 1, dim 2 -> (basis [E,4] fp32, weight_index [E,4] int64).

    torch_spline_conv cpu/basis_cpu.cpp: the per-dimension factor is evaluated in double and cast
    back to float, the product over dimensions runs in float.
    """
    assert degree == 1 and pseudo.shape[1] == 2
    E = pseudo.shape[0]
    pseudo = pseudo.to(torch.float32)
    basis = torch.empty(E, 4, dtype=torch.float32)
    wi = torch.empty(E, 4, dtype=torch.long)
    v = pseudo * float(kernel_size - degree)          # is_open_spline = 1
    fl = torch.floor(v)
    frac = v - fl
    base = fl.long()
    for s in range(4):
        k = s
        b = torch.ones(E, dtype=torch.float32)
        w = torch.zeros(E, dtype=torch.long)
        off = 1
        for d in range(2):
            k_mod = k % 2
            k //= 2
            w = w + ((base[:, d] + k_mod) % kernel_size) * off
            off *= kernel_size
            f = frac[:, d]
            fac = (1.0 - f.double() - k_mod + 2.0 * f.double() * k_mod).to(torch.float32)
            b = b * fac
        basis[:, s] = b
        wi[:, s] = w
    return basis, wi


def spline_conv(x: Tensor, edge_index: Tensor, pseudo: Tensor, weight: Tensor, root: Tensor,
                bias: Tensor) -> Tensor:
    """SplineConv(in, out, dim=2, kernel_size=5, aggr='max', root_weight, bias).

    message m_e = sum_s basis[e,s] * (x[src_e] @ weight[wi[e,s]]);  out_i = max_{e: dst_e = i} m_e
    (0 for nodes without in-edges, torch_scatter's fill), then + x @ root + bias.
    """
    dt = x.dtype
    src, dst = edge_index[0], edge_index[1]
    basis, wi = spline_basis(pseudo)
    basis = basis.to(dt)
    xj = x[src]
    msg = torch.zeros(xj.shape[0], weight.shape[2], dtype=dt)
    K = weight.shape[0]
    for s in range(4):
        for k in range(K):
            sel = torch.nonzero(wi[:, s] == k).flatten()
            if sel.numel():
                msg[sel] += (basis[sel, s:s + 1] * xj[sel]) @ weight[k]
    out = torch.zeros(x.shape[0], weight.shape[2], dtype=dt)
    if msg.shape[0]:
        idx = dst[:, None].expand_as(msg)
        out = out.scatter_reduce(0, idx, msg, reduce="amax", include_self=False)
    out = out + x @ root
    out = out + bias
    return out


def sconv_residual(x: Tensor, edge_index: Tensor, pseudo: Tensor, p: dict, prefix: str) -> Tensor:
    """SiameseSConvOnNodes: x + 0.1 * conv1(relu(conv0(x)))   (spline_conv.py:28-41,51-58)."""
    def rootw(i):
        k = f"{prefix}.mp_network.convs.{i}.root"
        if k in p:
            return p[k]
        return p[f"{prefix}.mp_network.convs.{i}.lin.weight"].t()      # PyG 2.x name
    h = spline_conv(x, edge_index, pseudo, p[f"{prefix}.mp_network.convs.0.weight"], rootw(0),
                    p[f"{prefix}.mp_network.convs.0.bias"])
    h = F.relu(h)
    h = spline_conv(h, edge_index, pseudo, p[f"{prefix}.mp_network.convs.1.weight"], rootw(1),
                    p[f"{prefix}.mp_network.convs.1.bias"])
    return x + 0.1 * h


# --------------------------------------------------------------------------------------------
# A4/A5  affinity  (/root/reference/src/model/affinity_layer.py:11-22)
# --------------------------------------------------------------------------------------------
def affinity(X: Tensor, Y: Tensor, w: Tensor, A_weight: Tensor, A_bias: Tensor) -> Tensor:
    coeff = torch.tanh(F.linear(w, A_weight, A_bias))
    res = torch.matmul(X * coeff, Y.transpose(0, 1))
    return F.softplus(res) - 0.5


# --------------------------------------------------------------------------------------------
# A8  pygmtools.sinkhorn, pytorch backend, 0.5.3  (PARITY UNPINNED)
# --------------------------------------------------------------------------------------------
def sinkhorn(s: Tensor, nrows: Optional[Tensor] = None, ncols: Optional[Tensor] = None,
             dummy_row: bool = False, max_iter: int = 10, tau: float = 1.0) -> Tensor:
    """batched_operation=False path: per-sample crop, alternate row/column log-normalisation."""
    B = s.shape[0]
    if s.shape[2] >= s.shape[1]:
        transposed = False
    else:
        s = s.transpose(1, 2)
        nrows, ncols = ncols, nrows
        transposed = True
    if nrows is None:
        nrows = torch.full((B,), s.shape[1], dtype=torch.long)
    if ncols is None:
        ncols = torch.full((B,), s.shape[2], dtype=torch.long)
    nrows = nrows.long(); ncols = ncols.long()

    transposed_batch = nrows > ncols
    if torch.any(transposed_batch):
        s_t = s.transpose(1, 2)
        s_t = torch.cat((s_t[:, :s.shape[1], :],
                         torch.full((B, s.shape[1], s.shape[2] - s.shape[1]), -float("inf"),
                                    dtype=s.dtype)), dim=2)
        s = torch.where(transposed_batch.view(B, 1, 1), s_t, s)
        nrows, ncols = (torch.where(transposed_batch, ncols, nrows),
                        torch.where(transposed_batch, nrows, ncols))

    log_s = s / tau
    if dummy_row:
        assert log_s.shape[2] >= log_s.shape[1]
        dummy_shape = list(log_s.shape)
        dummy_shape[1] = log_s.shape[2] - log_s.shape[1]
        ori_nrows = nrows
        nrows = ncols.clone()
        log_s = torch.cat((log_s, torch.full(dummy_shape, -float("inf"), dtype=log_s.dtype)), dim=1)
        for b in range(B):
            log_s[b, int(ori_nrows[b]):int(nrows[b]), :int(ncols[b])] = -100

    ret = torch.full(tuple(log_s.shape), -float("inf"), dtype=log_s.dtype)
    for b in range(B):
        r, c = int(nrows[b]), int(ncols[b])
        lb = log_s[b, :r, :c]
        for i in range(max_iter):
            if i % 2 == 0:
                lb = lb - torch.logsumexp(lb, 1, keepdim=True)
            else:
                lb = lb - torch.logsumexp(lb, 0, keepdim=True)
        ret[b, :r, :c] = lb

    if dummy_row:
        if dummy_shape[1] > 0:
            ret = ret[:, :-dummy_shape[1]]
        for b in range(B):
            ret[b, int(ori_nrows[b]):int(nrows[b]), :int(ncols[b])] = -float("inf")

    if torch.any(transposed_batch):
        s_t = ret.transpose(1, 2)
        s_t = torch.cat((s_t[:, :ret.shape[1], :],
                         torch.full((B, ret.shape[1], ret.shape[2] - ret.shape[1]), -float("inf"),
                                    dtype=ret.dtype)), dim=2)
        ret = torch.where(transposed_batch.view(B, 1, 1), s_t, ret)
    if transposed:
        ret = ret.transpose(1, 2)
    return torch.exp(ret)


# --------------------------------------------------------------------------------------------
# A6/A7  sparse association graph + PYGNNLayer  (ngm.py:317-348, gnn.py:171-226)
# --------------------------------------------------------------------------------------------
def sage_mean_aggregate(x: Tensor, row: Tensor, col: Tensor, N: int) -> Tensor:
    """``matmul(adj.t(), x, reduce='mean')`` of torch_sparse with values dropped.

    adj holds entries (row_t, col_t); adj.t() row c lists x[row_t] for col_t == c in ascending
    row_t order (SparseTensor sorts by row-major key), summed sequentially in fp32, divided by the
    entry count (0 entries -> 0).
    """
    key = col * N + row
    order = torch.argsort(key, stable=True)
    r, c = row[order], col[order]
    out = torch.zeros(N, x.shape[1], dtype=x.dtype)
    out.index_add_(0, c, x[r])          # index_add on CPU accumulates in index order
    cnt = torch.bincount(c, minlength=N).to(x.dtype).clamp(min=1)
    return out / cnt[:, None]


def pygnn_layer(x: Tensor, row: Tensor, col: Tensor, n1: int, n2: int, n1max: int, n2max: int,
                p: dict, prefix: str, sk_iter: int = 20, sk_tau: float = 0.01) -> Tensor:
    """x [N, Cin] -> [N, 17]; N = n1max*n2max, association index p = i2*n1max + i1."""
    N = n1max * n2max
    agg = sage_mean_aggregate(x, row, col, N)
    x1 = F.linear(agg, p[f"{prefix}.conv2.lin_l.weight"], p[f"{prefix}.conv2.lin_l.bias"]) \
        + F.linear(x, p[f"{prefix}.conv2.lin_r.weight"])
    h = F.relu(F.linear(x, p[f"{prefix}.n_self_func.0.weight"], p[f"{prefix}.n_self_func.0.bias"]))
    h = F.relu(F.linear(h, p[f"{prefix}.n_self_func.2.weight"], p[f"{prefix}.n_self_func.2.bias"]))
    x1 = x1 + h
    x2 = F.linear(x1, p[f"{prefix}.classifier.weight"], p[f"{prefix}.classifier.bias"])   # [N,1]
    x3 = x2.reshape(1, n2max, n1max).transpose(1, 2)                                    # [1,n1,n2]
    x4 = sinkhorn(x3, torch.tensor([n1]), torch.tensor([n2]), dummy_row=True,
                  max_iter=sk_iter, tau=sk_tau)
    x5 = x4.transpose(2, 1).contiguous().reshape(N, 1)
    return torch.cat((x1, x5), dim=-1)


# --------------------------------------------------------------------------------------------
# A9  AFA-U encoder  (/root/reference/src/model/afau.py:22-300)
# --------------------------------------------------------------------------------------------
def _instance_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """InstanceNorm1d(affine) over the node dim of x [B, n, E]  (afau.py:152,165-176)."""
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def afau_block(row_emb: Tensor, col_emb: Tensor, cost: Tensor, p: dict, prefix: str,
               head_num: int = 16, qkv_dim: int = 16) -> Tensor:
    """EncodingBlock.forward (afau.py:109-141) + CrossSet_MultiHeadAttention (afau.py:231-300)."""
    B, nr, _ = row_emb.shape
    nc = col_emb.shape[1]

    def heads(t):
        return t.reshape(B, -1, head_num, qkv_dim).transpose(1, 2)

    q = heads(F.linear(row_emb, p[f"{prefix}.Wq.weight"]))
    k = heads(F.linear(col_emb, p[f"{prefix}.Wk.weight"]))
    v = heads(F.linear(col_emb, p[f"{prefix}.Wv.weight"]))
    dot = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(qkv_dim)               # [B,H,nr,nc]
    w1 = p[f"{prefix}.mixed_score_MHA.mix1_weight"]                             # [H,2,16]
    b1 = p[f"{prefix}.mixed_score_MHA.mix1_bias"]                               # [H,16]
    w2 = p[f"{prefix}.mixed_score_MHA.mix2_weight"]                             # [H,16,1]
    b2 = p[f"{prefix}.mixed_score_MHA.mix2_bias"]                               # [H,1]
    out_heads = []
    for h in range(head_num):             # per head to bound memory; same arithmetic as the 5-D form
        two = torch.stack((dot[:, h], cost), dim=3)                             # [B,nr,nc,2]
        ms1 = F.relu(torch.matmul(two, w1[h]) + b1[h])                          # [B,nr,nc,16]
        ms2 = torch.matmul(ms1, w2[h]).squeeze(-1) + b2[h]                      # [B,nr,nc]
        wgt = torch.softmax(ms2, dim=2)
        out_heads.append(torch.matmul(wgt, v[:, h]))                            # [B,nr,16]
    out_concat = torch.stack(out_heads, dim=2).reshape(B, nr, head_num * qkv_dim)
    mh = F.linear(out_concat, p[f"{prefix}.multi_head_combine.weight"],
                  p[f"{prefix}.multi_head_combine.bias"])
    out1 = _instance_norm(row_emb + mh, p[f"{prefix}.add_n_normalization_1.norm.weight"],
                          p[f"{prefix}.add_n_normalization_1.norm.bias"])
    ff = F.linear(F.relu(F.linear(out1, p[f"{prefix}.feed_forward.W1.weight"],
                                  p[f"{prefix}.feed_forward.W1.bias"])),
                  p[f"{prefix}.feed_forward.W2.weight"], p[f"{prefix}.feed_forward.W2.bias"])
    return _instance_norm(out1 + ff, p[f"{prefix}.add_n_normalization_2.norm.weight"],
                          p[f"{prefix}.add_n_normalization_2.norm.bias"])


def afau_encoder(row_emb: Tensor, col_emb: Tensor, cost: Tensor, p: dict,
                 prefix: str = "encoder_k") -> Tuple[Tensor, Tensor]:
    """Encoder.forward with its single EncoderLayer (afau.py:41-85): both blocks read the INPUT
    embeddings; the column block gets cost^T."""
    r = afau_block(row_emb, col_emb, cost, p, f"{prefix}.layers.0.row_encoding_block")
    c = afau_block(col_emb, row_emb, cost.transpose(1, 2), p, f"{prefix}.layers.0.col_encoding_block")
    return r, c


# --------------------------------------------------------------------------------------------
# A10  soft_topk + Sinkhorn_m  (/root/reference/src/model/soft_topk.py:8-53,166-255)
# --------------------------------------------------------------------------------------------
def soft_topk_prob(scores: Tensor, ks: Tensor, max_iter: int, tau: float, nrows: Tensor,
                   ncols: Tensor) -> Tensor:
    """Returns ``output_s`` (= ds_mat).  The discarded greedy pass of the reference is not run here;
    ``soft_topk_full`` below runs it for the CPU-baseline timing."""
    B = scores.shape[0]
    out = torch.zeros_like(scores)
    col_prob = torch.zeros((B, 2), dtype=torch.float32)
    col_prob[:, 1] += ks
    col_prob[:, 0] += nrows * ncols - ks
    log_col_prob = torch.log(col_prob).to(scores.dtype)
    for b in range(B):
        n1, n2 = int(nrows[b]), int(ncols[b])
        S = scores[b, :n1, :n2].detach()
        anchors = torch.tensor([S.min(), S.max()], dtype=scores.dtype)
        dist = -torch.abs(scores[b, :n1, :n2].reshape(-1).unsqueeze(-1) - anchors.unsqueeze(0))
        log_s = dist / tau
        lc = log_col_prob[b].unsqueeze(0)

        def step(ls, i):
            if i % 2 == 0:
                ls = ls - torch.logsumexp(ls, 1, keepdim=True) + 0.0     # log row marginal = log 1
            else:
                ls = ls - torch.logsumexp(ls, 0, keepdim=True) + lc
            return torch.where(torch.isnan(ls), torch.full_like(ls, -float("inf")), ls)

        for i in range(max_iter):
            log_s = step(log_s, i)
        it = max_iter
        while torch.any(log_s > 0):
            log_s = step(log_s, it)
            it += 1
        out[b, :n1, :n2] = torch.exp(log_s[:, 1]).view(n1, n2)
    return out


def greedy_perm(x: Tensor, top_indices: Tensor, ks: Tensor) -> Tensor:
    """soft_topk.py:56-77 (python loop kept: this is what the reference executes)."""
    for b in range(x.shape[0]):
        matched = 0
        cur = 0
        want = round(ks[b].item())
        while matched < want and cur < top_indices.shape[1]:
            idx = int(top_indices[b][cur])
            r, c = idx // x.shape[2], idx % x.shape[2]
            if x[b, :, c].sum() < 1 and x[b, r, :].sum() < 1:
                x[b, r, c] = 1
                matched += 1
            cur += 1
    return x


def greedy_topk_fast(assign: Tensor, ds: Tensor, ks: Tensor) -> Tensor:
    """Closed form of ``greedy_perm(zeros, argsort(assign*ds, desc, stable), ks)``:
    positive assigned entries by (value desc, flat index asc), then a raster-order greedy over the
    zero-valued cells of the PADDED matrix.  Used to check the loop version at sizes where the
    python loop is too slow; tests assert both agree."""
    B, R, C = assign.shape
    out = torch.zeros_like(assign)
    val = (assign * ds).reshape(B, -1)
    for b in range(B):
        want = round(ks[b].item())
        v = val[b]
        pos = torch.nonzero(v > 0).flatten()
        order = pos[torch.argsort(v[pos], descending=True, stable=True)]
        rows_used = np.zeros(R, bool); cols_used = np.zeros(C, bool)
        matched = 0
        for idx in order.tolist():
            if matched >= want:
                break
            r, c = idx // C, idx % C
            if not rows_used[r] and not cols_used[c]:
                rows_used[r] = cols_used[c] = True
                out[b, r, c] = 1; matched += 1
        if matched < want:
            # walk cells in raster order, skipping cells whose value is > 0 (already visited)
            vb = v.reshape(R, C)
            for r in range(R):
                if matched >= want:
                    break
                if rows_used[r]:
                    continue
                for c in range(C):
                    if not cols_used[c] and not (vb[r, c] > 0):
                        rows_used[r] = cols_used[c] = True
                        out[b, r, c] = 1; matched += 1
                        break
    return out


# --------------------------------------------------------------------------------------------
# A11  hungarian  (/root/reference/utils/hungarian.py:8-65)
# --------------------------------------------------------------------------------------------
def hungarian(s: Tensor, n1: Optional[Tensor] = None, n2: Optional[Tensor] = None) -> Tensor:
    import scipy.optimize as opt
    if s.dim() == 2:
        s = s.unsqueeze(0); squeeze = True
    elif s.dim() == 3:
        squeeze = False
    else:
        raise ValueError("input data shape not understood: {}".format(s.shape))
    cost = s.detach().cpu().numpy() * -1
    B = cost.shape[0]
    out = np.zeros_like(cost)
    for b in range(B):
        r = cost.shape[1] if n1 is None else int(n1[b])
        c = cost.shape[2] if n2 is None else int(n2[b])
        row, col = opt.linear_sum_assignment(cost[b, :r, :c])
        out[b, row, col] = 1
    res = torch.from_numpy(out)
    return res.squeeze(0) if squeeze else res


# --------------------------------------------------------------------------------------------
# A13  MatchClassifier  (ngm.py:75-106) - stock torch in both implementations
# --------------------------------------------------------------------------------------------
def match_classifier(m: Tensor, p: dict, training: bool = False, prefix: str = "match_cls") -> Tensor:
    x = m.unsqueeze(1)
    for conv_i, bn_i in ((0, 2), (4, 6)):
        x = F.conv2d(x, p[f"{prefix}.conv.{conv_i}.weight"], p[f"{prefix}.conv.{conv_i}.bias"], padding=1)
        x = F.relu(x)
        x = F.batch_norm(x, p[f"{prefix}.conv.{bn_i}.running_mean"], p[f"{prefix}.conv.{bn_i}.running_var"],
                         p[f"{prefix}.conv.{bn_i}.weight"], p[f"{prefix}.conv.{bn_i}.bias"],
                         training=training, eps=1e-5)
        x = F.max_pool2d(x, 2)
    x = F.adaptive_avg_pool2d(x, 1).view(x.size(0), -1)
    return F.linear(x, p[f"{prefix}.fc.weight"], p[f"{prefix}.fc.bias"]).squeeze(-1)


def gnn_layer_dense(p: dict, A: Tensor, W: Tensor, x: Tensor, n1: Tensor, n2: Tensor, norm: bool = True,
                    sk_iter: int = 20, sk_tau: float = 0.05) -> tuple:
    """GNNLayer.forward of the dense NGM-v1 path (/root/reference/src/model/gnn.py:54-87), with the reference's own
    permute + matmul formulation of the aggregation.  ``p``: the layer's state_dict (e_func.* optional)."""
    lin = torch.nn.functional.linear
    mlp = lambda name, t: torch.relu(lin(torch.relu(lin(t, p[name + ".0.weight"], p[name + ".0.bias"])),
                                         p[name + ".2.weight"], p[name + ".2.bias"]))
    if "e_func.0.weight" in p:
        W1 = torch.mul(A.unsqueeze(-1), x.unsqueeze(1))
        W_new = mlp("e_func", torch.cat((W, W1), dim=-1))
    else:
        W_new = W
    if norm:
        A = torch.nn.functional.normalize(A, p=1, dim=2)
    x1 = mlp("n_func", x)
    x2 = torch.matmul((A.unsqueeze(-1) * W_new).permute(0, 3, 1, 2),
                      x1.unsqueeze(2).permute(0, 3, 1, 2)).squeeze(-1).transpose(1, 2)
    x2 = x2 + mlp("n_self_func", x)
    if "classifier.weight" in p:
        skc = p["classifier.weight"].shape[0]
        x3 = lin(x2, p["classifier.weight"], p["classifier.bias"])
        n1_rep = torch.repeat_interleave(n1, skc, dim=0)
        n2_rep = torch.repeat_interleave(n2, skc, dim=0)
        n1m, n2m = int(n1.max()), int(n2.max())
        x4 = x3.permute(0, 2, 1).reshape(x.shape[0] * skc, n2m, n1m).transpose(1, 2)
        x5 = sinkhorn(x4, n1_rep, n2_rep, dummy_row=True, max_iter=sk_iter, tau=sk_tau).transpose(2, 1).contiguous()
        x6 = x5.reshape(x.shape[0], skc, n1m * n2m).permute(0, 2, 1)
        return W_new, torch.cat((x2, x6), dim=-1)
    return W_new, x2
