"""CPU restatement of ``Net.forward`` (``/root/reference/src/model/ngm.py:205-491``) from the backbone
feature maps onward.  Test infrastructure - see the package docstring.

``forward_head(params, data, fmaps, ...)`` takes the reference-named ``state_dict`` of the model under
test, the ``data_dict`` (CPU tensors) and the raw layer3/layer4 feature maps of both images, and returns
the same keys ``Net.forward`` adds (``ds_mat``, ``perm_mat``, ``k_prob``, ``cls_prob`` ...) plus the
intermediates the parity tests compare stage by stage.  The per-pair / per-layer python loops of the
reference are kept (they are what its CPU path executes and what ``bench.py``'s CPU baseline times).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import ops

Tensor = torch.Tensor

SK_TAU = 0.01          # ngm.py:45
SK_ITER_NUM = 10       # ngm.py:53
GNN_SK_ITER = 20       # gnn.py:173 default, never overridden at ngm.py:151-157
GNN_LAYERS = 3         # ngm.py:48
UNIV_SIZE = 600        # ngm.py:52
K_FACTOR = 50.0        # ngm.py:55
RESCALE = (320, 240)   # ngm.py:160


def _pad_stack(ts: List[Tensor]) -> Tensor:
    """torch.stack(pad_tensor(list)) of /root/reference/utils/pad_tensor.py:5-31."""
    shape = [max(t.shape[i] for t in ts) for i in range(ts[0].dim())]
    out = torch.zeros([len(ts)] + shape, dtype=ts[0].dtype)
    for b, t in enumerate(ts):
        out[(b,) + tuple(slice(0, s) for s in t.shape)] = t
    return out


def forward_head(p: Dict[str, Tensor], data: dict, fmaps, regression: bool = True,
                 training: bool = False, feature_align_loops: bool = False,
                 compute_ke: bool = True, stable_sort: bool = True,
                 dtype: torch.dtype = torch.float32, keep_graph: bool = False) -> dict:
    """``keep_graph=True`` leaves the parameters attached so that torch autograd differentiates this restatement
    exactly as it differentiates the reference's forward (used by oracle/train.py)."""
    if not keep_graph:
        p = {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach()) for k, v in p.items()}
    points, n_points, graphs = data["Ps"], data["ns"], data["pyg_graphs"]
    B = data["gt_perm_mat"].shape[0]
    fa = ops.feature_align_loop if feature_align_loops else ops.feature_align
    inter = {}

    global_list, graph_feats, edge_feats = [], [], []
    for gi, ((nodes, edges), P, ns, graph) in enumerate(zip(fmaps, points, n_points, graphs)):
        nodes = nodes.to(dtype); edges = edges.to(dtype)
        global_list.append(F.adaptive_max_pool2d(edges, (1, 1)).reshape(B, -1))          # :238
        nodes_n = ops.normalize_over_channels(nodes)                                     # :241-243
        edges_n = ops.normalize_over_channels(edges)
        if dtype == torch.float32:
            U = ops.concat_features(fa(nodes_n, P, ns, RESCALE), ns)                     # :246-247
            Fe = ops.concat_features(fa(edges_n, P, ns, RESCALE), ns)
        else:   # fp64 "truth" run: same formulae, interpolation weights still formed in fp32
            U = ops.concat_features(_fa_any(nodes_n, P, ns), ns)
            Fe = ops.concat_features(_fa_any(edges_n, P, ns), ns)
        x0 = torch.cat((U, Fe), dim=1)                                                   # :248
        inter[f"node_feat_{gi}"] = x0
        x = ops.sconv_residual(x0, graph.edge_index, graph.edge_attr.to(dtype), p,
                               "message_pass_node_features")                              # :254
        inter[f"sconv_{gi}"] = x
        ptr, eptr = graph.ptr, graph.eptr
        xs, es = [], []
        for b in range(B):                                                               # :255
            xb = x[int(ptr[b]):int(ptr[b + 1])]
            ei = graph.edge_index[:, int(eptr[b]):int(eptr[b + 1])] - int(ptr[b])
            xs.append(xb)
            es.append(xb[ei[0]] - xb[ei[1]])                         # spline_conv.py:73-81
        graph_feats.append(xs); edge_feats.append(es)

    gw = torch.cat([global_list[0], global_list[1]], dim=-1)                             # :262-268
    gw = ops.normalize_over_channels(gw)

    unary = [ops.affinity(X, Y, w, p["vertex_affinity.A.weight"], p["vertex_affinity.A.bias"])
             for X, Y, w in zip(graph_feats[0], graph_feats[1], gw)]                      # :277-280
    if compute_ke:                                                                        # :282-287
        quad = [0.5 * ops.affinity(X, Y, w, p["edge_affinity.A.weight"], p["edge_affinity.A.bias"])
                for X, Y, w in zip(edge_feats[0], edge_feats[1], gw)]
        inter["Ke"] = quad
    Kp = _pad_stack(unary)                                                                # :317
    inter["Kp"] = Kp
    n1max, n2max = Kp.shape[1], Kp.shape[2]
    N = n1max * n2max
    emb = Kp.transpose(1, 2).contiguous().view(B, -1, 1)                                  # :321

    qap = []
    for b in range(B):                                                                    # :326-348
        idxG, idxH = data["KGHs_sparse"][b]
        n1b, n2b = int(n_points[0][b]), int(n_points[1][b])
        diag = torch.arange(n1b * n2b, dtype=torch.long)     # linspace(...).long() of factorize_graph_matching.py:93-94
        row = torch.cat((idxG.long(), diag)); col = torch.cat((idxH.long(), diag))
        # ngm.py:333-342: K_value = [Ke_b.flatten(); Kp_b.flatten()] (Ke_b is e1 x e2 of the PyG edge lists) and all
        # three tensors are cut to the smallest length.  With a partial gt permutation G2 / H2 lose different
        # columns, the lists are longer than the value vector and the cut removes the tail of the diagonal block.
        e1b = int(graphs[0].eptr[b + 1] - graphs[0].eptr[b]); e2b = int(graphs[1].eptr[b + 1] - graphs[1].eptr[b])
        common_len = min(row.numel(), col.numel(), e1b * e2b + n1b * n2b)
        row, col = row[:common_len], col[:common_len]
        t = emb[b]
        for i in range(GNN_LAYERS):
            t = ops.pygnn_layer(t, row, col, n1b, n2b, n1max, n2max, p, f"gnn_layer_{i}",
                                sk_iter=GNN_SK_ITER, sk_tau=SK_TAU)
        qap.append(t)
    emb = torch.stack(qap, 0)                                                             # :362
    inter["emb"] = emb
    v = F.linear(emb, p["classifier.weight"], p["classifier.bias"])                       # :368
    s = v.view(B, n2max, -1).transpose(1, 2)                                              # :369
    ss = ops.sinkhorn(s, n_points[0], n_points[1], dummy_row=True, max_iter=SK_ITER_NUM, tau=SK_TAU)
    inter["s"] = s; inter["ss"] = ss

    min_pts = torch.tensor([int(min(n_points[0][b], n_points[1][b])) for b in range(B)],
                           dtype=torch.float32)                                           # :374-378
    gt_ks = torch.tensor([torch.sum(data["gt_perm_mat"][i]) for i in range(B)],
                         dtype=torch.float32)                                             # :381-384
    if regression:                                                                        # :386-412
        row0 = torch.zeros((B, int(torch.max(n_points[0])), UNIV_SIZE), dtype=dtype)
        col0 = torch.zeros((B, int(torch.max(n_points[1])), UNIV_SIZE), dtype=dtype)
        for b in range(B):
            nb = int(n_points[1][b])
            col0[b, torch.arange(nb), torch.arange(nb)] = 1
        out_r, out_c = ops.afau_encoder(row0, col0, ss.detach(), p)
        inter["afa_row"] = out_r; inter["afa_col"] = out_c
        g_r = out_r.max(dim=1).values      # pad rows to 600 with -inf then MaxPool1d(600) == max over rows
        g_c = out_c.max(dim=1).values
        kr = F.linear(F.relu(F.linear(g_r, p["final_row.0.weight"], p["final_row.0.bias"])),
                      p["final_row.2.weight"], p["final_row.2.bias"]).squeeze(-1)
        kc = F.linear(F.relu(F.linear(g_c, p["final_col.0.weight"], p["final_col.0.bias"])),
                      p["final_col.2.weight"], p["final_col.2.bias"]).squeeze(-1)
        ks = torch.sigmoid((kr + kc) / 2).to(torch.float32)
    else:
        ks = gt_ks / min_pts                                                              # :416

    k_for_topk = gt_ks.view(-1) if training else ks.view(-1) * min_pts                    # :418-439
    ss_out = ops.soft_topk_prob(ss, k_for_topk, SK_ITER_NUM, SK_TAU, n_points[0], n_points[1])
    x = ops.hungarian(ss_out.to(torch.float32), n_points[0], n_points[1])                 # :444
    top_indices = torch.argsort(x.mul(ss_out.detach().to(torch.float32)).reshape(B, -1), descending=True,
                                dim=-1, stable=stable_sort)                               # :445-447
    k_greedy = ks.view(-1) * min_pts
    perm = ops.greedy_perm(torch.zeros(ss_out.shape), top_indices, k_greedy)              # :448-449
    matched_sim = s * perm.to(dtype)                                                      # :451
    cls_logits = ops.match_classifier(matched_sim, p, training=training and keep_graph)
    cls_prob = torch.sigmoid(cls_logits)
    out = {
        "ds_mat": ss_out, "perm_mat": perm, "k_prob": ks, "cls_prob": cls_prob,
        "hungarian": x, "k_int": torch.tensor([round(k.item()) for k in k_greedy]),
        "inter": inter,
    }
    if "label" in data:
        out["cls_loss"] = F.binary_cross_entropy_with_logits(cls_logits.float(), data["label"].view(-1).float())
    if regression:
        sup = gt_ks / min_pts
        out["ks_loss"] = F.mse_loss(ks, sup) * K_FACTOR                                   # :465
        out["ks_error"] = F.l1_loss(ks * min_pts, gt_ks)                                  # :466
    else:
        out["ks_loss"] = 0.0; out["ks_error"] = 0.0
    return out


def _fa_any(fm: Tensor, P: Tensor, ns: Tensor) -> Tensor:
    """feature_align with taps in fm.dtype (weights formed in fp32 as the reference does)."""
    B, C, Hf, Wf = fm.shape
    n_max = P.shape[1]
    ori = torch.tensor(RESCALE, dtype=torch.float32)
    feat_size = torch.tensor([Hf, Wf], dtype=torch.float32)
    step = ori / feat_size
    pt = (P.to(torch.float32) - step / 2) / ori * feat_size
    x, y = pt[..., 0], pt[..., 1]
    x0 = torch.floor(x); x1 = x0 + 1; y0 = torch.floor(y); y1 = y0 + 1
    x0 = torch.clamp(x0, 0, Wf - 1); x1 = torch.clamp(x1, 0, Wf - 1)
    y0 = torch.clamp(y0, 0, Hf - 1); y1 = torch.clamp(y1, 0, Hf - 1)
    xi0, xi1, yi0, yi1 = x0.long(), x1.long(), y0.long(), y1.long()
    flat = fm.reshape(B, C, Hf * Wf)
    tap = lambda yi, xi: torch.gather(flat, 2, (yi * Wf + xi)[:, None, :].expand(B, C, n_max))
    Ia, Ib, Ic, Id = tap(yi0, xi0), tap(yi1, xi0), tap(yi0, xi1), tap(yi1, xi1)
    eqx, eqy = xi0 == xi1, yi0 == yi1
    x0 = torch.where(eqx & (xi0 == 0), x0 - 1, x0); x1 = torch.where(eqx & (xi0 != 0), x1 + 1, x1)
    y0 = torch.where(eqy & (yi0 == 0), y0 - 1, y0); y1 = torch.where(eqy & (yi0 != 0), y1 + 1, y1)
    dt = fm.dtype
    wa = ((x1 - x) * (y1 - y)).to(dt)[:, None, :]; wb = ((x1 - x) * (y - y0)).to(dt)[:, None, :]
    wc = ((x - x0) * (y1 - y)).to(dt)[:, None, :]; wd = ((x - x0) * (y - y0)).to(dt)[:, None, :]
    out = Ia * wa + Ib * wb + Ic * wc + Id * wd
    valid = (torch.arange(n_max)[None, :] < ns.view(-1, 1))[:, None, :]
    return torch.where(valid, out, torch.zeros((), dtype=dt))
