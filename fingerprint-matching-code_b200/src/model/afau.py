"""Drop-in for the live part of ``/root/reference/src/model/afau.py`` (``Encoder`` and its blocks).

Same modules, parameter names and ``Encoder.forward(row_emb, col_emb, cost_mat)`` signature.  The
reference's attention materialises ``[B, n, 16, n, 16]`` tensors (``afau.py:262-282``); here each block is
q/k/v GEMMs -> one fused mixed-score attention kernel -> combine GEMM -> add+InstanceNorm kernel ->
feed-forward GEMMs -> add+InstanceNorm kernel (``csrc/afau.cu``, ``csrc/gemm_*.cu``).

``Encoder.forward_k_inputs`` is the form ``Net.forward`` uses: it exploits that the row embedding is all
zeros and the column embedding one-hot (``ngm.py:392-399``) - projections of zeros are zeros and of a
one-hot vector a weight column, both exactly - and returns only the per-channel maxima the k head needs.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from fpmatch import ops


class Encoder(nn.Module):
    """AFA-U graph attention module to generate bipartite node embeddings."""
    model_params = {
        'embedding_dim': 600,
        'head_num': 16,
        'qkv_dim': 16,
        'ff_hidden_dim': 256,
        'ms_hidden_dim': 16,
        'ms_layer1_init': 10,
        'ms_layer2_init': 10,
        'sqrt_qkv_dim': math.sqrt(16),
    }

    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([EncoderLayer(**self.model_params)])

    def forward(self, row_emb, col_emb, cost_mat):
        for layer in self.layers:
            row_emb, col_emb = layer(row_emb, col_emb, cost_mat)
        return row_emb, col_emb

    def forward_k_inputs(self, cost_mat, n2, n1max, n2max, cost_t=None):
        """(max over rows of row block output [B,600], max over rows of col block output [B,600]) for
        row_emb = 0 [B,n1max,600] and col_emb = one-hot(j < n2_b) [B,n2max,600].  ``cost_t`` = the contiguous
        transposed copy of ``cost_mat`` when the caller has one (the Sinkhorn kernel writes it for free): the attention
        kernel then reads a column of 32 rows with one coalesced load and stages nothing."""
        assert len(self.layers) == 1
        layer = self.layers[0]
        g_row = layer.row_encoding_block.forward_zero_rows(cost_mat, n2, n1max, n2max, cost_t=cost_t)
        g_col = layer.col_encoding_block.forward_onehot_rows_zero_cols(n2, n2max)
        return g_row, g_col


    def forward_k_inputs_train(self, cost_mat, n2, n1max, n2max):
        """Differentiable twin of ``forward_k_inputs`` (stages 2-5 of train.py train this branch): the same kernels
        behind ``fpmatch.autograd`` Functions.  Returns (g_row, g_col, zero) where ``zero`` is an exact 0 that is
        connected to the parameters whose gradient is identically zero in the reference's graph (projections of the
        all-zero row embedding, the column block's attention): autograd then hands the optimiser zero tensors rather
        than ``None`` for them, exactly as it does in the reference (AdamW still applies weight decay to those)."""
        assert len(self.layers) == 1
        layer = self.layers[0]
        g_row = layer.row_encoding_block.forward_zero_rows_train(cost_mat, n2, n1max, n2max)
        g_col = layer.col_encoding_block.forward_onehot_rows_zero_cols_train(n2, n2max)
        rb, cb = layer.row_encoding_block, layer.col_encoding_block
        dead = [rb.Wq.weight, cb.Wq.weight, cb.Wk.weight, cb.Wv.weight, cb.multi_head_combine.weight,
                cb.mixed_score_MHA.mix1_weight, cb.mixed_score_MHA.mix1_bias, cb.mixed_score_MHA.mix2_weight,
                cb.mixed_score_MHA.mix2_bias]
        zero = sum((p.sum() * 0.0 for p in dead))
        return g_row, g_col, zero


class EncoderLayer(nn.Module):
    def __init__(self, **model_params):
        super().__init__()
        self.row_encoding_block = EncodingBlock(**model_params)
        self.col_encoding_block = EncodingBlock(**model_params)

    def forward(self, row_emb, col_emb, cost_mat):
        row_emb_out = self.row_encoding_block(row_emb, col_emb, cost_mat, transposed_cost=False)
        col_emb_out = self.col_encoding_block(col_emb, row_emb, cost_mat, transposed_cost=True)
        return row_emb_out, col_emb_out


def _lin(x3, weight, bias=None, act=0):
    B, n, K = x3.shape
    # the Parameter object itself is handed down: the operand-split cache is keyed on the live tensor object, and
    # .detach() would mint a new one per call
    w = weight if weight.is_contiguous() else weight.detach().contiguous()
    out = ops.gemm_nt(x3.reshape(B * n, K), w, None if bias is None else bias.detach().contiguous(), act,
                      weight_operand=True)
    return out.view(B, n, -1)


class EncodingBlock(nn.Module):
    def __init__(self, **model_params):
        super().__init__()
        self.model_params = model_params
        embedding_dim = self.model_params['embedding_dim']
        head_num = self.model_params['head_num']
        qkv_dim = self.model_params['qkv_dim']
        self.Wq = nn.Linear(embedding_dim, head_num * qkv_dim, bias=False)
        self.Wk = nn.Linear(embedding_dim, head_num * qkv_dim, bias=False)
        self.Wv = nn.Linear(embedding_dim, head_num * qkv_dim, bias=False)
        self.mixed_score_MHA = CrossSet_MultiHeadAttention(**model_params)
        self.multi_head_combine = nn.Linear(head_num * qkv_dim, embedding_dim)
        self.add_n_normalization_1 = AddAndInstanceNormalization(**model_params)
        self.feed_forward = FeedForward(**model_params)
        self.add_n_normalization_2 = AddAndInstanceNormalization(**model_params)

    def _tail(self, row_emb, attn_out, want_rowmax=False):
        """combine -> add+norm -> feed-forward -> add+norm (afau.py:133-139)."""
        mh = _lin(attn_out, self.multi_head_combine.weight, self.multi_head_combine.bias)
        out1 = self.add_n_normalization_1(row_emb, mh)
        out2 = self.feed_forward(out1)
        return self.add_n_normalization_2(out1, out2, want_rowmax=want_rowmax)

    def forward(self, row_emb, col_emb, cost_mat, transposed_cost=False):
        """``cost_mat`` is always the un-transposed [B, n1, n2] matrix; the column block sets
        ``transposed_cost`` instead of receiving a strided view."""
        row_emb = row_emb.detach().to(torch.float32).contiguous()
        col_emb = col_emb.detach().to(torch.float32).contiguous()
        cost = cost_mat.detach().to(torch.float32)
        if not cost.is_contiguous():
            # a caller following the reference passes cost.transpose(1, 2) to the column block
            if cost.transpose(1, 2).is_contiguous():
                cost, transposed_cost = cost.transpose(1, 2), not transposed_cost
            else:
                cost = cost.contiguous()
        q = _lin(row_emb, self.Wq.weight)
        k = _lin(col_emb, self.Wk.weight)
        v = _lin(col_emb, self.Wv.weight)
        att = self.mixed_score_MHA(q, k, v, cost, transposed_cost=transposed_cost)
        return self._tail(row_emb, att)

    # ---- structured forms used by Net.forward --------------------------------------------------
    def forward_zero_rows(self, cost_mat, n2, n1max, n2max, cost_t=None):
        """Row block with row_emb = 0 and col_emb = one-hot: q = 0, k/v = weight columns."""
        B = cost_mat.shape[0]
        dev = cost_mat.device
        E = self.Wq.out_features
        v = ops.onehot_proj(self.Wv.weight.detach().contiguous(), n2, n2max)
        if ops.afau_zero_query_kernel_fits(n1max, n2max):
            q = k = None                                   # the zero-query kernel reads neither
        else:
            q = torch.zeros((B, n1max, E), dtype=torch.float32, device=dev)
            k = ops.onehot_proj(self.Wk.weight.detach().contiguous(), n2, n2max)
        if cost_t is not None and q is None:
            att = self.mixed_score_MHA(q, k, v, cost_t, transposed_cost=True, q_zero=True)
        else:
            att = self.mixed_score_MHA(q, k, v, cost_mat, transposed_cost=False, q_zero=True)
        mh = _lin(att, self.multi_head_combine.weight, self.multi_head_combine.bias)
        out1 = self.add_n_normalization_1(mh, None)            # row_emb + mh with row_emb = 0
        out2 = self.feed_forward(out1)
        _, rowmax = self.add_n_normalization_2(out1, out2, want_rowmax=True, want_out=False)
        return rowmax

    def forward_onehot_rows_zero_cols(self, n2, n2max):
        """Column block: its keys/values are projections of the zero row embedding, so the attention
        output is exactly 0 and multi_head_combine contributes only its bias."""
        B = n2.shape[0]
        dev = n2.device
        emb = self.Wq.in_features
        norm = self.add_n_normalization_1.norm
        if n2max <= 112 and emb % 4 == 0 and n2max <= emb:
            out1 = ops.onehot_instnorm(n2, n2max, self.multi_head_combine.bias.detach().contiguous(),
                                       norm.weight.detach().contiguous(), norm.bias.detach().contiguous(), norm.eps)
        else:
            onehot = ops.onehot_proj(torch.eye(emb, dtype=torch.float32, device=dev), n2, n2max)
            out1 = self.add_n_normalization_1(onehot, self.multi_head_combine.bias.detach().contiguous())
        out2 = self.feed_forward(out1)
        _, rowmax = self.add_n_normalization_2(out1, out2, want_rowmax=True, want_out=False)
        return rowmax


    # ---- the same two forms, differentiable ---------------------------------------------------------------------
    def _tail_train(self, first, second):
        from fpmatch import autograd as fa
        n1, n2 = self.add_n_normalization_1.norm, self.add_n_normalization_2.norm
        out1 = fa.AddInstNormFn.apply(first, second, n1.weight, n1.bias, n1.eps, False)
        ff = self.feed_forward
        hid = fa.LinearFn.apply(out1, ff.W1.weight, ff.W1.bias, 1)
        out2 = fa.LinearFn.apply(hid, ff.W2.weight, ff.W2.bias, 0)
        _, rowmax = fa.AddInstNormFn.apply(out1, out2, n2.weight, n2.bias, n2.eps, True)
        return rowmax

    def forward_zero_rows_train(self, cost_mat, n2, n1max, n2max):
        from fpmatch import autograd as fa
        B, dev = cost_mat.shape[0], cost_mat.device
        q = torch.zeros((B, n1max, self.Wq.out_features), dtype=torch.float32, device=dev)
        k = fa.OnehotProjFn.apply(self.Wk.weight, n2, n2max)
        v = fa.OnehotProjFn.apply(self.Wv.weight, n2, n2max)
        m = self.mixed_score_MHA
        att = fa.AfauAttentionFn.apply(q, k, v, cost_mat, False, m.mix1_weight, m.mix1_bias, m.mix2_weight, m.mix2_bias)
        mh = fa.LinearFn.apply(att, self.multi_head_combine.weight, self.multi_head_combine.bias, 0)
        return self._tail_train(mh, None)

    def forward_onehot_rows_zero_cols_train(self, n2, n2max):
        dev = n2.device
        onehot = ops.onehot_proj(torch.eye(self.Wq.in_features, dtype=torch.float32, device=dev), n2, n2max)
        return self._tail_train(onehot, self.multi_head_combine.bias)


class AddAndInstanceNormalization(nn.Module):
    def __init__(self, **model_params):
        super().__init__()
        embedding_dim = model_params['embedding_dim']
        self.norm = nn.InstanceNorm1d(embedding_dim, affine=True, track_running_stats=False)

    def forward(self, input1, input2, want_rowmax=False, want_out=True):
        return ops.add_instnorm(input1.contiguous(), None if input2 is None else input2.contiguous(),
                                self.norm.weight.detach().contiguous(), self.norm.bias.detach().contiguous(),
                                want_rowmax=want_rowmax, eps=self.norm.eps, want_out=want_out)


class FeedForward(nn.Module):
    def __init__(self, **model_params):
        super().__init__()
        embedding_dim = model_params['embedding_dim']
        ff_hidden_dim = model_params['ff_hidden_dim']
        self.W1 = nn.Linear(embedding_dim, ff_hidden_dim)
        self.W2 = nn.Linear(ff_hidden_dim, embedding_dim)

    def forward(self, input1):
        return _lin(_lin(input1, self.W1.weight, self.W1.bias, act=1), self.W2.weight, self.W2.bias)


class CrossSet_MultiHeadAttention(nn.Module):
    def __init__(self, **model_params):
        super().__init__()
        self.model_params = model_params
        head_num = model_params['head_num']
        ms_hidden_dim = model_params['ms_hidden_dim']
        mix1_init = model_params['ms_layer1_init']
        mix2_init = model_params['ms_layer2_init']
        U = torch.distributions.Uniform
        self.mix1_weight = nn.Parameter(U(low=-mix1_init, high=mix1_init).sample((head_num, 2, ms_hidden_dim)))
        self.mix1_bias = nn.Parameter(U(low=-mix1_init, high=mix1_init).sample((head_num, ms_hidden_dim)))
        self.mix2_weight = nn.Parameter(U(low=-mix2_init, high=mix2_init).sample((head_num, ms_hidden_dim, 1)))
        self.mix2_bias = nn.Parameter(U(low=-mix2_init, high=mix2_init).sample((head_num, 1)))

    def forward(self, q, k, v, cost_mat, transposed_cost=False, q_zero=False):
        """q [B, nr, 256] (heads concatenated), k/v [B, nc, 256]; returns [B, nr, 256].  ``q_zero``: the caller
        guarantees q == 0, the kernel then skips the q.k products (identical result)."""
        if q is not None and q.dim() == 4:      # reference layout [B, H, n, d]
            q, k, v = (t.transpose(1, 2).reshape(t.shape[0], t.shape[2], -1) for t in (q, k, v))
        d = lambda t: t.detach().contiguous()
        if q is not None:
            q, k = q.contiguous(), k.contiguous()
        return ops.afau_attention(q, k, v.contiguous(), d(cost_mat), transposed_cost,
                                  d(self.mix1_weight), d(self.mix1_bias), d(self.mix2_weight), d(self.mix2_bias),
                                  q_zero=q_zero)


def reshape_by_heads(qkv, head_num):
    """[B, n, H*d] -> [B, H, n, d] (afau.py helper kept for importers)."""
    batch_s, n = qkv.size(0), qkv.size(1)
    return qkv.reshape(batch_s, n, head_num, -1).transpose(1, 2)
