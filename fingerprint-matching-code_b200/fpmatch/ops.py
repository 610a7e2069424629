"""torch-tensor wrappers over the C ABI (``include/fpmatch.h``).

PyTorch is plumbing here: it owns device memory (caching allocator) and the current stream.  Every op
takes contiguous CUDA tensors, launches on ``torch.cuda.current_stream()``, never synchronises the
host and raises ``RuntimeError`` for CPU tensors - there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

Tensor = torch.Tensor

# dense-contraction engine: "3xtf32" / "3xf16" (tcgen05, error-compensated, fp32-faithful), "tf32" (tcgen05,
# 1 pass, opt-in), "fp32" (CUDA cores; also the on-device checker for the tensor-core kernels)
GEMM_MODES = ("fp32", "3xtf32", "3xf16", "tf32")
_GEMM_MODE = os.environ.get("FPMATCH_GEMM", "3xf16")
_LAUNCHES = 0          # kernels launched through this module (bench.py reports it)


def set_gemm_mode(mode: str) -> None:
    global _GEMM_MODE
    if mode not in GEMM_MODES:
        raise ValueError(f"unknown GEMM mode {mode!r}")
    _GEMM_MODE = mode


def gemm_mode() -> str:
    return _GEMM_MODE


_GEMM_PAIR = os.environ.get("FPMATCH_GEMM_PAIR", "1") != "0"
_SLAB_PLAN = os.environ.get("FPMATCH_SLAB_PLAN", "1") != "0"


def set_gemm_pair(on: bool) -> None:
    """True (default): the error-compensated modes use the persistent CTA-pair kernel; False: one tile per CTA."""
    global _GEMM_PAIR
    _lib.check(_lib.lib().fpm_gemm_set_pair(int(bool(on))), "fpm_gemm_set_pair")
    _GEMM_PAIR = bool(on)


def set_gemm_max_clusters(n: int) -> None:
    """Cap on the SM pairs the persistent slab GEMM occupies (0 = all 74).  Pipelines with several batches in flight
    leave a few SMs to the other batch's tail kernels (fpmatch.prefetch.MatchingPipeline sets 70)."""
    _lib.check(_lib.lib().fpm_gemm_set_max_clusters(int(n)), "fpm_gemm_set_max_clusters")


def gemm_pair_enabled() -> bool:
    return _GEMM_PAIR


def set_slab_plan(on: bool) -> None:
    """True (default): SplineConv computes only the (node, weight slab) products its edges read (tile-table GEMM);
    False: the dense 26-slab product."""
    global _SLAB_PLAN
    _SLAB_PLAN = bool(on)


def slab_plan_enabled() -> bool:
    return _SLAB_PLAN


def launch_count() -> int:
    return _LAUNCHES


_OP_TRACE = None        # diagnostic (bench.py --trace-ops): host time + name of every wrapper that launched something


def op_trace_start() -> None:
    global _OP_TRACE
    _OP_TRACE = []


def op_trace_stop():
    global _OP_TRACE
    t, _OP_TRACE = _OP_TRACE, None
    return t


def _count(n: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += n
    if _OP_TRACE is not None:
        import sys as _sys
        import time as _time
        _OP_TRACE.append((_time.perf_counter(), _sys._getframe(1).f_code.co_name))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _first_cuda_index(args):
    for x in args:
        if isinstance(x, Tensor):
            if x.is_cuda:
                return x.device.index
        elif isinstance(x, (list, tuple)):
            r = _first_cuda_index(x)
            if r is not None:
                return r
    return None


def on_tensor_device(fn):
    """The C ABI launches into the calling thread's CURRENT CUDA device and takes that device's current stream.  torch
    ops work on whatever device their tensors live on, and so must these (the reference's extension had no device guard
    at all): every public wrapper runs under ``torch.cuda.device(<device of its first CUDA
    tensor>)`` when that is not the current one.  ``_chk`` then refuses any tensor that is not on the current device,
    so a mixed-device call fails loudly instead of launching into the wrong context."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        idx = _first_cuda_index(args)
        if idx is None and kwargs:
            idx = _first_cuda_index(kwargs.values())
        if idx is None or idx == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(idx):
            return fn(*args, **kwargs)
    return wrapped


def _chk(t: Optional[Tensor], name: str, dtype=torch.float32) -> Optional[int]:
    if t is None:
        return None
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise RuntimeError(f"fpmatch: {name} must be a CUDA tensor (no CPU fallback exists)")
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"fpmatch: {name} is on {t.device} but the op runs on cuda:{torch.cuda.current_device()} "
                           "(all tensors of one op must live on one device)")
    if t.dtype != dtype:
        raise RuntimeError(f"fpmatch: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"fpmatch: {name} must be contiguous")
    return t.data_ptr()


def _ptr_array(ts: Sequence[Tensor]):
    arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    return arr


def _i64(t: Tensor) -> Tensor:
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.contiguous()


# ---------------------------------------------------------------------------------------------------
# feature_align family
# ---------------------------------------------------------------------------------------------------
def feature_align(raw_feature: Tensor, P: Tensor, ns: Tensor, ori_size, feat_coords: bool = False) -> Tensor:
    B, Cc, Hf, Wf = raw_feature.shape
    nmax = P.shape[1]
    ns = _i64(ns)
    out = torch.empty((B, Cc, nmax), dtype=torch.float32, device=raw_feature.device)
    rc = _lib.lib().fpm_feature_align(_chk(raw_feature, "raw_feature"), _chk(P, "P"), _chk(ns, "ns", torch.int64),
                                      out.data_ptr(), B, Cc, Hf, Wf, nmax, float(ori_size[0]), float(ori_size[1]),
                                      int(feat_coords), _stream())
    _lib.check(rc, "fpm_feature_align"); _count()
    return out


def fmap_prep(fmap: Tensor) -> Tensor:
    """NCHW raw map -> [B, H*W, C] channels-last, divided by the channel L2 norm."""
    B, Cc, Hf, Wf = fmap.shape
    out = torch.empty((B, Hf * Wf, Cc), dtype=torch.float32, device=fmap.device)
    rc = _lib.lib().fpm_fmap_prep(_chk(fmap, "fmap"), out.data_ptr(), B, Cc, Hf, Wf, _stream())
    _lib.check(rc, "fpm_fmap_prep"); _count()
    return out


def global_max_into(fmap: Tensor, out: Tensor, offset: int) -> None:
    B, Cc, Hf, Wf = fmap.shape
    rc = _lib.lib().fpm_global_max(_chk(fmap, "fmap"), _chk(out, "out"), B, Cc, Hf * Wf, out.shape[1], offset,
                                   _stream())
    _lib.check(rc, "fpm_global_max"); _count()


def node_features(nodes_nhwc: Tensor, edges_nhwc: Tensor, hw1, hw2, P: Tensor, ns: Tensor, ptr: Tensor,
                  total_nodes: int, ori_size) -> Tensor:
    B = P.shape[0]
    nmax = P.shape[1]
    C1, C2 = nodes_nhwc.shape[2], edges_nhwc.shape[2]
    X = torch.empty((total_nodes, C1 + C2), dtype=torch.float32, device=P.device)
    rc = _lib.lib().fpm_node_features(_chk(nodes_nhwc, "nodes"), _chk(edges_nhwc, "edges"), _chk(P, "P"),
                                      _chk(ns, "ns", torch.int64), _chk(ptr, "ptr", torch.int64), X.data_ptr(),
                                      B, nmax, C1, hw1[0], hw1[1], C2, hw2[0], hw2[1],
                                      float(ori_size[0]), float(ori_size[1]), _stream())
    _lib.check(rc, "fpm_node_features"); _count()
    return X


def affinity_coeff(gcat: Tensor, W: Tensor, bias: Tensor) -> Tensor:
    B, IN = gcat.shape
    OUT = W.shape[0]
    out = torch.empty((B, OUT), dtype=torch.float32, device=gcat.device)
    rc = _lib.lib().fpm_affinity_coeff(_chk(gcat, "gcat"), _chk(W, "W"), _chk(bias, "bias"), out.data_ptr(),
                                       B, IN, OUT, _stream())
    _lib.check(rc, "fpm_affinity_coeff"); _count()
    return out


# ---------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------
def gemm_nt(A: Tensor, Bt: Tensor, bias: Optional[Tensor] = None, act: int = 0, out: Optional[Tensor] = None,
            mode: Optional[str] = None, weight_operand: bool = False) -> Tensor:
    """out[M,N] = act(A[M,K] @ Bt[N,K]^T + bias).  ``weight_operand`` marks Bt as a static weight whose
    tf32 split may be cached between calls."""
    mode = mode or _GEMM_MODE
    M, K = A.shape
    N = Bt.shape[0]
    assert Bt.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    a, b, c = _chk(A, "A"), _chk(Bt, "Bt"), _chk(out, "out")
    bp = _chk(bias, "bias")
    L = _lib.lib()
    ev = None
    if mode == "fp32":
        if _GEMM_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)); ev[0].record()
        rc = L.fpm_gemm_nt_f32(a, b, bp, c, M, N, K, K, K, N, act, _stream())
        _lib.check(rc, "fpm_gemm_nt_f32"); _count()
    elif mode == "tf32":
        if _GEMM_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)); ev[0].record()
        rc = L.fpm_gemm_nt_tc(a, None, b, None, bp, c, M, N, K, K, K, N, act, 1, _stream())
        _lib.check(rc, "fpm_gemm_nt_tc"); _count()
    elif mode == "3xf16":
        a_hi, a_lo, a_inv = f16_split_rows(A)
        b_hi, b_lo, b_inv = f16_split_rows(Bt, cache=weight_operand)
        if _GEMM_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)); ev[0].record()
        rc = L.fpm_gemm_nt_f16x3(a_hi.data_ptr(), a_lo.data_ptr(), a_inv.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(),
                                 b_inv.data_ptr(), bp, c, M, N, K, K, K, N, act, _stream())
        _lib.check(rc, "fpm_gemm_nt_f16x3"); _count()
    else:
        a_hi, a_lo = tf32_split(A)
        b_hi, b_lo = tf32_split(Bt, cache=weight_operand)
        if _GEMM_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)); ev[0].record()
        rc = L.fpm_gemm_nt_tc(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), bp, c, M, N, K,
                              K, K, N, act, 3, _stream())
        _lib.check(rc, "fpm_gemm_nt_tc"); _count()
    if ev is not None:
        ev[1].record()
        _GEMM_EVENTS.append((mode, M, N, K, ev[0], ev[1]))
    return out


_GEMM_EVENTS = None
_SPLIT_CACHE = {}


def _cache_get(kind: str, x: Tensor):
    """Split of a static weight, valid only while the SAME tensor object is alive and unmodified.  (Keying on
    data_ptr alone served a freed tensor's split to a new tensor allocated at the same address.)"""
    ent = _SPLIT_CACHE.get((kind, id(x)))
    if ent is not None and ent[0]() is x and ent[1] == x._version and ent[2] == x.data_ptr():
        return ent[3]
    return None


def _cache_put(kind: str, x: Tensor, value) -> None:
    dead = [k for k, e in _SPLIT_CACHE.items() if e[0]() is None]
    for k in dead:
        del _SPLIT_CACHE[k]
    if len(_SPLIT_CACHE) > 64:
        _SPLIT_CACHE.clear()
    _SPLIT_CACHE[(kind, id(x))] = (weakref.ref(x), x._version, x.data_ptr(), value)


def gemm_profile_start() -> None:
    """Record a CUDA-event pair around every GEMM kernel launch (bench.py's live roofline measurement)."""
    global _GEMM_EVENTS
    _GEMM_EVENTS = []


def gemm_profile_stop():
    global _GEMM_EVENTS
    ev, _GEMM_EVENTS = _GEMM_EVENTS, None
    return ev


def f16_split_rows(x: Tensor, cache: bool = False, out=None):
    """x[r,:] * s_r = hi + 2^-11 lo with hi, lo fp16 and s_r a power of two; returns (hi, lo, 1/s).
    ``out`` = (hi, lo, inv) buffers with at least x.shape[0] rows to write into (rows beyond are left untouched)."""
    if cache:
        hit = _cache_get("f16", x)
        if hit is not None:
            return hit
    rows, K = x.shape
    if out is not None:
        hi, lo, inv = out
    else:
        hi = torch.empty((rows, K), dtype=torch.float16, device=x.device)
        lo = torch.empty((rows, K), dtype=torch.float16, device=x.device)
        inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
    rc = _lib.lib().fpm_f16_split_rows(_chk(x, "x"), hi.data_ptr(), lo.data_ptr(), inv.data_ptr(), rows, K, _stream())
    _lib.check(rc, "fpm_f16_split_rows"); _count()
    if cache:
        _cache_put("f16", x, (hi, lo, inv))
    return hi, lo, inv


def tf32_split(x: Tensor, cache: bool = False):
    """x = hi + lo with both parts exactly representable in tf32.  ``cache=True`` memoises the split of a
    weight matrix (keyed on storage + version), so static weights are split once, not per forward."""
    if cache:
        hit = _cache_get("tf32", x)
        if hit is not None:
            return hit
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    rc = _lib.lib().fpm_tf32_split(_chk(x, "x"), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "fpm_tf32_split"); _count()
    if cache:
        _cache_put("tf32", x, (hi, lo))
    return hi, lo


# ---------------------------------------------------------------------------------------------------
# SplineConv
# ---------------------------------------------------------------------------------------------------
def csr_by_dst(edge_index: Tensor, ptr: Tensor, eptr: Tensor, total_nodes: int, max_edges: int):
    E = edge_index.shape[1]
    dev = edge_index.device
    in_ptr = torch.empty((total_nodes + 1,), dtype=torch.int32, device=dev)
    in_eid = torch.empty((max(E, 1),), dtype=torch.int32, device=dev)
    dst = edge_index[1]
    rc = _lib.lib().fpm_csr_by_dst(_chk(dst, "edge_index[1]", torch.int64), _chk(ptr, "ptr", torch.int64),
                                   _chk(eptr, "eptr", torch.int64), in_ptr.data_ptr(), in_eid.data_ptr(),
                                   ptr.numel() - 1, total_nodes, max_edges, _stream())
    _lib.check(rc, "fpm_csr_by_dst"); _count()
    return in_ptr, in_eid


class SlabPlan:
    """Device-side plan of one graph batch for the slab-sparse SplineConv GEMM (csrc/spline.cu, planner): which
    (node block, weight slab) tiles to compute.  Depends on the graph only, so both conv layers share it."""

    def __init__(self, edge_index: Tensor, pseudo: Tensor, total: int, channels: int, kernel_size: int = 5):
        dev = edge_index.device
        NS = kernel_size * kernel_size + 1
        self.T, self.C, self.NS = total, channels, NS
        self.T_pad = (total + 255) // 256 * 256
        tiles_m, ntn = self.T_pad // 256, channels // 128
        # sparse slabs hold < T/4 nodes each (padded to 256-row groups)
        self.rowmap_cap = (NS - 1) * ((total // 4 + 255) // 256 * 256 + 256)
        self.max_tiles = tiles_m * NS * ntn + (self.rowmap_cap // 256) * ntn
        self.mask = torch.zeros((total,), dtype=torch.int32, device=dev)
        self.meta = torch.zeros((2 + 3 * NS + 2 + 33,), dtype=torch.int32, device=dev)    # + planner scratch
        self.tab = torch.empty((self.max_tiles, 4), dtype=torch.int32, device=dev)
        self.rowmap = torch.full((self.rowmap_cap,), -1, dtype=torch.int32, device=dev)
        src = edge_index[0]
        rc = _lib.lib().fpm_spline_plan(_chk(src, "edge_index[0]", torch.int64), _chk(pseudo, "pseudo"),
                                        self.mask.data_ptr(), self.meta.data_ptr(), self.tab.data_ptr(),
                                        self.rowmap.data_ptr(), total, edge_index.shape[1], channels, kernel_size,
                                        self.max_tiles, self.rowmap_cap, _stream())
        _lib.check(rc, "fpm_spline_plan"); _count(3)

    def tiles_used(self) -> int:
        """Number of 256 x 128 tiles the GEMM computes (forces a device sync; for reporting only)."""
        return int(self.meta[0].item())


class _PlanInfo:
    def __init__(self, plan: SlabPlan):
        self.T, self.T_pad, self.meta = plan.T, plan.T_pad, plan.meta

    def tiles_used(self) -> int:
        return int(self.meta[0].item())


_GATHER_SPLIT = os.environ.get("FPMATCH_GATHER_SPLIT", "1") != "0"


def gather_split_enabled() -> bool:
    """True (default): the hidden layer of an SConv is written by the gather kernel directly as the fp16 operand of the
    next slab GEMM; FPMATCH_GATHER_SPLIT=0 keeps the fp32 tensor + split pass (A/B runs)."""
    return _GATHER_SPLIT


def slab_operand_buffers(plan: "SlabPlan", K: int, device):
    """Uninitialised A-operand buffers (fp16 hi, lo, fp32 row scales) of ``spline_slab_gemm`` for ``plan``: the dense
    node rows first, the compacted rows of the sparse slabs behind them."""
    rows = plan.T_pad + plan.rowmap_cap
    return (torch.empty((rows, K), dtype=torch.float16, device=device),
            torch.empty((rows, K), dtype=torch.float16, device=device),
            torch.empty((rows,), dtype=torch.float32, device=device))


def spline_slab_gemm(x: Optional[Tensor], packed: Tensor, plan: SlabPlan, presplit=None) -> Tensor:
    """Y [total, NS*C] = x @ packed^T restricted to the (node, slab) blocks of ``plan`` (other blocks are left
    uninitialised: no edge reads them).  Error-compensated fp16 on the persistent CTA-pair kernel.  ``presplit`` =
    buffers from ``slab_operand_buffers`` whose first T rows already hold the split of x (written by the previous
    layer's ``spline_gather_max``): the split pass over x is skipped."""
    K = packed.shape[1]
    T = plan.T
    assert packed.shape[0] == plan.NS * plan.C
    dev = packed.device
    rows = plan.T_pad + plan.rowmap_cap
    if presplit is None:
        assert x.shape == (T, K)
        a_hi, a_lo, a_inv = slab_operand_buffers(plan, K, dev)
        f16_split_rows(x, out=(a_hi, a_lo, a_inv))
    else:
        a_hi, a_lo, a_inv = presplit
        assert a_hi.shape == (rows, K) and a_lo.shape == (rows, K) and a_inv.shape == (rows,)
    L = _lib.lib()
    rc = L.fpm_spline_gather_rows(plan.meta.data_ptr(), plan.rowmap.data_ptr(), a_hi.data_ptr(), a_lo.data_ptr(),
                                  a_inv.data_ptr(), plan.T_pad, K, plan.rowmap_cap, _stream())
    _lib.check(rc, "fpm_spline_gather_rows"); _count()
    b_hi, b_lo, b_inv = f16_split_rows(packed, cache=True)
    Y = torch.empty((T, plan.NS * plan.C), dtype=torch.float32, device=dev)
    ev = None
    if _GEMM_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)); ev[0].record()
    rc = L.fpm_gemm_nt_f16x3_tiles(a_hi.data_ptr(), a_lo.data_ptr(), a_inv.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(),
                                   b_inv.data_ptr(), Y.data_ptr(), rows, packed.shape[0], K, K, K, Y.shape[1],
                                   plan.tab.data_ptr(), plan.meta.data_ptr(), plan.rowmap.data_ptr(), plan.max_tiles,
                                   T, _stream())
    _lib.check(rc, "fpm_gemm_nt_f16x3_tiles"); _count()
    if ev is not None:
        ev[1].record()
        # keep only the 88-int meta tensor alive (holding the whole plan would pin its buffers in the allocator)
        _GEMM_EVENTS.append(("3xf16-slabs", _PlanInfo(plan), packed.shape[0], K, ev[0], ev[1]))
    return Y


def spline_gather_max(Y: Tensor, xin: Optional[Tensor], edge_index: Tensor, pseudo: Tensor, in_ptr: Tensor,
                      in_eid: Tensor, bias: Tensor, mode: int, kernel_size: int = 5, want_argmax: bool = False,
                      split_out=None, want_out: bool = True):
    """``split_out`` = (hi, lo, inv) from ``slab_operand_buffers``: the result rows are also written as the
    error-compensated fp16 operand of the next layer's slab GEMM (bit-identical to ``f16_split_rows`` of the fp32
    result); with ``want_out=False`` the fp32 tensor is then not written at all (returned as None)."""
    total, Cc = Y.shape[0], bias.shape[0]
    want_out = want_out or split_out is None
    out = torch.empty((total, Cc), dtype=torch.float32, device=Y.device) if want_out else None
    arg = torch.empty((total, Cc), dtype=torch.int32, device=Y.device) if want_argmax else None
    if split_out is not None:
        hi, lo, inv = split_out
        assert hi.shape[1] == Cc and hi.shape[0] >= total and hi.dtype == torch.float16 and hi.is_contiguous()
        assert lo.shape == hi.shape and lo.is_contiguous() and inv.shape[0] == hi.shape[0]
    rc = _lib.lib().fpm_spline_gather_max(_chk(Y, "Y"), _chk(xin, "xin"), _chk(edge_index[0], "edge_index[0]", torch.int64),
                                          _chk(pseudo, "pseudo"), _chk(in_ptr, "in_ptr", torch.int32),
                                          _chk(in_eid, "in_eid", torch.int32), _chk(bias, "bias"),
                                          out.data_ptr() if want_out else None,
                                          arg.data_ptr() if want_argmax else None,
                                          hi.data_ptr() if split_out is not None else None,
                                          lo.data_ptr() if split_out is not None else None,
                                          inv.data_ptr() if split_out is not None else None,
                                          total, Cc, kernel_size, mode, _stream())
    _lib.check(rc, "fpm_spline_gather_max"); _count()
    return (out, arg) if want_argmax else out


# ---------------------------------------------------------------------------------------------------
# affinities
# ---------------------------------------------------------------------------------------------------
_AFFINITY_TC = os.environ.get("FPMATCH_AFFINITY_TC", "1") != "0"


def set_affinity_tc(on: bool) -> None:
    """True (default): node affinities on the tcgen05 tile-table GEMM (csrc/affinity_tc.cu); False: CUDA-core kernel."""
    global _AFFINITY_TC
    _AFFINITY_TC = bool(on)


def affinity_nodes(XA: Tensor, XB: Tensor, coeff: Tensor, ptrA: Tensor, ptrB: Tensor, Rmax: int, Cmax: int,
                   scale: float = 1.0, want_t: bool = True, raw: bool = False):
    """softplus((XA_b (.) coeff_b) XB_b^T) - 0.5 per pair, zero padded to [B,Rmax,Cmax] (+ the transposed copy)."""
    B = coeff.shape[0]
    out = torch.empty((B, Rmax, Cmax), dtype=torch.float32, device=XA.device)
    out_t = torch.empty((B, Cmax, Rmax), dtype=torch.float32, device=XA.device) if want_t else None
    K = XA.shape[1]
    if _AFFINITY_TC and _GEMM_MODE == "3xf16" and _GEMM_PAIR and K % 64 == 0 and B > 0 and XA.shape[0] > 0 \
            and XB.shape[0] > 0:
        L = _lib.lib()
        dev = XA.device
        tA, tB = (Rmax + 255) // 256, (Cmax + 127) // 128
        a_hi = torch.empty((XA.shape[0], K), dtype=torch.float16, device=dev)
        a_lo = torch.empty_like(a_hi)
        a_inv = torch.empty((XA.shape[0],), dtype=torch.float32, device=dev)
        rc = L.fpm_f16_split_rows_scaled(_chk(XA, "XA"), _chk(coeff, "coeff"), _chk(ptrA, "ptrA", torch.int64), B,
                                         a_hi.data_ptr(), a_lo.data_ptr(), a_inv.data_ptr(), XA.shape[0], K, _stream())
        _lib.check(rc, "fpm_f16_split_rows_scaled"); _count()
        b_hi, b_lo, b_inv = f16_split_rows(XB, cache=True)      # Kp and the raw products for Ke share X2's split
        ntile = B * tA * tB
        tab = torch.empty((ntile, 4), dtype=torch.int32, device=dev)
        meta = torch.empty((1,), dtype=torch.int32, device=dev)
        rowmap = torch.empty((B * tA * 256,), dtype=torch.int32, device=dev)
        rc = L.fpm_affinity_tiles(ptrA.data_ptr(), _chk(ptrB, "ptrB", torch.int64), B, Rmax, Cmax, tab.data_ptr(),
                                  meta.data_ptr(), rowmap.data_ptr(), _stream())
        _lib.check(rc, "fpm_affinity_tiles"); _count()
        ldp = 128 * tB
        P = torch.empty((B * Rmax, ldp), dtype=torch.float32, device=dev)
        rc = L.fpm_gemm_nt_f16x3_tiles(a_hi.data_ptr(), a_lo.data_ptr(), a_inv.data_ptr(), b_hi.data_ptr(),
                                       b_lo.data_ptr(), b_inv.data_ptr(), P.data_ptr(), XA.shape[0], XB.shape[0], K, K, K,
                                       ldp, tab.data_ptr(), meta.data_ptr(), rowmap.data_ptr(), ntile, 0, _stream())
        _lib.check(rc, "fpm_gemm_nt_f16x3_tiles"); _count()
        rc = L.fpm_affinity_finish(P.data_ptr(), ptrA.data_ptr(), ptrB.data_ptr(), out.data_ptr(),
                                   out_t.data_ptr() if want_t else None, B, Rmax, Cmax, ldp, scale, int(raw), _stream())
        _lib.check(rc, "fpm_affinity_finish"); _count()
        return out, out_t
    rc = _lib.lib().fpm_affinity(_chk(XA, "XA"), _chk(XB, "XB"), _chk(coeff, "coeff"),
                                 _chk(ptrA, "ptrA", torch.int64), _chk(ptrB, "ptrB", torch.int64),
                                 None, None, None, None, 0, 0, out.data_ptr(),
                                 out_t.data_ptr() if want_t else None, B, Rmax, Cmax, XA.shape[1], scale, int(raw),
                                 _stream())
    _lib.check(rc, "fpm_affinity"); _count()
    return out, out_t


def affinity_edges(XA: Tensor, XB: Tensor, coeff: Tensor, eptrA: Tensor, eptrB: Tensor, eidxA: Tensor,
                   eidxB: Tensor, Rmax: int, Cmax: int, scale: float = 0.5) -> Tensor:
    B = coeff.shape[0]
    out = torch.empty((B, Rmax, Cmax), dtype=torch.float32, device=XA.device)
    rc = _lib.lib().fpm_affinity(_chk(XA, "XA"), _chk(XB, "XB"), _chk(coeff, "coeff"), None, None,
                                 _chk(eptrA, "eptrA", torch.int64), _chk(eptrB, "eptrB", torch.int64),
                                 _chk(eidxA, "edge_index A", torch.int64), _chk(eidxB, "edge_index B", torch.int64),
                                 eidxA.shape[1], eidxB.shape[1], out.data_ptr(), None, B, Rmax, Cmax, XA.shape[1],
                                 scale, 0, _stream())
    _lib.check(rc, "fpm_affinity"); _count()
    return out


def affinity_edges_factored(XA: Tensor, XB: Tensor, coeff: Tensor, ptrA: Tensor, ptrB: Tensor, eptrA: Tensor,
                            eptrB: Tensor, eidxA: Tensor, eidxB: Tensor, n1max: int, n2max: int, e1max: int,
                            e2max: int, scale: float = 0.5) -> Tensor:
    """Ke from the [n1, n2] node products by linearity (edge feature = x[src] - x[dst]); see gemm_simt.cu."""
    B = coeff.shape[0]
    P, _ = affinity_nodes(XA, XB, coeff, ptrA, ptrB, n1max, n2max, want_t=False, raw=True)
    out = torch.empty((B, e1max, e2max), dtype=torch.float32, device=XA.device)
    rc = _lib.lib().fpm_affinity_edges_factored(P.data_ptr(), _chk(eidxA, "edge_index A", torch.int64),
                                                _chk(eptrA, "eptrA", torch.int64), _chk(ptrA, "ptrA", torch.int64),
                                                _chk(eidxB, "edge_index B", torch.int64),
                                                _chk(eptrB, "eptrB", torch.int64), _chk(ptrB, "ptrB", torch.int64),
                                                eidxA.shape[1], eidxB.shape[1], out.data_ptr(), B, n1max, n2max,
                                                e1max, e2max, scale, _stream())
    _lib.check(rc, "fpm_affinity_edges_factored"); _count()
    return out


# ---------------------------------------------------------------------------------------------------
# association-graph GNN
# ---------------------------------------------------------------------------------------------------
def assoc_in_csr(edges: Tensor, nmax: int, want_col: bool = False):
    """In-neighbour lists of a [B, 2, emax] edge table (a column is an edge only if both its ends are >= 0).
    Returns (in_ptr [B,nmax+1], in_src [B,emax]) and, with ``want_col``, the column id of every list entry."""
    B, _, emax = edges.shape
    in_ptr = torch.empty((B, nmax + 1), dtype=torch.int32, device=edges.device)
    in_src = torch.empty((B, max(emax, 1)), dtype=torch.int32, device=edges.device)
    in_col = torch.empty((B, max(emax, 1)), dtype=torch.int32, device=edges.device) if want_col else None
    rc = _lib.lib().fpm_assoc_in_csr(_chk(edges, "edges", torch.int32), in_ptr.data_ptr(), in_src.data_ptr(),
                                     in_col.data_ptr() if want_col else None, B, nmax, emax, _stream())
    _lib.check(rc, "fpm_assoc_in_csr"); _count()
    return (in_ptr, in_src, in_col) if want_col else (in_ptr, in_src)


class AssocStructure:
    """The association graph of a batch of pairs, kept factorised (csrc/gnn.cu).

    Built from the two per-pair edge tables ``[B, 2, emax]`` (the (G-node, H-node) of every G / H column, -1 where
    the column is all-zero) exactly as the reference's index lists describe it: kron(G2,G1) and kron(H2,H1) drop
    their zero columns independently (gmdataset.py:623-642) and ngm.py:333-342 cuts [idx; diag] to the length of
    K_value.  For complete tables this is the plain Kronecker structure; for the partial permutations of real
    genuine pairs it reproduces the reference's (mis-paired, truncated) lists - see ``assoc_effective_kernel``.
    ``csr1`` / ``csr2`` = (in_ptr, in_src, in_col) of the effective tables, ``ocsr*`` the out-neighbour lists
    (training), ``ndiag`` [B] int64, ``part`` [B,4] int32, ``status`` [1] int32."""

    def __init__(self, edges1: Tensor, edges2: Tensor, eptr1: Tensor, eptr2: Tensor, n1: Tensor, n2: Tensor,
                 n1max: int, n2max: int, with_out: bool = False):
        B, _, e1max = edges1.shape
        e2max = edges2.shape[2]
        dev = edges1.device
        self.eff1 = torch.empty_like(edges1)
        self.eff2 = torch.empty_like(edges2)
        self.ndiag = torch.empty((B,), dtype=torch.int64, device=dev)
        self.part = torch.empty((B, 4), dtype=torch.int32, device=dev)
        self.status = torch.zeros((1,), dtype=torch.int32, device=dev)
        rc = _lib.lib().fpm_assoc_effective(_chk(edges1, "edges1", torch.int32), _chk(edges2, "edges2", torch.int32),
                                            _chk(eptr1, "eptr1", torch.int64), _chk(eptr2, "eptr2", torch.int64),
                                            _chk(n1, "n1", torch.int64), _chk(n2, "n2", torch.int64),
                                            self.eff1.data_ptr(), self.eff2.data_ptr(), self.ndiag.data_ptr(),
                                            self.part.data_ptr(), self.status.data_ptr(), B, e1max, e2max, _stream())
        _lib.check(rc, "fpm_assoc_effective"); _count()
        self.csr1 = assoc_in_csr(self.eff1, n1max, want_col=True)
        self.csr2 = assoc_in_csr(self.eff2, n2max)
        self.ocsr1 = self.ocsr2 = None
        if with_out:
            swap = lambda t: torch.stack((t[:, 1], t[:, 0]), 1).contiguous()
            self.ocsr1 = assoc_in_csr(swap(self.eff1), n1max, want_col=True)
            self.ocsr2 = assoc_in_csr(swap(self.eff2), n2max)
        self.n1max, self.n2max, self.e1max, self.e2max = n1max, n2max, e1max, e2max


def gnn_layer(xprev: Optional[Tensor], mprev_t: Tensor, st: AssocStructure, weights):
    B = mprev_t.shape[0]
    n1max, n2max = st.n1max, st.n2max
    N = n1max * n2max
    dev = mprev_t.device
    xout = torch.empty((B, N, 16), dtype=torch.float32, device=dev)
    score = torch.empty((B, n1max, n2max), dtype=torch.float32, device=dev)
    for i, w in enumerate(weights):
        _chk(w, f"gnn weight {i}")
    wp = _ptr_array(weights)
    rc = _lib.lib().fpm_gnn_layer(_chk(xprev, "xprev"), _chk(mprev_t, "mprev_t"),
                                  st.csr1[0].data_ptr(), st.csr1[1].data_ptr(), st.csr1[2].data_ptr(),
                                  st.csr2[0].data_ptr(), st.csr2[1].data_ptr(), st.ndiag.data_ptr(),
                                  st.part.data_ptr(), wp, xout.data_ptr(), score.data_ptr(), B, n1max, n2max,
                                  st.e1max, st.e2max, 1 if xprev is None else 17, _stream())
    _lib.check(rc, "fpm_gnn_layer"); _count(3)
    return xout, score


def final_classifier(x1: Tensor, sk_t: Tensor, cw: Tensor, cb: Tensor, n1max: int, n2max: int) -> Tensor:
    B = x1.shape[0]
    s = torch.empty((B, n1max, n2max), dtype=torch.float32, device=x1.device)
    rc = _lib.lib().fpm_final_classifier(_chk(x1, "x1"), _chk(sk_t, "sk_t"), _chk(cw, "classifier.weight"),
                                         _chk(cb, "classifier.bias"), s.data_ptr(), B, n1max, n2max, _stream())
    _lib.check(rc, "fpm_final_classifier"); _count()
    return s


# ---------------------------------------------------------------------------------------------------
# Sinkhorn / soft-top-k
# ---------------------------------------------------------------------------------------------------
def sinkhorn_log(s: Tensor, n1: Optional[Tensor], n2: Optional[Tensor], max_iter: int, tau: float,
                 dummy_row: bool, want_t: bool = False):
    B, R, Cc = s.shape
    L = _lib.lib()
    out = torch.empty_like(s)
    out_t = torch.empty((B, Cc, R), dtype=torch.float32, device=s.device) if want_t else None
    wsb = int(L.fpm_sinkhorn_workspace_bytes(B, R, Cc, int(dummy_row)))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=s.device) if wsb else None
    n1 = _i64(n1) if n1 is not None else None
    n2 = _i64(n2) if n2 is not None else None
    rc = L.fpm_sinkhorn_log(_chk(s, "s"), _chk(n1, "nrows", torch.int64), _chk(n2, "ncols", torch.int64),
                            out.data_ptr(), out_t.data_ptr() if want_t else None,
                            ws.data_ptr() if ws is not None else None, B, R, Cc, int(max_iter), float(tau),
                            int(bool(dummy_row)), _stream())
    _lib.check(rc, "fpm_sinkhorn_log"); _count()
    return (out, out_t) if want_t else out


def soft_topk(scores: Tensor, ks: Tensor, n1: Optional[Tensor], n2: Optional[Tensor], max_iter: int,
              tau: float) -> Tensor:
    B, R, Cc = scores.shape
    L = _lib.lib()
    out = torch.empty_like(scores)
    wsb = int(L.fpm_soft_topk_workspace_bytes(B, R, Cc))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=scores.device) if wsb else None
    n1 = _i64(n1) if n1 is not None else None
    n2 = _i64(n2) if n2 is not None else None
    ks = ks.to(torch.float32).contiguous()
    rc = L.fpm_soft_topk(_chk(scores, "scores"), _chk(ks, "ks"), _chk(n1, "nrows", torch.int64),
                         _chk(n2, "ncols", torch.int64), out.data_ptr(), ws.data_ptr() if ws is not None else None,
                         B, R, Cc, int(max_iter), float(tau), _stream())
    _lib.check(rc, "fpm_soft_topk"); _count()
    return out


# ---------------------------------------------------------------------------------------------------
# AFA-U
# ---------------------------------------------------------------------------------------------------
def afau_zero_query_kernel_fits(nr: int, nc: int) -> bool:
    """True when ``afau_attention(..., q_zero=True)`` runs the dedicated zero-query kernel (csrc/afau.cu), which reads
    neither q nor k: they may then be passed as None."""
    if os.environ.get("FPMATCH_AFAU_QZERO") == "0":
        return False
    smem = (((nr * (nc | 1) + 3) & ~3) + 2 * nc * 16) * 4
    return 2 * nr <= 256 and smem <= 72 * 1024


def afau_attention(q: Optional[Tensor], k: Optional[Tensor], v: Tensor, cost: Tensor, transposed_cost: bool,
                   mix1_w: Tensor, mix1_b: Tensor, mix2_w: Tensor, mix2_b: Tensor, q_zero: bool = False) -> Tensor:
    """cost is the ORIGINAL [B, n1, n2] matrix; ``transposed_cost`` makes the kernel read cost^T.  With ``q_zero`` (the
    caller guarantees q == 0) and ``afau_zero_query_kernel_fits`` q and k may be None."""
    B, nc, E = v.shape
    R, Cc = cost.shape[1], cost.shape[2]
    nr = Cc if transposed_cost else R
    if q is not None:
        assert q.shape == (B, nr, E) and k.shape == v.shape
    else:
        assert q_zero and k is None and afau_zero_query_kernel_fits(nr, nc)
    out = torch.empty((B, nr, E), dtype=torch.float32, device=v.device)
    if transposed_cost:
        assert (nr, nc) == (Cc, R)
        cs_b, cs_r, cs_c = R * Cc, 1, Cc
    else:
        assert (nr, nc) == (R, Cc)
        cs_b, cs_r, cs_c = R * Cc, Cc, 1
    rc = _lib.lib().fpm_afau_attention(_chk(q, "q"), _chk(k, "k"), _chk(v, "v"), _chk(cost, "cost"), cs_b, cs_r, cs_c,
                                       _chk(mix1_w, "mix1_weight"), _chk(mix1_b, "mix1_bias"),
                                       _chk(mix2_w, "mix2_weight"), _chk(mix2_b, "mix2_bias"), out.data_ptr(),
                                       B, nr, nc, int(q_zero), _stream())
    _lib.check(rc, "fpm_afau_attention"); _count()
    return out


def add_instnorm(a: Tensor, other: Optional[Tensor], gamma: Tensor, beta: Tensor, want_rowmax: bool = False,
                 eps: float = 1e-5, want_out: bool = True):
    """``want_out=False`` (with ``want_rowmax``): only the per-channel maximum over rows is produced and the normalised
    tensor is never written (returned as None)."""
    B, n, E = a.shape
    if not want_rowmax or n > 112 or E % 4 or os.environ.get("FPMATCH_INSTNORM_TILE") == "0":
        want_out = True
    out = torch.empty_like(a) if want_out else None
    rowmax = torch.empty((B, E), dtype=torch.float32, device=a.device) if want_rowmax else None
    mode = 0 if other is None else (2 if other.dim() == 1 else 1)
    rc = _lib.lib().fpm_add_instnorm(_chk(a, "a"), _chk(other, "other"), mode, _chk(gamma, "norm.weight"),
                                     _chk(beta, "norm.bias"), out.data_ptr() if want_out else None,
                                     rowmax.data_ptr() if want_rowmax else None, B, n, E, float(eps), _stream())
    _lib.check(rc, "fpm_add_instnorm"); _count()
    return (out, rowmax) if want_rowmax else out


def onehot_instnorm(hot: Tensor, nmax: int, vec: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    """``add_instnorm(onehot, vec)`` for the one-hot column embedding ``onehot[b, r, c] = (c == r and r < hot[b])`` of
    ngm.py:396-399 without materialising it ([B, nmax, E] with E = len(vec))."""
    B, E = hot.shape[0], vec.shape[0]
    out = torch.empty((B, nmax, E), dtype=torch.float32, device=vec.device)
    rc = _lib.lib().fpm_onehot_instnorm(_chk(_i64(hot), "hot", torch.int64), _chk(vec, "vec"), _chk(gamma, "norm.weight"),
                                        _chk(beta, "norm.bias"), out.data_ptr(), None, B, nmax, E, float(eps), _stream())
    _lib.check(rc, "fpm_onehot_instnorm"); _count()
    return out


def onehot_proj(W: Tensor, n: Tensor, nmax: int) -> Tensor:
    B = n.shape[0]
    OUT, IN = W.shape
    out = torch.empty((B, nmax, OUT), dtype=torch.float32, device=W.device)
    rc = _lib.lib().fpm_onehot_proj(_chk(W, "W"), _chk(n, "n", torch.int64), out.data_ptr(), B, nmax, OUT, IN,
                                    _stream())
    _lib.check(rc, "fpm_onehot_proj"); _count()
    return out


def k_head(g_row: Tensor, g_col: Tensor, weights, n1: Tensor, n2: Tensor, mean_k: bool = True):
    B, E = g_row.shape
    Hd = weights[0].shape[0]
    ks = torch.empty((B,), dtype=torch.float32, device=g_row.device)
    kscaled = torch.empty((B,), dtype=torch.float32, device=g_row.device)
    for i, w in enumerate(weights):
        _chk(w, f"k-head weight {i}")
    rc = _lib.lib().fpm_k_head(_chk(g_row, "g_row"), _chk(g_col, "g_col"), _ptr_array(weights),
                               _chk(n1, "n1", torch.int64), _chk(n2, "n2", torch.int64), ks.data_ptr(),
                               kscaled.data_ptr(), B, E, Hd, int(mean_k), _stream())
    _lib.check(rc, "fpm_k_head"); _count()
    return ks, kscaled


# ---------------------------------------------------------------------------------------------------
# LAP + greedy
# ---------------------------------------------------------------------------------------------------
def lap_topk(ds: Tensor, n1: Optional[Tensor], n2: Optional[Tensor], ks: Optional[Tensor] = None,
             want_hungarian: bool = True, want_perm: bool = False, want_status: bool = False):
    """(hungarian, perm_mat[, status]).  ``status [B]`` int32 is 1 where the pair's cost matrix was infeasible (NaN /
    inf entries): the outputs of that pair are all-zero and scipy's ``linear_sum_assignment`` would have raised
    ``ValueError`` (utils/hungarian.py:63)."""
    B, R, Cc = ds.shape
    hung = torch.empty_like(ds) if want_hungarian else None
    perm = torch.empty_like(ds) if want_perm else None
    status = torch.empty((B,), dtype=torch.int32, device=ds.device) if want_status else None
    n1 = _i64(n1) if n1 is not None else None
    n2 = _i64(n2) if n2 is not None else None
    if ks is not None:
        ks = ks.to(torch.float32).contiguous()
    rc = _lib.lib().fpm_lap_topk(_chk(ds, "s"), _chk(n1, "n1", torch.int64), _chk(n2, "n2", torch.int64),
                                 _chk(ks, "ks"), hung.data_ptr() if want_hungarian else None,
                                 perm.data_ptr() if want_perm else None,
                                 status.data_ptr() if want_status else None, B, R, Cc, _stream())
    _lib.check(rc, "fpm_lap_topk"); _count()
    return (hung, perm, status) if want_status else (hung, perm)


def greedy_perm(x: Tensor, top_indices: Tensor, ks: Tensor) -> Tensor:
    B, R, Cc = x.shape
    top_indices = _i64(top_indices)
    ks = ks.to(torch.float32).contiguous()
    rc = _lib.lib().fpm_greedy_perm(_chk(x, "x"), _chk(top_indices, "top_indices", torch.int64), _chk(ks, "ks"),
                                    B, R, Cc, top_indices.shape[1], _stream())
    _lib.check(rc, "fpm_greedy_perm"); _count()
    return x


# ---------------------------------------------------------------------------------------------------
# training: vector-Jacobian products (wired into autograd by fpmatch/autograd.py)
# ---------------------------------------------------------------------------------------------------
def node_features_bwd(dX: Tensor, P: Tensor, ns: Tensor, ptr: Tensor, shape1, shape2, ori_size):
    """dX [total, C1+C2] -> gradients of the prepared channels-last maps [B,H1*W1,C1], [B,H2*W2,C2]."""
    B, nmax = P.shape[0], P.shape[1]
    (C1, H1, W1), (C2, H2, W2) = shape1, shape2
    d1 = torch.empty((B, H1 * W1, C1), dtype=torch.float32, device=dX.device)
    d2 = torch.empty((B, H2 * W2, C2), dtype=torch.float32, device=dX.device)
    rc = _lib.lib().fpm_node_features_bwd(_chk(dX, "dX"), _chk(P, "P"), _chk(ns, "ns", torch.int64),
                                          _chk(ptr, "ptr", torch.int64), d1.data_ptr(), d2.data_ptr(), B, nmax,
                                          C1, H1, W1, C2, H2, W2, float(ori_size[0]), float(ori_size[1]), _stream())
    _lib.check(rc, "fpm_node_features_bwd"); _count(2)
    return d1, d2


def fmap_prep_bwd(fmap: Tensor, dy_nhwc: Tensor) -> Tensor:
    B, Cc, Hf, Wf = fmap.shape
    dx = torch.empty_like(fmap)
    rc = _lib.lib().fpm_fmap_prep_bwd(_chk(fmap, "fmap"), _chk(dy_nhwc, "dy"), dx.data_ptr(), B, Cc, Hf, Wf, _stream())
    _lib.check(rc, "fpm_fmap_prep_bwd"); _count()
    return dx


def spline_scatter_bwd(G: Tensor, argmax: Tensor, edge_index: Tensor, pseudo: Tensor, out_ptr: Tensor,
                       out_eid: Tensor, kernel_size: int = 5) -> Tensor:
    total, Cc = G.shape
    NS = kernel_size * kernel_size + 1
    dY = torch.empty((total, NS * Cc), dtype=torch.float32, device=G.device)
    rc = _lib.lib().fpm_spline_scatter_bwd(_chk(G, "G"), _chk(argmax, "argmax", torch.int32),
                                           _chk(edge_index[1], "edge_index[1]", torch.int64), _chk(pseudo, "pseudo"),
                                           _chk(out_ptr, "out_ptr", torch.int32), _chk(out_eid, "out_eid", torch.int32),
                                           dY.data_ptr(), total, Cc, kernel_size, _stream())
    _lib.check(rc, "fpm_spline_scatter_bwd"); _count()
    return dY


class SlabGroups:
    """Column / row compaction of the SplineConv backward for one graph batch (built from the slab plan's per-node
    slab mask; shared by both conv layers).  Slabs read by at least ``T / wide_div`` nodes - and the root slab - form
    the WIDE group (all rows); the other slabs that some edge reads form the NARROW group, restricted to the rows
    ``R`` that read any of them.  Two host reads (slab counts, row list) per graph batch."""

    def __init__(self, plan: "SlabPlan", wide_div: int = 8):
        dev = plan.mask.device
        NS, T = plan.NS, plan.T
        ar = torch.arange(NS - 1, device=dev, dtype=torch.int32)
        cnt = ((plan.mask.view(-1, 1) >> ar) & 1).sum(0).tolist()              # host read 1
        self.wide = [k for k in range(NS - 1) if cnt[k] * wide_div >= T and cnt[k] > 0] + [NS - 1]
        self.narrow = [k for k in range(NS - 1) if 0 < cnt[k] and cnt[k] * wide_div < T]
        colmap = [-(2 ** 31)] * NS
        for g, k in enumerate(self.wide):
            colmap[k] = g
        for g, k in enumerate(self.narrow):
            colmap[k] = -(g + 1)
        self.colmap = torch.tensor(colmap, dtype=torch.int32).to(dev)
        self.rows = None
        self.rowpos = None
        if self.narrow:
            bits = 0
            for k in self.narrow:
                bits |= 1 << k
            self.rows = torch.nonzero(plan.mask & bits).view(-1)                # host read 2 (output size)
            self.rowpos = torch.full((T,), -1, dtype=torch.int32, device=dev)
            self.rowpos[self.rows] = torch.arange(self.rows.numel(), device=dev, dtype=torch.int32)
            if self.rows.numel() == 0:
                self.narrow, self.rows, self.rowpos = [], None, None
        self.NS, self.T = NS, T


def spline_scatter_bwd_compact(G: Tensor, argmax: Tensor, edge_index: Tensor, pseudo: Tensor, out_ptr: Tensor,
                               out_eid: Tensor, groups: SlabGroups, kernel_size: int = 5):
    """(dYd [total, nD*C], dYs [nR, nS*C] or None): the gradient of the slab products, zero blocks left out."""
    total, Cc = G.shape
    nD, nS = len(groups.wide), len(groups.narrow)
    dYd = torch.empty((total, nD * Cc), dtype=torch.float32, device=G.device)
    dYs = torch.empty((groups.rows.numel(), nS * Cc), dtype=torch.float32, device=G.device) if nS else None
    rc = _lib.lib().fpm_spline_scatter_bwd_compact(
        _chk(G, "G"), _chk(argmax, "argmax", torch.int32), _chk(edge_index[1], "edge_index[1]", torch.int64),
        _chk(pseudo, "pseudo"), _chk(out_ptr, "out_ptr", torch.int32), _chk(out_eid, "out_eid", torch.int32),
        _chk(groups.colmap, "colmap", torch.int32), _chk(groups.rowpos, "rowpos", torch.int32), dYd.data_ptr(),
        dYs.data_ptr() if dYs is not None else None, total, Cc, kernel_size, nD, nS, _stream())
    _lib.check(rc, "fpm_spline_scatter_bwd_compact"); _count()
    return dYd, dYs


def transpose_pad(x: Tensor, multiple: int = 8) -> Tensor:
    """[R, C] -> [C, ldo] with ldo = R rounded up to `multiple`, zero padded (K-major GEMM operand)."""
    R, Cc = x.shape
    ldo = (R + multiple - 1) // multiple * multiple
    out = torch.empty((Cc, ldo), dtype=torch.float32, device=x.device)
    rc = _lib.lib().fpm_transpose_f32(_chk(x, "x"), out.data_ptr(), R, Cc, ldo, _stream())
    _lib.check(rc, "fpm_transpose_f32"); _count()
    return out


def bmm_ragged(Mat: Tensor, trans: bool, X: Tensor, ptrX: Tensor, ptrO: Tensor, total_out: int,
               coeff_in: Optional[Tensor] = None, coeff_out: Optional[Tensor] = None) -> Tensor:
    B, Rmax, Cmax = Mat.shape
    D = X.shape[1]
    out = torch.empty((total_out, D), dtype=torch.float32, device=X.device)
    rc = _lib.lib().fpm_bmm_ragged(_chk(Mat, "Mat"), B, Rmax, Cmax, int(trans), _chk(X, "X"),
                                   _chk(ptrX, "ptrX", torch.int64), _chk(ptrO, "ptrO", torch.int64),
                                   _chk(coeff_in, "coeff_in"), _chk(coeff_out, "coeff_out"), out.data_ptr(), D,
                                   _stream())
    _lib.check(rc, "fpm_bmm_ragged"); _count()
    return out


def segment_rowdot(X: Tensor, Y: Tensor, ptr: Tensor) -> Tensor:
    B, D = ptr.numel() - 1, X.shape[1]
    out = torch.empty((B, D), dtype=torch.float32, device=X.device)
    rc = _lib.lib().fpm_segment_rowdot(_chk(X, "X"), _chk(Y, "Y"), _chk(ptr, "ptr", torch.int64), out.data_ptr(),
                                       B, D, _stream())
    _lib.check(rc, "fpm_segment_rowdot"); _count()
    return out


GNN_GRAD_SIZES = lambda cin: [16 * cin, 16, 16 * cin, 16 * cin, 16, 256, 16, 16, 1]


def gnn_layer_bwd(xprev: Optional[Tensor], mprev_t: Tensor, st: AssocStructure, weights, dxout: Tensor,
                  dscore: Tensor):
    """Returns (dxprev [B,N,16] or None, dm [B,n1max,n2max], list of the 9 weight gradients)."""
    assert st.ocsr1 is not None, "AssocStructure was built without the out-neighbour lists (with_out=True)"
    B = mprev_t.shape[0]
    n1max, n2max = st.n1max, st.n2max
    N = n1max * n2max
    dev = mprev_t.device
    cin = 1 if xprev is None else 17
    cp = (cin + 3) // 4 * 4
    dxprev = torch.empty((B, N, 16), dtype=torch.float32, device=dev) if cin > 1 else None
    dm = torch.empty((B, n1max, n2max), dtype=torch.float32, device=dev)
    gagg = torch.empty((B, N, cp), dtype=torch.float32, device=dev)
    sizes = GNN_GRAD_SIZES(cin)
    grads = torch.zeros((sum(sizes),), dtype=torch.float32, device=dev)
    for i, w in enumerate(weights):
        _chk(w, f"gnn weight {i}")
    rc = _lib.lib().fpm_gnn_layer_bwd(
        _chk(xprev, "xprev"), _chk(mprev_t, "mprev_t"),
        st.csr1[0].data_ptr(), st.csr1[1].data_ptr(), st.csr1[2].data_ptr(), st.csr2[0].data_ptr(), st.csr2[1].data_ptr(),
        st.ocsr1[0].data_ptr(), st.ocsr1[1].data_ptr(), st.ocsr1[2].data_ptr(), st.ocsr2[0].data_ptr(),
        st.ocsr2[1].data_ptr(), st.ndiag.data_ptr(), st.part.data_ptr(), _ptr_array(weights),
        _chk(dxout, "dxout"), _chk(dscore, "dscore"), dxprev.data_ptr() if cin > 1 else None, dm.data_ptr(),
        gagg.data_ptr(), grads.data_ptr(), B, n1max, n2max, st.e1max, st.e2max, cin, _stream())
    _lib.check(rc, "fpm_gnn_layer_bwd"); _count(4)
    return dxprev, dm, list(torch.split(grads, sizes))


def sinkhorn_log_bwd(s: Tensor, n1: Optional[Tensor], n2: Optional[Tensor], gout: Tensor, max_iter: int,
                     tau: float, dummy_row: bool) -> Tensor:
    B, R, Cc = s.shape
    L = _lib.lib()
    gs = torch.empty_like(s)
    wsb = int(L.fpm_sinkhorn_bwd_workspace_bytes(B, R, Cc, int(max_iter)))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=s.device) if wsb else None
    n1 = _i64(n1) if n1 is not None else None
    n2 = _i64(n2) if n2 is not None else None
    rc = L.fpm_sinkhorn_log_bwd(_chk(s, "s"), _chk(n1, "nrows", torch.int64), _chk(n2, "ncols", torch.int64),
                                _chk(gout, "gout"), gs.data_ptr(), ws.data_ptr() if ws is not None else None,
                                B, R, Cc, int(max_iter), float(tau), int(bool(dummy_row)), _stream())
    _lib.check(rc, "fpm_sinkhorn_log_bwd"); _count()
    return gs


def soft_topk_bwd(scores: Tensor, ks: Tensor, n1: Optional[Tensor], n2: Optional[Tensor], gout: Tensor,
                  max_iter: int, tau: float) -> Tensor:
    B, R, Cc = scores.shape
    L = _lib.lib()
    gs = torch.empty_like(scores)
    wsb = int(L.fpm_soft_topk_bwd_workspace_bytes(B, R, Cc))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=scores.device) if wsb else None
    n1 = _i64(n1) if n1 is not None else None
    n2 = _i64(n2) if n2 is not None else None
    ks = ks.to(torch.float32).contiguous()
    rc = L.fpm_soft_topk_bwd(_chk(scores, "scores"), _chk(ks, "ks"), _chk(n1, "nrows", torch.int64),
                             _chk(n2, "ncols", torch.int64), _chk(gout, "gout"), gs.data_ptr(),
                             ws.data_ptr() if ws is not None else None, B, R, Cc, int(max_iter), float(tau), _stream())
    _lib.check(rc, "fpm_soft_topk_bwd"); _count()
    return gs


# ---------------------------------------------------------------------------------------------------
# batched CSR / CSC products and the dense FGM affinity (src.sparse_torch, src.sparse, RebuildFGM)
# ---------------------------------------------------------------------------------------------------
def _csx(m, name):
    return (_chk(m.indices, name + ".indices", torch.int64), _chk(m.indptr, name + ".indptr", torch.int64),
            _chk(m.data, name + ".data"))


def csr_dot_diag(indices: Tensor, indptr: Tensor, data: Tensor, diag: Tensor, shape) -> Tensor:
    B, h, w = shape
    out = torch.empty_like(data)
    rc = _lib.lib().fpm_csr_dot_diag(_chk(indices, "indices", torch.int64), _chk(indptr, "indptr", torch.int64),
                                     _chk(data, "data"), _chk(diag.contiguous(), "diag"), out.data_ptr(), B, h, w,
                                     _stream())
    _lib.check(rc, "fpm_csr_dot_diag"); _count()
    return out


def csr_dot_csc_dense(t1, t2) -> Tensor:
    B, h, _ = t1.shape
    w = t2.shape[2]
    out = torch.empty((B, h, w), dtype=torch.float32, device=t1.device)
    rc = _lib.lib().fpm_csr_dot_csc_dense(*_csx(t1, "t1"), *_csx(t2, "t2"), out.data_ptr(), B, h, w, _stream())
    _lib.check(rc, "fpm_csr_dot_csc_dense"); _count()
    return out


def dense_dot_csc_dense(d: Tensor, t2) -> Tensor:
    B, h, k = d.shape
    w = t2.shape[2]
    out = torch.empty((B, h, w), dtype=torch.float32, device=d.device)
    rc = _lib.lib().fpm_dense_dot_csc_dense(_chk(d.contiguous(), "dense"), *_csx(t2, "t2"), out.data_ptr(), B, h, k, w,
                                            _stream())
    _lib.check(rc, "fpm_dense_dot_csc_dense"); _count()
    return out


def bilinear_diag(t1, d: Tensor, t3) -> Tensor:
    B, x, f = t1.shape
    out = torch.empty((B, x), dtype=torch.float32, device=d.device)
    i1, p1, d1 = _csx(t1, "t1")
    i3, p3, d3 = _csx(t3, "t3")
    rc = _lib.lib().fpm_bilinear_diag(i1, p1, d1, _chk(d.contiguous(), "t2"), i3, p3, d3, out.data_ptr(), B, x, f,
                                      _stream())
    _lib.check(rc, "fpm_bilinear_diag"); _count()
    return out


def fgm_rebuild(KroGt, KroHt, ke_vec: Tensor, kp_vec: Tensor) -> Tensor:
    """K [B,N,N] from the transposed Kronecker factors (CSR [B,E,N], CSC [B,N,E]) and vec(Ke), vec(Kp)."""
    B, E, N = KroGt.shape
    K = torch.empty((B, N, N), dtype=torch.float32, device=ke_vec.device)
    rc = _lib.lib().fpm_fgm_rebuild(*_csx(KroGt, "KroGt"), *_csx(KroHt, "KroHt"), _chk(ke_vec, "vec(Ke)"),
                                    _chk(kp_vec, "vec(Kp)"), K.data_ptr(), B, E, N, _stream())
    _lib.check(rc, "fpm_fgm_rebuild"); _count(2)
    return K


# ---------------------------------------------------------------------------------------------------
# loss / metrics (src.loss_func.PermutationLoss, src.evaluation_metric.matching_recall)
# ---------------------------------------------------------------------------------------------------
def permutation_loss_pairs(pred: Tensor, gt: Tensor, n1: Tensor, n2: Tensor) -> Tensor:
    """Per-pair summed BCE over the valid block, [B]."""
    B, R, Cc = pred.shape
    out = torch.empty((B,), dtype=torch.float32, device=pred.device)
    rc = _lib.lib().fpm_permutation_loss(_chk(pred, "pred"), _chk(gt, "gt"), _chk(n1, "n1", torch.int64),
                                         _chk(n2, "n2", torch.int64), out.data_ptr(), B, R, Cc, _stream())
    _lib.check(rc, "fpm_permutation_loss"); _count()
    return out


def permutation_loss_bwd(pred: Tensor, gt: Tensor, n1: Tensor, n2: Tensor, gscale: Tensor) -> Tensor:
    B, R, Cc = pred.shape
    grad = torch.empty_like(pred)
    rc = _lib.lib().fpm_permutation_loss_bwd(_chk(pred, "pred"), _chk(gt, "gt"), _chk(n1, "n1", torch.int64),
                                             _chk(n2, "n2", torch.int64), _chk(gscale, "gscale"), grad.data_ptr(),
                                             B, R, Cc, _stream())
    _lib.check(rc, "fpm_permutation_loss_bwd"); _count()
    return grad


def matching_stats(pred: Tensor, gt: Tensor, ns: Tensor) -> Tensor:
    """[B, 3]: sum(pred * gt), sum(gt), sum(pred) over rows < ns[b]."""
    B, R, Cc = pred.shape
    out = torch.empty((B, 3), dtype=torch.float32, device=pred.device)
    rc = _lib.lib().fpm_matching_stats(_chk(pred, "pred"), _chk(gt, "gt"), _chk(ns, "ns", torch.int64), out.data_ptr(),
                                       B, R, Cc, _stream())
    _lib.check(rc, "fpm_matching_stats"); _count()
    return out


def head_losses(logits: Tensor, label: Optional[Tensor], ks: Optional[Tensor], gt_perm: Tensor, n1: Tensor, n2: Tensor,
                k_factor: float):
    """The scalar tail of ``Net.forward`` in eval mode (ngm.py:456-469) in one launch.  Returns ``(cls_prob [B],
    cls_loss, ks_loss, ks_error)``, the three losses as 0-dim views of one tensor; ``label`` / ``ks`` None -> the
    corresponding losses are 0."""
    B, R, Cc = gt_perm.shape
    dev = logits.device
    ws = torch.zeros((3 * B + 4,), dtype=torch.float32, device=dev)       # per-pair terms + the ticket counter
    cls_prob = torch.empty((B,), dtype=torch.float32, device=dev)
    scalars = torch.empty((3,), dtype=torch.float32, device=dev)
    rc = _lib.lib().fpm_head_losses(_chk(logits, "cls_logits"), _chk(label, "label"), _chk(ks, "ks"),
                                    _chk(gt_perm, "gt_perm_mat"), _chk(_i64(n1), "n1", torch.int64),
                                    _chk(_i64(n2), "n2", torch.int64), float(k_factor), cls_prob.data_ptr(),
                                    ws.data_ptr(), scalars.data_ptr(), B, R, Cc, _stream())
    _lib.check(rc, "fpm_head_losses"); _count()
    return cls_prob, scalars[0], scalars[1], scalars[2]


def fgm_aggregate(A: Tensor, W: Tensor, X: Tensor, norm: bool = True, trans: bool = False,
                  inv: Optional[Tensor] = None):
    """Dense NGM-v1 aggregation (GNNLayer.forward, gnn.py:54-68).  Forward: returns (out [B,N,F], inv [B,N] or None);
    ``trans=True`` evaluates the transposed product with the given row scales (backward of the forward)."""
    B, N, F_ = X.shape
    fe = W.shape[-1]
    out = torch.empty((B, N, F_), dtype=torch.float32, device=X.device)
    if not trans:
        inv = torch.empty((B, N), dtype=torch.float32, device=X.device) if norm else None
    rc = _lib.lib().fpm_fgm_aggregate(_chk(A, "A"), _chk(W, "W"), _chk(X, "X"), _chk(inv, "inv"), out.data_ptr(), B, N,
                                      F_, fe, int(norm), int(trans), _stream())
    _lib.check(rc, "fpm_fgm_aggregate"); _count()
    return (out, inv) if not trans else out


def fgm_aggregate_dw(A: Tensor, inv: Optional[Tensor], dx2: Tensor, x1: Tensor, fe: int) -> Tensor:
    B, N, F_ = x1.shape
    dW = torch.empty((B, N, N, fe), dtype=torch.float32, device=x1.device)
    rc = _lib.lib().fpm_fgm_aggregate_dw(_chk(A, "A"), _chk(inv, "inv"), _chk(dx2, "dx2"), _chk(x1, "x1"),
                                         dW.data_ptr(), B, N, F_, fe, _stream())
    _lib.check(rc, "fpm_fgm_aggregate_dw"); _count()
    return dW


def match_classifier(s: Tensor, perm: Optional[Tensor], w1: Tensor, b1: Tensor, bn1: Sequence[Tensor], w2: Tensor,
                     b2: Tensor, bn2: Sequence[Tensor], fcw: Tensor, fcb: Tensor, eps: float = 1e-5) -> Tensor:
    """Logits [B] of the genuine / imposter CNN on ``s * perm`` (MatchClassifier.forward in eval mode, ngm.py:75-106).
    ``bn1`` / ``bn2`` = (weight, bias, running_mean, running_var) of the two BatchNorm2d layers."""
    B, H, W = s.shape
    if tuple(w1.shape) != (16, 1, 3, 3) or tuple(w2.shape) != (32, 16, 3, 3) or fcw.numel() != 32:
        raise RuntimeError("fpmatch: match_classifier is built for the reference's (16, 32) channel plan")
    L = _lib.lib()
    ws = torch.empty(L.fpm_match_classifier_workspace_floats(B, H, W), dtype=torch.float32, device=s.device)
    out = torch.empty(B, dtype=torch.float32, device=s.device)
    for t in list(bn1) + list(bn2):
        _chk(t, "batch-norm tensor")
    rc = L.fpm_match_classifier(_chk(s, "s"), _chk(perm, "perm"), _chk(w1, "w1"), _chk(b1, "b1"), _ptr_array(bn1),
                                _chk(w2, "w2"), _chk(b2, "b2"), _ptr_array(bn2), _chk(fcw, "fcw"), _chk(fcb, "fcb"),
                                float(eps), ws.data_ptr(), out.data_ptr(), B, H, W, _stream())
    _lib.check(rc, "fpm_match_classifier"); _count(3)
    return out


# ---------------------------------------------------------------------------------------------------
# AFA-U backward (k-branch training, stages 2-5)
# ---------------------------------------------------------------------------------------------------
def _cost_strides(cost: Tensor, nr: int, nc: int, transposed_cost: bool):
    R, Cc = cost.shape[1], cost.shape[2]
    if transposed_cost:
        assert (nr, nc) == (Cc, R)
        return R * Cc, 1, Cc
    assert (nr, nc) == (R, Cc)
    return R * Cc, Cc, 1


def afau_attention_bwd(q: Tensor, k: Tensor, v: Tensor, cost: Tensor, transposed_cost: bool, mix1_w: Tensor,
                       mix1_b: Tensor, mix2_w: Tensor, mix2_b: Tensor, out: Tensor, dout: Tensor):
    """Returns (dq, dk, dv, dmix1_w [16,2,16], dmix1_b [16,16], dmix2_w [16,16,1], dmix2_b [16,1])."""
    B, nr, E = q.shape
    nc = k.shape[1]
    dq = torch.empty_like(q)
    dk = torch.zeros_like(k); dv = torch.zeros_like(v)
    dmix = torch.zeros((16, 65), dtype=torch.float32, device=q.device)
    cs_b, cs_r, cs_c = _cost_strides(cost, nr, nc, transposed_cost)
    rc = _lib.lib().fpm_afau_attention_bwd(_chk(q, "q"), _chk(k, "k"), _chk(v, "v"), _chk(cost, "cost"), cs_b, cs_r, cs_c,
                                           _chk(mix1_w, "mix1_weight"), _chk(mix1_b, "mix1_bias"),
                                           _chk(mix2_w, "mix2_weight"), _chk(mix2_b, "mix2_bias"), _chk(out, "out"),
                                           _chk(dout, "dout"), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                           dmix.data_ptr(), B, nr, nc, _stream())
    _lib.check(rc, "fpm_afau_attention_bwd"); _count()
    dm1w = torch.stack((dmix[:, 0:16], dmix[:, 16:32]), dim=1)
    return dq, dk, dv, dm1w, dmix[:, 32:48].clone(), dmix[:, 48:64].reshape(16, 16, 1).clone(), dmix[:, 64:65].clone()


def add_instnorm_bwd(a: Tensor, other: Optional[Tensor], gamma: Tensor, dy: Optional[Tensor],
                     drowmax: Optional[Tensor], eps: float = 1e-5):
    """Returns (dx [B,n,E], dgamma [E], dbeta [E], dvec [E] or None)."""
    B, n, E = a.shape
    mode = 0 if other is None else (2 if other.dim() == 1 else 1)
    dx = torch.empty_like(a)
    dgamma = torch.zeros((E,), dtype=torch.float32, device=a.device)
    dbeta = torch.zeros((E,), dtype=torch.float32, device=a.device)
    dvec = torch.zeros((E,), dtype=torch.float32, device=a.device) if mode == 2 else None
    rc = _lib.lib().fpm_add_instnorm_bwd(_chk(a, "a"), _chk(other, "other"), mode, _chk(gamma, "norm.weight"),
                                         _chk(dy, "dy"), _chk(drowmax, "drowmax"), dx.data_ptr(), dgamma.data_ptr(),
                                         dbeta.data_ptr(), dvec.data_ptr() if dvec is not None else None, B, n, E,
                                         float(eps), _stream())
    _lib.check(rc, "fpm_add_instnorm_bwd"); _count()
    return dx, dgamma, dbeta, dvec


# ---------------------------------------------------------------------------------------------------
# device guard: every public wrapper (and the constructors of the plan / structure classes) runs on the device of its
# tensors - see on_tensor_device
# ---------------------------------------------------------------------------------------------------
def _install_device_guards():
    import types
    g = globals()
    skip = {"set_gemm_mode", "gemm_mode", "set_affinity_tc", "set_gemm_pair", "set_gemm_max_clusters", "gemm_pair_enabled", "set_slab_plan", "slab_plan_enabled",
            "launch_count", "gemm_profile_start", "gemm_profile_stop", "on_tensor_device"}
    for name, obj in list(g.items()):
        if name.startswith("_") or name in skip:
            continue
        if isinstance(obj, types.FunctionType) and obj.__module__ == __name__:
            g[name] = on_tensor_device(obj)
    for cls in (SlabPlan, SlabGroups, AssocStructure):
        cls.__init__ = on_tensor_device(cls.__init__)


_install_device_guards()
