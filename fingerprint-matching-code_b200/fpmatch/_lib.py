"""ctypes binding of ``libfpmatch_b200.so`` (the C ABI of ``include/fpmatch.h``).

There is no fallback: if the shared library is missing it is built in-tree with nvcc, and if that is
impossible the import of any op raises.  Calls fail loudly (``RuntimeError`` with the library's message)
on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

from . import build as _build

_LIB = None

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_LL = C.c_longlong
_D = C.c_double

# name -> (restype, argtypes); kept in the order of include/fpmatch.h
SIGNATURES = {
    "fpm_abi_version": (_I, []),
    "fpm_device_ok": (_I, []),
    "fpm_last_error": (C.c_char_p, []),
    "fpm_set_error": (None, [C.c_char_p]),
    "fpm_feature_align": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _I, _P]),
    "fpm_fmap_prep": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "fpm_global_max": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "fpm_node_features": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "fpm_affinity_coeff": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_gemm_nt_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fpm_tf32_split": (_I, [_P, _P, _P, _LL, _P]),
    "fpm_gemm_nt_tc": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fpm_gemm_set_trace": (_I, [_P, _I]),
    "fpm_gemm_set_pair": (_I, [_I]),
    "fpm_gemm_set_max_clusters": (_I, [_I]),
    "fpm_f16_split_rows": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "fpm_gemm_nt_f16x3": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fpm_gemm_nt_f16x3_tiles": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _LL, _I, _P]),
    "fpm_spline_plan": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "fpm_spline_gather_rows": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_csr_by_dst": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_spline_gather_max": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_affinity": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "fpm_f16_split_rows_scaled": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _P]),
    "fpm_affinity_tiles": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "fpm_affinity_finish": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "fpm_affinity_edges_factored": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _F, _P]),
    "fpm_assoc_effective": (_I, [_P] * 11 + [_I, _I, _I, _P]),
    "fpm_assoc_in_csr": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_gnn_layer": (_I, [_P] * 12 + [_I] * 6 + [_P]),
    "fpm_final_classifier": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_sinkhorn_workspace_bytes": (_LL, [_I, _I, _I, _I]),
    "fpm_sinkhorn_log": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "fpm_soft_topk_workspace_bytes": (_LL, [_I, _I, _I]),
    "fpm_soft_topk": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "fpm_afau_attention": (_I, [_P, _P, _P, _P, _LL, _LL, _LL, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_add_instnorm": (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "fpm_onehot_instnorm": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "fpm_onehot_proj": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_k_head": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_lap_topk": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_greedy_perm": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    # loss / metrics
    "fpm_permutation_loss": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_permutation_loss_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_matching_stats": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_head_losses": (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _P]),
    # batched CSR / CSC containers + dense FGM affinity
    "fpm_csr_dot_diag": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_csr_dot_csc_dense": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_dense_dot_csc_dense": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_bilinear_diag": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_fgm_rebuild": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    # training (backward) entry points
    "fpm_node_features_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "fpm_fmap_prep_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_spline_scatter_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "fpm_spline_scatter_bwd_compact": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "fpm_transpose_f32": (_I, [_P, _P, _I, _I, _I, _P]),
    "fpm_bmm_ragged": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P]),
    "fpm_segment_rowdot": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "fpm_gnn_layer_bwd": (_I, [_P] * 21 + [_I] * 6 + [_P]),
    "fpm_afau_attention_bwd": (_I, [_P, _P, _P, _P, _LL, _LL, _LL] + [_P] * 10 + [_I, _I, _I, _P]),
    "fpm_add_instnorm_bwd": (_I, [_P, _P, _I] + [_P] * 7 + [_I, _I, _I, _F, _P]),
    "fpm_sinkhorn_bwd_workspace_bytes": (_LL, [_I, _I, _I, _I]),
    "fpm_sinkhorn_log_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "fpm_soft_topk_bwd_workspace_bytes": (_LL, [_I, _I, _I]),
    "fpm_soft_topk_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "fpm_match_classifier_workspace_floats": (_LL, [_I, _I, _I]),
    "fpm_match_classifier": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _P]),
    "fpm_fgm_aggregate": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "fpm_fgm_aggregate_dw": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    # keypoint-graph construction
    "fpm_graph_adjacency": (_I, [_P, _P, _P, _I, _I, _I, _D, _P]),
    "fpm_graph_row_counts": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "fpm_graph_edges": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _LL, _I, _I, _D, _P]),
    "fpm_graph_permute": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_graph_incidence": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "fpm_graph_kron_index": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
}


def header_symbols() -> list:
    """Function names declared in include/fpmatch.h (used by the ABI test)."""
    hdr = (Path(__file__).resolve().parents[2] / "include" / "fpmatch.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fpm_[a-z0-9_]+)\s*\(", hdr)))


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = _build.LIB_PATH
        if not path.exists() or _build.stale():     # sources edited since the library was linked: rebuild, never
            _build.build()                          # run an out-of-date binary silently
        handle = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError = the library is stale: fail loudly
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().fpm_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
