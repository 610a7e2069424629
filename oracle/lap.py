"""ctypes wrapper of oracle/lap_ref.c (test infrastructure; see oracle/__init__.py)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "_build" / "liblap_ref.so"
        if not so.exists() or so.stat().st_mtime < (_HERE / "lap_ref.c").stat().st_mtime:
            subprocess.run(["make", "-s", "-C", str(_HERE)], check=True)
        _LIB = C.CDLL(str(so))
        _LIB.lap_ref_solve.restype = C.c_int
        _LIB.lap_ref_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return _LIB


def solve(cost: np.ndarray):
    """(row_ind, col_ind) of the minimum-cost assignment of a float32 matrix, scipy's exact result."""
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    nr, nc = cost.shape
    k = min(nr, nc)
    rows = np.empty(k, dtype=np.int64); cols = np.empty(k, dtype=np.int64)
    rc = _lib().lap_ref_solve(cost.ctypes.data, nr, nc, rows.ctypes.data, cols.ctypes.data)
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    return rows, cols
