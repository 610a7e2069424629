"""In-tree build of ``libfpmatch_b200.so`` (sm_100a only) with plain nvcc.

The library has no torch / python dependency: it is the C-ABI drop-in boundary declared in
``include/fpmatch.h``.  ``python -m fpmatch.build`` (or ``__graft_entry__.build()``) compiles every
``csrc/*.cu`` to an object and links them; objects are rebuilt only when their source (or a header)
is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent.parent          # fingerprint-matching-code_b200/
CSRC = PKG_DIR / "csrc"
BUILD = CSRC / "build"
LIB_PATH = PKG_DIR / "fpmatch" / "libfpmatch_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the fpmatch CUDA library cannot be built")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


MANIFEST = LIB_PATH.with_suffix(".manifest")


def _inputs():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted((PKG_DIR.parent / "include").glob("*.h"))


def source_digest() -> str:
    """Content hash of everything the library is compiled from (robust against copied trees with fresh mtimes)."""
    import hashlib
    h = hashlib.sha256()
    for f in _inputs():
        h.update(f.name.encode()); h.update(f.read_bytes())
    return h.hexdigest()


def stale() -> bool:
    """True when the sources differ from the ones the existing library was built from."""
    try:
        return MANIFEST.read_text().strip() != source_digest()
    except OSError:
        return not LIB_PATH.exists()        # a library without a manifest (older build) is taken as it is


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    BUILD.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted((PKG_DIR.parent / "include").glob("*.h"))
    objs, jobs = [], []
    for src in sources:
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-I", str(PKG_DIR.parent / "include"),
               "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr[-4000:]}")
        if verbose:
            print(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if force or jobs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", str(LIB_PATH), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    MANIFEST.write_text(source_digest() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
