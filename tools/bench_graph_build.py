"""Time the batched GPU graph construction (N1) next to scipy + numpy, the reference's host route.

    python tools/bench_graph_build.py [--out gpurun_out/graph_build_bench.json]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "fingerprint-matching-code_b200"))
sys.path.insert(0, str(ROOT))

from fpmatch import graph_build as gb            # noqa: E402
from oracle import graphs as og                  # noqa: E402  (CPU baseline leg only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rows = []
    for n, B in ((100, 512), (400, 64)):
        rng = np.random.RandomState(n)
        pts = np.stack([rng.uniform(0, 320, (B, n)), rng.uniform(0, 240, (B, n))], -1)
        P = torch.tensor(pts, device="cuda")
        ns = torch.full((B,), n, device="cuda", dtype=torch.int64)
        for _ in range(3):
            gb.build_graph_batch(P, ns, "tri")
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 10
        ev[0].record()
        for _ in range(reps):
            built = gb.build_graph_batch(P, ns, "tri")
        ev[1].record()
        torch.cuda.synchronize()
        gpu_ms = ev[0].elapsed_time(ev[1]) / reps
        ev[0].record()
        for _ in range(reps):
            gb.graph_adjacency(P, ns, "tri")
        ev[1].record()
        torch.cuda.synchronize()
        adj_ms = ev[0].elapsed_time(ev[1]) / reps
        A = gb.graph_adjacency(P, ns, "tri")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            gb.graph_edges(A, P, ns)
        torch.cuda.synchronize()
        edges_wall_ms = (time.perf_counter() - t0) / reps * 1e3
        m = min(B, 16)
        og.delaunay_adjacency_ref(pts[0])                                  # scipy import / first-call cost
        t0 = time.perf_counter()
        for b in range(m):
            A = og.delaunay_adjacency_ref(pts[b])
            og.build_graphs(pts[b], n, stg="tri", ref=True)
            og.pyg_graph(A, pts[b])
        cpu_ms = (time.perf_counter() - t0) / m * B * 1e3
        rows.append({"n": n, "graphs": B, "edges_total": int(built.graph.edge_index.shape[1]),
                     "gpu_ms_per_batch": gpu_ms, "gpu_adjacency_ms": adj_ms, "edges_wall_ms": edges_wall_ms, "graphs_per_s": B / gpu_ms * 1e3,
                     "cpu_scipy_numpy_ms_per_batch_1core": cpu_ms, "cpu_sample_graphs": m})
        print(json.dumps(rows[-1]))
    if args.out:
        Path(args.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
