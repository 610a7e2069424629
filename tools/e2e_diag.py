"""Where does an end-to-end step spend its time?  Per step: CPU time to stage the next batch, device time of the H2D
copy (copy stream), of the forward (main stream) and wall time; plus the plain pinned-copy link rate."""
import sys, time, json, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import synth
from fpmatch.prefetch import CudaPrefetcher, HostResultRing
from src.model.ngm import Net

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(regression=True).eval().to(dev)
host = synth.make_batch(256, 100, seed=1234, with_kron=False, with_dense_gh=False)
def pin(v):
    if isinstance(v, torch.Tensor): return v.pin_memory()
    if isinstance(v, (list, tuple)): return type(v)(pin(x) for x in v)
    if hasattr(v, "edge_index"):
        for k in ("x", "edge_index", "edge_attr", "ptr", "eptr"): setattr(v, k, getattr(v, k).pin_memory())
        return v
    return v
host = {k: pin(v) for k, v in host.items()}
print("pinned:", host["fmaps"][0][0].is_pinned(), "cpus", os.cpu_count(), "threads", torch.get_num_threads())
probe = host["fmaps"][0][0]; dst = torch.empty_like(probe, device=dev)
for _ in range(2): dst.copy_(probe, non_blocking=True)
torch.cuda.synchronize()
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0.record()
for _ in range(4): dst.copy_(probe, non_blocking=True)
c1.record(); torch.cuda.synchronize()
print("plain pinned H2D GB/s:", 4 * probe.numel() * 4 / (c0.elapsed_time(c1) / 1e3) / 1e9)

feeder = CudaPrefetcher([], device=dev); ring = HostResultRing(device=dev)
keys = ("ds_mat", "perm_mat", "k_prob", "cls_prob")
def run(n, log):
    feeder.batches = [host] * n
    t_prev = time.perf_counter()
    for d in feeder:
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad(): o = net(d)
        e1.record()
        t1 = time.perf_counter()
        ring.push([o[k] for k in keys])
        t2 = time.perf_counter()
        log.append({"stage_ms": (t0 - t_prev) * 1e3, "enqueue_ms": (t1 - t0) * 1e3, "push_wait_ms": (t2 - t1) * 1e3, "ev": (e0, e1)})
        t_prev = t2
    ring.flush()
run(3, [])
torch.cuda.synchronize()
log = []
t0 = time.perf_counter(); run(10, log); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 10 * 1e3
for r in log: r["fwd_ms"] = r.pop("ev")[0].elapsed_time(r["ev"][1]) if False else None
print("wall ms/step", wall)
for i, r in enumerate(log): print(i, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k != "fwd_ms"})
