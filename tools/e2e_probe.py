#!/usr/bin/env python
"""Where does the end-to-end step (pinned host batch -> device -> Net.forward -> pinned host results) spend its time?
Wall-clock ms per step of fpmatch.prefetch.MatchingPipeline at bench.py's shape under a few switches."""
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import synth
from fpmatch.prefetch import CudaPrefetcher, HostResultRing, MatchingPipeline
from src.model.ngm import Net

dev = torch.device("cuda")
torch.manual_seed(0)
net = Net(regression=True).eval().to(dev)
host = synth.make_batch(256, 100, seed=1234, with_dense_gh=False)


def pin(v):
    if isinstance(v, torch.Tensor):
        return v.pin_memory()
    if isinstance(v, (list, tuple)):
        return type(v)(pin(x) for x in v)
    if hasattr(v, "edge_index"):
        for k in ("x", "edge_index", "edge_attr", "ptr", "eptr"):
            setattr(v, k, getattr(v, k).pin_memory())
        return v
    return v


host = {k: pin(v) for k, v in host.items()}
keys = ("ds_mat", "perm_mat", "k_prob", "cls_prob")


class Identity:
    def __call__(self, d):
        z = torch.zeros(4, device=dev)
        return {k: z for k in keys}


def run(model, inflight, n=10, batch=host, extra=None):
    feeder = CudaPrefetcher([], device=dev)
    ring = HostResultRing(device=dev)
    def once(k):
        c = 0
        for _ in MatchingPipeline(model, [batch] * k, keys=keys, device=dev, inflight=inflight, feeder=feeder, ring=ring,
                                  extra=extra):
            c += 1
        assert c == k
    once(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    once(n)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


res = {}
res["full_inflight1"] = run(net, 1)
res["full_inflight2"] = run(net, 2)
res["copies_only_inflight1"] = run(Identity(), 1)
net.track_lap_status = False
res["full_inflight1_no_lap_status"] = run(net, 1)
net.track_lap_status = True
net.graph_fork = False
res["full_inflight1_no_graph_fork"] = run(net, 1)
net.graph_fork = True
resident = synth.batch_to(synth.clone_batch(host), dev)
nomaps = {k: v for k, v in host.items() if k != "fmaps"}
res["resident_maps_inflight1"] = run(net, 1, batch=nomaps, extra={"fmaps": resident["fmaps"]})
# plain loop: synchronous reference-style data_to_cuda + forward + .cpu()
from utils.data_to_cuda import data_to_cuda
from fpmatch.prefetch import _shallow
def plain(n=6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        d = data_to_cuda(_shallow(host), device=dev)
        with torch.no_grad():
            o = net(d)
        _ = [o[k].cpu() for k in keys]
    return (time.perf_counter() - t0) / n * 1e3
plain(2)
res["synchronous_reference_style_loop"] = plain()
# host cost of ENQUEUEING one forward (python + ctypes + allocator + launches): a 2-pair batch keeps the GPU far ahead of
# the host, so the wall clock per step is the host's alone; it does not depend on the batch size
tiny = synth.batch_to(synth.make_batch(2, 100, seed=5, with_dense_gh=False), dev)
def enqueue(n=30):
    for _ in range(5):
        with torch.no_grad():
            net(synth.clone_batch(tiny))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        with torch.no_grad():
            net(synth.clone_batch(tiny))
    t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3
res["host_enqueue_ms_per_forward"] = enqueue()
print(json.dumps(res, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "e2e_probe.json").write_text(json.dumps(res))
