#!/usr/bin/env python
"""Does the SplineConv gather / max kernel run beside the persistent slab GEMM of the OTHER graph (two streams)?
Times, at bench.py's shape (256 pairs x 100 keypoints): the slab GEMM alone, the gather alone, and both launched
together on two streams in either order.  max(t_gemm, t_gather) = they overlap; the sum = they do not."""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import ops, synth
from src.model.ngm import Net

dev = torch.device("cuda")
torch.manual_seed(0)
net = Net(regression=True).eval().to(dev)
data = synth.batch_to(synth.make_batch(256, 100, seed=1234, with_dense_gh=False), dev)
g = data["pyg_graphs"][0]
conv = net.message_pass_node_features.mp_network.convs[0]
total = g.x.shape[0]
x = torch.randn(total, 768, device=dev) * 0.05
ei, ea = g.edge_index.contiguous(), g.edge_attr.contiguous()
emax = int((g.eptr[1:] - g.eptr[:-1]).max())
csr = ops.csr_by_dst(ei, g.ptr, g.eptr, total, emax)
plan = ops.SlabPlan(ei, ea, total, 768, 5)
packed = conv.packed_weight()
bias = conv.bias.detach().contiguous()
Y = ops.spline_slab_gemm(x, packed, plan)
Y2 = Y.clone()
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()


def gemm():
    return ops.spline_slab_gemm(x, packed, plan)


def gather():
    return ops.spline_gather_max(Y2, None, ei, ea, csr[0], csr[1], bias, 0, 5)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def both(first_gemm=True):
    cur = torch.cuda.current_stream()
    sA.wait_stream(cur); sB.wait_stream(cur)
    if first_gemm:
        with torch.cuda.stream(sA):
            gemm()
        with torch.cuda.stream(sB):
            gather()
    else:
        with torch.cuda.stream(sB):
            gather()
        with torch.cuda.stream(sA):
            gemm()
    cur.wait_stream(sA); cur.wait_stream(sB)


def both_detail():
    """Event times of one concurrent launch: when does each kernel start / finish relative to the common fork?"""
    cur = torch.cuda.current_stream()
    out = []
    for _ in range(5):
        torch.cuda.synchronize()
        f = torch.cuda.Event(enable_timing=True); a0 = torch.cuda.Event(enable_timing=True)
        a1 = torch.cuda.Event(enable_timing=True); b0 = torch.cuda.Event(enable_timing=True)
        b1 = torch.cuda.Event(enable_timing=True)
        f.record(cur)
        sA.wait_stream(cur); sB.wait_stream(cur)
        with torch.cuda.stream(sA):
            a0.record(); gemm(); a1.record()
        with torch.cuda.stream(sB):
            b0.record(); gather(); b1.record()
        cur.wait_stream(sA); cur.wait_stream(sB)
        torch.cuda.synchronize()
        out.append({"gemm_chain_start": f.elapsed_time(a0), "gemm_chain_end": f.elapsed_time(a1),
                    "gather_start": f.elapsed_time(b0), "gather_end": f.elapsed_time(b1)})
    return out[-1]


res = {"detail_ms_since_fork": both_detail(), "gemm_chain_ms (split + gather_rows + GEMM)": timed(gemm), "gather_ms": timed(gather),
       "both_gemm_first_ms": timed(lambda: both(True)), "both_gather_first_ms": timed(lambda: both(False))}
print(json.dumps(res, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "overlap_probe.json").write_text(json.dumps(res))
