#!/usr/bin/env python
"""Benchmark of the matching head (BASELINE.json metric: matched pairs/sec at 100 keypoints).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the matching head (normalise + feature_align -> 2x SplineConv -> affinities ->
3 NGM layers with Sinkhorn -> Sinkhorn -> AFA-U k -> soft-top-k -> exact LAP -> greedy top-k -> match
classifier) over one batch of 256 synthetic fingerprint pairs with 100 keypoints per image
(BASELINE.json configs[1]); the ResNet backbone is out of scope, so the step starts from its feature maps.

`value`   device-timed pairs/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`     the same metric through the public API with pinned HOST inputs: fpmatch.prefetch.CudaPrefetcher (the
          drop-in for the reference's data_to_cuda call, copying step i+1 on a side stream while step i runs) ->
          Net.forward(data_dict) -> fpmatch.prefetch.HostResultRing (outputs copied to pinned host buffers on a side
          stream, handed to the caller one step later); every step's copies are inside the timed region and the
          region ends only when the last step's results are on the host.
`roofline` for the dominant kernel (the SplineConv slab GEMM): algorithmic FLOPs / live CUDA-event time.
`cpu_baseline` the oracle port of the reference's PyTorch+scipy CPU path on a bounded sample.
--impl reference times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "fingerprint-matching-code_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

N_KPTS = 100
BATCH = 256
WORKLOAD = f"matching-head inference, {N_KPTS} keypoints/image, batch {BATCH} pairs, fp32"
METRIC = "matched pairs/sec"
# why the CPU arm is a port and what that means for the figure (BASELINE.md section 3 planned to import the reference)
PORT_NOTE = ("the reference itself cannot be imported here (torch_geometric, torch_sparse, torch_spline_conv, pygmtools "
             "absent; its CUDA extension does not compile against torch 2.11), so this times oracle/: the reference's own "
             "modules' arithmetic where they import (feature_align, soft_topk, AFA-U, affinity, scipy LAP) and "
             "restatements elsewhere (SplineConv as per-edge index_select + bmm, SAGE mean as argsort + index_add_ instead "
             "of torch_sparse's CSR spmm, pygmtools' per-pair Sinkhorn loop) - an approximation of the reference's CPU "
             "cost, not a measurement of its code")


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # first use of a query in a process (and of NVML on a fresh box) can take tens of milliseconds inside the
            # driver: pay that here, not in the sampling thread during the timed region (seen once as a 25 ms gap in
            # kernel submission: 8.8 instead of 6.5 ms per step)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown",
                 "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown",
                 "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, attr in names.items():
                    if mask & getattr(nv, attr, 0):
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        if self.nv is not None:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def build_model(regression=True):
    from src.model.ngm import Net
    torch.manual_seed(0)
    return Net(regression=regression).eval()


def cpu_reference_run(steps, warmup, sample_pairs):
    """The reference's CPU path (oracle port, reference loop structure) on `sample_pairs` pairs per step."""
    from fpmatch import synth
    from oracle import head
    net = build_model()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    data = synth.make_batch(sample_pairs, N_KPTS, seed=1234, with_kron=True, with_dense_gh=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        head.forward_head(sd, synth.clone_batch(data), data["fmaps"], regression=True, feature_align_loops=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return sample_pairs / sec, sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--gemm", default=os.environ.get("FPMATCH_GEMM", None))
    ap.add_argument("--cpu-sample", type=int, default=32, help="pairs per step of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace-ops", action="store_true",
                    help="diagnostic: report the largest host-side gaps between op launches of the timed loop")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostic: no NVML sampling thread")
    ap.add_argument("--e2e-inflight", type=int, default=int(os.environ.get("FPMATCH_BENCH_E2E_INFLIGHT", "2")),
                    help="batches in flight in the end-to-end loop whose backbone maps cross PCIe every step (r2: 31.3 k "
                         "pairs/s with 2, 29.3 k with 1); the resident-maps variant uses --inflight")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("FPMATCH_BENCH_INFLIGHT", "2")),
                    help="batches in flight in the device-timed loop: step i is issued on stream i %% inflight, so the "
                         "latency-bound tail of one batch (Sinkhorn, LAP, AFA-U) overlaps the tensor-bound front of the next")
    args = ap.parse_args()
    rank, world, local = dist_env()
    cores = os.cpu_count()
    torch.set_num_threads(max(1, cores or 1))

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        warm = min(args.warmup, 1)
        pps, sec = cpu_reference_run(steps, warm, args.cpu_sample)
        line = {
            "impl": "reference", "metric": METRIC, "value": pps, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "keypoints": N_KPTS, "note": "CPU path, bounded sample"},
            "cpu_baseline": {"value": pps, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.cpu_sample} pairs x {N_KPTS} keypoints per step, oracle port of the "
                                       "reference's PyTorch+scipy path with its python loop structure",
                             "note": PORT_NOTE},
            "e2e": {"value": pps, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback exists)"
    if world > 1 and os.environ.get("FPMATCH_BENCH_AFFINITY", "1") != "0" and hasattr(os, "sched_setaffinity"):
        # one contiguous slice of the host cores per rank: the rank's pinned staging buffers are first touched (and its
        # copies driven) from cores of one NUMA node instead of wherever the scheduler puts the process
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]) or set(cpus))
        except OSError:
            pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from fpmatch import ops, synth
    if args.gemm:
        ops.set_gemm_mode(args.gemm)

    B = args.batch
    net = build_model().to(dev)
    host = synth.make_batch(B, N_KPTS, seed=1234 + rank, with_kron=False, with_dense_gh=False)

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
    resident = synth.batch_to(synth.clone_batch(host), dev)

    def tensor_bytes(v):
        if isinstance(v, torch.Tensor):
            return v.numel() * v.element_size()
        if isinstance(v, (list, tuple)):
            return sum(tensor_bytes(x) for x in v)
        if hasattr(v, "edge_index"):
            return sum(tensor_bytes(getattr(v, k)) for k in ("x", "edge_index", "edge_attr", "ptr", "eptr"))
        return 0
    h2d = sum(tensor_bytes(v) for v in host.values())

    def step_resident():
        with torch.no_grad():
            return net(dict(resident))    # the forward overwrites graph.x in place, as the reference does

    from fpmatch.prefetch import MatchingPipeline, lanes as _lanes
    lanes = _lanes(dev, args.inflight) if args.inflight > 1 else []

    stagger = os.environ.get("FPMATCH_STAGGER", "1") != "0"
    enqueue_marks = []
    step_marks = None       # per-step completion events of the timed loop (diagnostic: `step_end_ms` in the JSON line)

    def run_steps(n):
        """n steps; with --inflight k > 1 step i runs on stream i % k (each batch's kernels stay ordered on their own
        stream; consecutive batches are independent, as in serving).  All lanes join the current stream at the end."""
        out = None
        # The host enqueues a step in ~3 ms, the device needs ~6.5: left alone the host runs the whole loop ahead and the
        # caching allocator has to find fresh 2 GB blocks for every step still in flight (cudaMalloc inside the timed
        # region, and - once, on a fresh box - an allocator retry that drained the device: 20-80 ms gaps in
        # submission, 8.5 / 19 ms per step instead of 6.4).  A serving loop reads its results, which bounds the
        # run-ahead (MatchingPipeline: the result ring); the device-timed loop bounds it the same way: at most
        # `depth` steps are in flight, the host waits for the oldest one's completion event before it issues the next.
        depth = max(1, len(lanes)) + 1
        pending = []

        def throttle():
            if len(pending) >= depth:
                pending.pop(0).synchronize()

        def issued():
            ev = torch.cuda.Event()
            ev.record()
            pending.append(ev)

        if not lanes:
            for _ in range(n):
                throttle()
                out = step_resident()
                issued()
            return out
        cur = torch.cuda.current_stream(dev)
        for s_ in lanes:
            s_.wait_stream(cur)
        ops.set_gemm_max_clusters(MatchingPipeline.GEMM_CLUSTERS_IN_FLIGHT)      # as the streaming API does
        for i in range(n):
            throttle()
            with torch.cuda.stream(lanes[i % len(lanes)]):
                if stagger:                                  # as MatchingPipeline does: fronts of consecutive batches in turn
                    net.front_gate = getattr(net, "front_done", None) if i else None
                out = step_resident()
                issued()
                if step_marks is not None:
                    ev = torch.cuda.Event(enable_timing=True); ev.record(); step_marks.append(ev)
                    enqueue_marks.append(time.perf_counter())
        ops.set_gemm_max_clusters(0)
        for s_ in lanes:
            cur.wait_stream(s_)
        return out

    out_keys = ("ds_mat", "perm_mat", "k_prob", "cls_prob")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)                # NVML is initialised and queried once before any timing
    if args.no_clock_sampler:
        sampler.nv = None
    for _ in range(max(args.warmup, 3)):
        step_resident()
    run_steps(4 * max(1, args.inflight) + 2)     # the throttled steady state of run_steps, long enough for the allocator
    barrier()
    # the collector is parked for the timed regions (as timeit does): a generation-2 pass over the process's objects
    # takes several milliseconds of host time in the middle of a 65 ms measurement
    import gc
    gc.collect()
    gc.disable()

    # ---- device-resident throughput + live GEMM timing for the roofline ----
    with sampler as clk:
        l0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        step_marks = []
        t_host0 = time.perf_counter()
        if args.trace_ops:
            ops.op_trace_start()
        out = run_steps(args.steps)
        if args.trace_ops:
            tr = ops.op_trace_stop()
            gaps = sorted(((tr[i + 1][0] - tr[i][0]) * 1e3, tr[i][1], tr[i + 1][1], (tr[i][0] - t_host0) * 1e3)
                          for i in range(len(tr) - 1))[-6:]
            print("host gaps (ms, after op, before op, at ms):", [(round(g, 2), a, b_, round(t_, 1)) for g, a, b_, t_ in gaps],
                  file=sys.stderr)
        marks, step_marks = step_marks, None
        e1.record()
        barrier()
        launches = ops.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    step_end_ms = [round(e0.elapsed_time(m), 3) for m in marks]
    step_enqueued_ms = [round((t - t_host0) * 1e3, 3) for t in enqueue_marks[-len(marks):]] if marks else []
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = B * world / (ms_step / 1e3)

    # ---- live timing of the dominant kernel for the roofline: CUDA events around every GEMM launch ON ITS OWN STREAM.
    # The product path runs the two images' chains on two streams; an event pair on one of them would also measure the
    # time the launch spends queued behind the other stream's GEMM, so this pass keeps everything on one stream.
    fork_was = net.graph_fork
    net.graph_fork = False
    rf_steps = max(2, min(args.steps, 5))
    step_resident()
    barrier()
    ops.gemm_profile_start()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(rf_steps):
        step_resident()
    r1.record()
    barrier()
    gemm_events = ops.gemm_profile_stop()
    ms_step_single_stream = r0.elapsed_time(r1) / rf_steps
    net.graph_fork = fork_was

    # dominant kernel = the SplineConv slab GEMM (largest FLOP count in the step).  With the slab planner the launch
    # computes only the tiles listed in its device-side table: FLOPs = executed tiles x 256 x 128 x K x 2.
    def flops_of(e):
        if e[0] == "3xf16-slabs":
            return 2.0 * e[1].tiles_used() * 256 * 128 * e[3]
        return 2.0 * e[1] * e[2] * e[3]
    big = max(gemm_events, key=flops_of)
    if big[0] == "3xf16-slabs":
        same = [e for e in gemm_events if e[0] == big[0] and e[1].T == big[1].T and e[2] == big[2]]
        flops = sum(flops_of(e) for e in same) / len(same)
        kernel_label = (f"spline slab GEMM [3xf16, tile table]: {big[1].tiles_used()} tiles of 256x128x{big[3]} "
                        f"(the dense product would be M={big[1].T} N={big[2]} K={big[3]} = "
                        f"{(big[1].T_pad // 256) * (big[2] // 128)} tiles)")
        big = ("3xf16", big[1].T, big[2], big[3], big[4], big[5])
    else:
        same = [e for e in gemm_events if (e[0], e[1], e[2], e[3]) == big[:4]]
        flops = 2.0 * big[1] * big[2] * big[3]
        kernel_label = None
    gemm_ms = sum(e[4].elapsed_time(e[5]) for e in same) / len(same)
    peaks, peak_kind = measured_peaks()
    mode = big[0]
    if mode == "fp32":
        # CUDA-core path: bounded by the fp32 FMA pipe, 148 SMs x 128 lanes x 2 flop x max clock
        peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
        peak_note = "fp32 CUDA-core FMA peak at max SM clock (nominal; no measured figure)"
    else:
        # the kernel is timed inside a long step -> the SUSTAINED measured dense bf16 figure of MEASURED_PEAKS.json
        peak = peaks["bf16_tflops_sustained"]
        peak_note = (f"{peak_kind} MEASURED_PEAKS.json bf16_tflops_sustained = {peaks['bf16_tflops_sustained']} "
                     f"(burst {peaks['bf16_tflops']}); fp16 and bf16 MMAs run at the same rate")
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this shape: read from the summary
    # of the last `ncu --set full` capture of THIS workload (tools/summarize_profiles.py writes profiles/LATEST_NCU.json
    # next to profiles/<tag>_ncu_full.md); null when no capture of this kernel at this grid is on file
    traffic, traffic_src = None, None
    latest = ROOT / "profiles" / "LATEST_NCU.json"
    if latest.exists() and kernel_label is not None and B == BATCH:
        rec = json.loads(latest.read_text())
        g = rec.get("slab_gemm")
        if g:
            traffic, traffic_src = g["dram_bytes"], f"{rec['file']} ({g['kernel']}, grid {g['grid']}, {g['time_us']:.0f} us under ncu)"
    achieved = flops / (gemm_ms / 1e3) / 1e12
    gemm_share = gemm_ms * len(same) / rf_steps / ms_step_single_stream

    # ---- end to end through the public API with pinned host inputs: CudaPrefetcher (the host->device copy of step
    # i+1 runs on a side stream while step i is matched) -> Net.forward -> outputs copied to the host.  Every step's
    # H2D and D2H happen inside the timed region; wall clock, max over ranks.
    from fpmatch.prefetch import CudaPrefetcher, HostResultRing, MatchingPipeline
    e2e_steps = max(3, min(args.steps, 10))

    ring = HostResultRing(device=dev)                        # pinned result buffers and device staging buffers
    feeder = CudaPrefetcher([], device=dev)                  # are allocated once and reused across steps

    def e2e_run(nsteps, hb=None, extra=None, fd=None, inflight=1):
        """fpmatch.prefetch.MatchingPipeline = the public streaming API: pinned host batches in, pinned host results
        out, `inflight` batches on the device at once."""
        pipe = MatchingPipeline(net, [hb if hb is not None else host] * nsteps, keys=out_keys, device=dev,
                                inflight=max(1, inflight), feeder=fd or feeder, ring=ring, extra=extra)
        res, got = None, 0
        for res in pipe:
            got += 1
        assert res is not None and got == nsteps
        return res

    e2e_run(6, inflight=args.e2e_inflight)       # more steps than the prefetcher has staging slots: all buffers exist
    barrier()
    t0 = time.perf_counter()
    res = e2e_run(e2e_steps, inflight=args.e2e_inflight)
    barrier()
    wall = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([wall], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world / t.item()
    d2h = sum(r.numel() * r.element_size() for r in res)

    # Second end-to-end variant: the backbone maps are already on the device - the situation of the reference's own
    # scripts, where the ResNet runs in the same process right before the head (ngm.py:228-238) - and only the
    # keypoints, graphs, ground truth and labels (what collate_fn produces on the host) cross PCIe every step.
    host_nomaps = {k: v for k, v in host.items() if k != "fmaps"}
    feeder2 = CudaPrefetcher([], device=dev)
    maps_dev = {"fmaps": resident["fmaps"]}
    h2d_nomaps = sum(tensor_bytes(v) for v in host_nomaps.values())
    e2e_run(6, host_nomaps, maps_dev, feeder2, inflight=args.inflight)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps, host_nomaps, maps_dev, feeder2, inflight=args.inflight)
    barrier()
    wall2 = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([wall2], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_resident_value = B * world / t.item()
    # the e2e step moves 262 MB host->device: report what this box's link gives for a plain pinned copy of that size,
    # so a link-bound e2e figure can be told from a compute-bound one
    probe = host["fmaps"][0][0]
    dst = torch.empty_like(probe, device=dev)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(4):
        dst.copy_(probe, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * probe.numel() * probe.element_size() / (c0.elapsed_time(c1) / 1e3) / 1e9

    gc.enable()
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pps, sec = cpu_reference_run(4, 1, args.cpu_sample)           # ~10 s of CPU work on the box's host cores
        cpu_base = {"value": pps, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port", "note": PORT_NOTE,
                    "sample": f"4 timed passes (+1 warm-up) over {args.cpu_sample} pairs x {N_KPTS} keypoints ({sec:.1f} s each), oracle port of the reference's "
                              "PyTorch+scipy CPU path with its per-point / per-pair python loops"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "keypoints": N_KPTS, "pairs_per_gpu": B, "gemm_mode": ops.gemm_mode(),
                       "l2": "inputs + intermediates per step (>1 GB) exceed the 126 MB L2; no explicit flush",
                       "dead_ke_computed": True,
                       "batches_in_flight": max(1, args.inflight),
                       "pipeline": (f"step i issued on stream i % {args.inflight}: consecutive batches are independent, so "
                                    "the latency-bound tail of one overlaps the tensor-bound front of the next; every "
                                    "step is a complete forward of one 256-pair batch") if args.inflight > 1 else "one stream",
                       "parallelism": f"dp{world} (independent pairs, no data-path collective)"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kernel_label or f"gemm_nt[{mode}] M={big[1]} N={big[2]} K={big[3]}",
                         "launch_ms": gemm_ms, "share_of_step": gemm_share,
                         "ms_per_step_single_stream": ms_step_single_stream, "peak_source": peak_note,
                         "flops_per_launch": flops,
                         "note": ("achieved = algorithmic 2*M*N*K of the executed tiles / live CUDA-event time of the launch. "
                                  "fp32-faithful results need 3 fp16 MMAs per product (hi*hi, hi*lo, lo*hi), which are NOT "
                                  "counted as work: the scheme's own ceiling is frac = 1/3; issued-MMA rate = 3 x achieved = "
                                  f"{3 * achieved:.0f} TFLOP/s = {3 * achieved / peak:.2f} of the peak") if mode == "3xf16" else None},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "h2d_link_gbs_measured": h2d_gbs, "h2d_ms_per_step_at_that_rate": h2d / h2d_gbs / 1e6,
                    "batches_in_flight": max(1, args.e2e_inflight),
                    "maps_resident": {"value": e2e_resident_value, "batches_in_flight": max(1, args.inflight), "unit": "pairs/s", "h2d_bytes_per_step": h2d_nomaps,
                                      "d2h_bytes_per_step": d2h,
                                      "note": "same public-API path, backbone maps produced on the device (in-process "
                                              "backbone, as in the reference's scripts); keypoints, graphs, ground "
                                              "truth and labels still cross PCIe every step"}},
            "gpu_launches": launches,
            "step_end_ms": step_end_ms,
            "step_enqueued_ms": step_enqueued_ms,
            "clocks": clk.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
