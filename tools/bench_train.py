#!/usr/bin/env python
"""BASELINE.json config 3: stage-1 training step (100 keypoints, global batch 64 = 8 pairs per GPU on 8 GPUs), data-parallel
with the NCCL gradient all-reduce of DistributedDataParallel (`fpmatch.dist.wrap_ddp`).

    python tools/bench_train.py                               # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_train.py --pairs-per-rank 8               # N GPUs, weak scaling (global batch 8 N)

The step follows /root/reference/src/train/training_loop.py:21-67 for stage 1 (train.py:157-181): zero_grad ->
Net.forward (train mode) -> PermutationLoss -> backward -> clip_grad_norm_(5.0) -> AdamW(lr 1e-4 warm-up, wd 1e-4), from
the backbone's feature maps on (the backbone is out of scope and frozen here).  Device-timed with CUDA events, barrier on
both sides, max over ranks.  Per run it reports
  ms_step                 the full DDP step,
  ms_step_local           the same step under `no_sync()` (no all-reduce at all)  -> exposed all-reduce = difference,
  allreduce_alone_ms      one NCCL all-reduce of a flat buffer of the same number of gradient bytes, nothing else running,
  graph_ms_step           (1 GPU only, --graph) the whole step captured in ONE CUDA graph and replayed: the floor once
                          the ~300 kernel launches of an 8-pair step stop being issued one by one from python.
One JSON line on rank 0, appended to gpurun_out/train_scale.jsonl."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs-per-rank", type=int, default=8)
    ap.add_argument("--keypoints", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from fpmatch import dist as fdist, ops, synth
    from src.loss_func import PermutationLoss
    from src.model.ngm import Net

    torch.manual_seed(0)
    net = Net(regression=False).to(dev).train()
    frozen = ("encoder_k.", "final_row.", "final_col.", "match_cls.", "node_layers.", "edge_layers.")
    for k, p in net.named_parameters():
        if k.startswith(frozen):
            p.requires_grad_(False)                       # train.py:168-181 freezes the k-branch in stage 1
    params = [p for p in net.parameters() if p.requires_grad]
    model = fdist.wrap_ddp(net, dev) if world > 1 else net
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4, capturable=args.graph)
    B = args.pairs_per_rank
    data = synth.make_batch(B, args.keypoints, seed=7 + rank, imposter_every=0, with_kron=False, with_dense_gh=False,
                            fmap_noise=1.0)
    data.pop("label")
    devd = synth.batch_to(data, dev)
    crit = PermutationLoss()
    losses = []

    def step(sync=True, record=False):
        d = dict(devd)
        d["pyg_graphs"] = [g.to(dev) for g in devd["pyg_graphs"]]
        opt.zero_grad(set_to_none=not args.graph)
        ctx = model.no_sync() if (world > 1 and not sync) else torch.autograd.profiler.record_function("step")
        with ctx:
            out = model(d)
            loss = crit(out["ds_mat"], d["gt_perm_mat"], *d["ns"])
            loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
        opt.step()
        if record:
            losses.append(loss.detach())
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(args.warmup):
        step()
    l0 = ops.launch_count()
    ms = timed(lambda: step(record=True), args.steps)
    launches = (ops.launch_count() - l0) / args.steps
    ms_local = timed(lambda: step(sync=False), args.steps) if world > 1 else ms
    t0 = time.perf_counter()
    step(); torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3

    grad_bytes = sum(p.numel() * p.element_size() for p in params if p.grad is not None)
    ar_ms = None
    if world > 1:
        flat = torch.empty(grad_bytes // 4, dtype=torch.float32, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        ar_ms = timed(lambda: dist.all_reduce(flat), 10)

    graph_ms, graph_err = None, None
    if args.graph and world == 1:
        try:
            net.track_lap_status = False
            crit.check_range = False            # the reference's range assertion reads the device: not capturable
            static = dict(devd)
            static["pyg_graphs"] = [g.to(dev) for g in devd["pyg_graphs"]]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=False)
            with torch.cuda.graph(g):
                gl = step()
            torch.cuda.synchronize()
            graph_ms = timed(g.replay, args.steps)
        except Exception as e:                                    # noqa: BLE001  report, do not hide
            graph_err = f"{type(e).__name__}: {str(e)[:300]}"

    if rank == 0:
        lv = [float(x) for x in torch.stack(losses).tolist()] if losses else []
        rec = {"config": f"stage-1 training step, {args.keypoints} keypoints, {B} pairs per GPU, dp{world}",
               "n_gpus": world, "pairs_per_rank": B, "global_batch": B * world, "ms_step": ms,
               "pairs_per_s": B * world / ms * 1e3, "ms_step_local_no_allreduce": ms_local,
               "exposed_allreduce_ms": ms - ms_local if world > 1 else 0.0, "allreduce_alone_ms": ar_ms,
               "allreduce_bytes": grad_bytes if world > 1 else 0, "grad_bytes": grad_bytes,
               "allreduce_busbw_gbs": (2 * (world - 1) / world * grad_bytes / (ar_ms / 1e3) / 1e9) if ar_ms else None,
               "fpmatch_launches_per_step": launches, "host_wall_ms_one_step": host_ms,
               "graph_ms_step": graph_ms, "graph_error": graph_err,
               "loss_first": lv[0] if lv else None, "loss_last": lv[-1] if lv else None,
               "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30, "tag": args.tag}
        print(json.dumps(rec), flush=True)
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        with open(out / "train_scale.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
