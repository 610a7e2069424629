#!/usr/bin/env python
"""Device-timed throughput of the matching head at any BASELINE.json inference configuration on N GPUs of one node
(weak scaling: every rank matches its own shard; no collective on the data path).

    python tools/bench_scale.py --keypoints 400 --pairs-per-rank 32                 # config 4 on 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_scale.py --keypoints 400 --pairs-per-rank 32                    # config 4 on N GPUs

CUDA events around K steps after W warm-ups, barrier + synchronize on both sides, max over ranks (as bench.py).
One JSON line on rank 0, appended to gpurun_out/scale_configs.jsonl."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--keypoints", type=int, default=400)
    ap.add_argument("--pairs-per-rank", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ragged", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from fpmatch import ops, synth
    from src.model.ngm import Net
    torch.manual_seed(0)
    net = Net(regression=True).to(dev).eval()
    B = args.pairs_per_rank
    data = synth.batch_to(synth.make_batch(B, args.keypoints, seed=9 + rank, ragged=args.ragged,
                                           n_min=(3 * args.keypoints) // 4, with_kron=False, with_dense_gh=False), dev)

    def step():
        with torch.no_grad():
            net(dict(data))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    if rank == 0:
        rec = {"config": f"matching-head inference, {args.keypoints} keypoints{' (ragged)' if args.ragged else ''}, "
                         f"{B} pairs per GPU, dp{world}", "n_gpus": world, "pairs_per_rank": B, "ms_per_step": ms,
               "pairs_per_s": B * world / ms * 1e3, "fpmatch_launches_per_step": (ops.launch_count() - l0) / args.steps,
               "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30, "scaling": "weak"}
        print(json.dumps(rec), flush=True)
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        with open(out / "scale_configs.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
