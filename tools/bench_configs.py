"""Device-timed numbers for the BASELINE.json configurations that are not the bench.py headline:
  config 3  stage-1 training step (forward + backward + AdamW), 100 keypoints: 64 pairs on one GPU and the 8-pair
            per-GPU share of "batch 64 over 8 GPUs";
  config 4  pore-level inference, 400 keypoints/image, batch 32.
CUDA events, 3 warm-up + 5 timed steps, inputs resident.  Usage: python tools/bench_configs.py"""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import ops, synth
from src.loss_func import PermutationLoss
from src.model.ngm import Net

DEV = "cuda"
res = []


def timed(fn, warm=3, steps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_cfg(B, n=100):
    torch.manual_seed(0)
    net = Net(regression=False).to(DEV).train()
    frozen = ("encoder_k.", "final_row.", "final_col.", "match_cls.", "node_layers.", "edge_layers.")
    params = [p for k, p in net.named_parameters() if not k.startswith(frozen)]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)
    data = synth.make_batch(B, n, seed=7, imposter_every=0, with_kron=False, with_dense_gh=False)
    data.pop("label")
    dev = synth.batch_to(data, DEV)
    crit = PermutationLoss()

    def step():
        d = dict(dev)
        d["pyg_graphs"] = [g.to(DEV) for g in dev["pyg_graphs"]]
        opt.zero_grad(set_to_none=True)
        out = net(d)
        loss = crit(out["ds_mat"], d["gt_perm_mat"], *d["ns"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
        opt.step()

    l0 = ops.launch_count()
    ms = timed(step)
    launches = (ops.launch_count() - l0) // 8
    torch.cuda.synchronize()
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    rec = {"config": f"stage-1 training step, {n} keypoints, {B} pairs on 1 GPU", "ms_per_step": ms,
           "pairs_per_s": B / ms * 1e3, "fpmatch_launches_per_step": launches, "peak_mem_gib": mem}
    res.append(rec); print(json.dumps(rec), flush=True)
    del net, opt
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()


def infer_cfg(B, n):
    torch.manual_seed(0)
    net = Net(regression=True).to(DEV).eval()
    data = synth.make_batch(B, n, seed=9, with_kron=False, with_dense_gh=False)
    dev = synth.batch_to(data, DEV)

    def step():
        with torch.no_grad():
            net(dict(dev))

    ms = timed(step)
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    rec = {"config": f"matching-head inference, {n} keypoints, {B} pairs on 1 GPU", "ms_per_step": ms,
           "pairs_per_s": B / ms * 1e3, "peak_mem_gib": mem}
    res.append(rec); print(json.dumps(rec), flush=True)
    del net
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()


infer_cfg(32, 400)
train_cfg(8)
train_cfg(64)
infer_cfg(256, 100)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_configs.json").write_text(json.dumps(res))
