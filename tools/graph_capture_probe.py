#!/usr/bin/env python
"""Which part of the stage-1 training step (or of the inference forward) cannot be captured in a CUDA graph?
Captures growing prefixes of the step under torch.cuda.graph in 'global' and 'relaxed' error modes and reports the first
exception of each, plus replay timings of what does capture."""
import json
import sys
import traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import synth
from src.loss_func import PermutationLoss
from src.model.ngm import Net

dev = torch.device("cuda")
torch.manual_seed(0)
net = Net(regression=False).to(dev).train()
net.track_lap_status = False
frozen = ("encoder_k.", "final_row.", "final_col.", "match_cls.", "node_layers.", "edge_layers.")
for k, p in net.named_parameters():
    if k.startswith(frozen):
        p.requires_grad_(False)
params = [p for p in net.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4, capturable=True)
data = synth.make_batch(8, 100, seed=7, imposter_every=0, with_kron=False, with_dense_gh=False, fmap_noise=1.0)
data.pop("label")
devd = synth.batch_to(data, dev)
crit = PermutationLoss()
crit.check_range = False


def fwd():
    d = dict(devd)
    d["pyg_graphs"] = [g.to(dev) for g in devd["pyg_graphs"]]
    out = net(d)
    return out, d


def stage(n):
    opt.zero_grad(set_to_none=False)
    out, d = fwd()
    if n >= 2:
        loss = crit(out["ds_mat"], d["gt_perm_mat"], *d["ns"])
    if n >= 3:
        loss.backward()
    if n >= 4:
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
    if n >= 5:
        opt.step()


def infer():
    net.eval()
    try:
        with torch.no_grad():
            fwd()
    finally:
        net.train()


def timed(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        stage(5)
        infer()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
res = {"eager_step_ms": timed(lambda: stage(5)), "eager_infer_ms": timed(infer)}
names = {0: "inference forward", 1: "train forward", 2: "+ loss", 3: "+ backward", 4: "+ clip", 5: "+ AdamW"}
for mode in ("global", "relaxed"):
    for n in (1, 2, 3, 4, 5, 0):
        key = f"{mode}: {names[n]}"
        try:
            g = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=False)
            with torch.cuda.graph(g, capture_error_mode=mode):
                infer() if n == 0 else stage(n)
            torch.cuda.synchronize()
            res[key] = {"ok": True, "replay_ms": timed(g.replay)}
        except Exception as e:                                      # noqa: BLE001
            tb = traceback.extract_tb(e.__traceback__)
            mine = [f"{Path(f.filename).name}:{f.lineno} {f.name}" for f in tb if "fingerprint-matching" in f.filename or "tools" in f.filename]
            res[key] = {"ok": False, "error": f"{type(e).__name__}: {str(e)[:160]}", "where": mine[-4:]}
            try:
                torch.cuda.synchronize()
            except Exception:                                       # noqa: BLE001
                pass
        print(key, json.dumps(res[key]), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "graph_capture_probe.json").write_text(json.dumps(res, indent=1))
