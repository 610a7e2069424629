"""How sensitive is the stage-1 loss trajectory to rounding?  Runs the 100-step golden set-up on the GPU with
different dense-contraction engines (same mathematics, different summation order) and twice with the same engine
(the GNN weight-gradient atomics are the only run-to-run difference), and prints pairwise relative differences
next to the distance to the CPU oracle's golden trajectory."""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import ops, synth
from oracle import train as otrain
from src.model.ngm import Net

gold = json.loads((ROOT / "tests" / "golden" / "train_trajectory.json").read_text())
DEV = "cuda"


def loss_fn(ds, gt, n1, n2):
    B, R, C = ds.shape
    mask = (torch.arange(R, device=ds.device)[None, :, None] < n1.view(B, 1, 1)) & \
           (torch.arange(C, device=ds.device)[None, None, :] < n2.view(B, 1, 1))
    return (torch.nn.functional.binary_cross_entropy(ds, gt, reduction="none") * mask).sum() / n1.sum().float()


def run(mode, steps):
    ops.set_gemm_mode(mode)
    torch.manual_seed(0)
    net = Net(regression=False)
    sd = net.state_dict()
    names = set(otrain.trainable_names(sd))
    net = net.to(DEV).train()
    params = [q for k, q in net.named_parameters() if k in names]
    lr_at = lambda t: otrain.warmup_lr(t, gold["lr"], gold.get("warmup_epochs", 1), gold.get("steps_per_epoch", 10 ** 9)) \
        if "warmup_epochs" in gold else gold["lr"]
    opt = torch.optim.AdamW(params, lr=lr_at(0), weight_decay=gold["weight_decay"])
    out_l = []
    for t in range(steps):
        for gp in opt.param_groups:
            gp["lr"] = lr_at(t)
        d = synth.make_batch(gold["B"], gold["n"], seed=gold["seed_base"] + t, imposter_every=0, with_kron=False,
                             fmap_noise=gold["fmap_noise"])
        d.pop("label")
        d = synth.batch_to(d, DEV)
        opt.zero_grad()
        out = net(d)
        loss = loss_fn(out["ds_mat"], d["gt_perm_mat"], d["ns"][0], d["ns"][1])
        loss.backward()
        torch.nn.utils.clip_grad_norm_([q for q in params if q.grad is not None], max_norm=gold["clip"])
        opt.step()
        out_l.append(loss.item())
    return out_l


steps = gold["steps"]
runs = {"oracle_fp32": gold["loss_fp32"]}
if "loss_fp64" in gold:
    runs["oracle_fp64"] = gold["loss_fp64"]
for name, mode in (("gpu_3xf16_a", "3xf16"), ("gpu_3xf16_b", "3xf16"), ("gpu_fp32", "fp32"), ("gpu_3xtf32", "3xtf32")):
    runs[name] = run(mode, steps)
ops.set_gemm_mode("3xf16")
at = [0, 1, 2, 4, 9, 19, 29, 49, 69, 99]
rel = lambda a, b: [abs(a[i] - b[i]) / max(abs(b[i]), 1e-30) for i in at]
table = {}
keys = list(runs)
for i, a in enumerate(keys):
    for b in keys[i + 1:]:
        table[f"{a} vs {b}"] = rel(runs[a], runs[b])
print("steps", at)
for k, v in table.items():
    print(f"{k:32s}", " ".join(f"{x:9.2e}" for x in v))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "train_chaos_probe.json").write_text(json.dumps({"steps_shown": at, "runs": runs, "rel": table}))
