"""Debug aid: train the CPU oracle k steps, load its parameters into the GPU model and compare the gradients of
step k parameter by parameter (finds regimes where a backward kernel departs from autograd)."""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import synth
from oracle import head, train as otrain
from src.model.ngm import Net

gold = json.loads((ROOT / "tests" / "golden" / "train_trajectory.json").read_text())
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
DEV = "cuda"


def batch(t):
    d = synth.make_batch(gold["B"], gold["n"], seed=gold["seed_base"] + t, imposter_every=0, with_kron=True)
    d.pop("label")
    return d


torch.manual_seed(0)
net = Net(regression=False)
p = {k: v.detach().clone() for k, v in net.state_dict().items()}
names = otrain.trainable_names(p)
params = [p[k].requires_grad_(True) for k in names]
opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-4)
for t in range(K + 1):
    d = batch(t)
    opt.zero_grad()
    out = head.forward_head(p, d, d["fmaps"], regression=False, training=True, keep_graph=True)
    loss = otrain.permutation_loss(out["ds_mat"], d["gt_perm_mat"], d["ns"][0], d["ns"][1])
    loss.backward()
    if t == K:
        break
    torch.nn.utils.clip_grad_norm_([q for q in params if q.grad is not None], 5.0)
    opt.step()
print("oracle loss at step", K, float(loss))
sd = {k: v.detach().clone() for k, v in p.items()}
net.load_state_dict(sd)
net = net.to(DEV).train()
dg = synth.batch_to(batch(K), DEV)
outg = net(dg)
ds, gt, n1, n2 = outg["ds_mat"], dg["gt_perm_mat"], dg["ns"][0], dg["ns"][1]
B, R, C = ds.shape
mask = (torch.arange(R, device=DEV)[None, :, None] < n1.view(B, 1, 1)) & (torch.arange(C, device=DEV)[None, None, :] < n2.view(B, 1, 1))
lg = (torch.nn.functional.binary_cross_entropy(ds, gt, reduction="none") * mask).sum() / n1.sum().float()
lg.backward()
print("gpu loss", lg.item())
inter = outg["_fpm_inter"]
print("fwd: s", (inter["s"].detach().cpu() - out["inter"]["s"].detach()).abs().max().item(),
      "ss", (inter["ss"].detach().cpu() - out["inter"]["ss"].detach()).abs().max().item(),
      "ds", (ds.detach().cpu() - out["ds_mat"].detach()).abs().max().item(),
      "Kp", (inter["Kp"].detach().cpu() - out["inter"]["Kp"].detach()).abs().max().item())
named = dict(net.named_parameters())
rows = []
for k in names:
    go = p[k].grad
    gg = named[k].grad
    if go is None and gg is None:
        continue
    if go is None or gg is None:
        rows.append((k, "MISSING", go is None, gg is None)); continue
    e = (gg.cpu() - go).abs().max().item(); m = go.abs().max().item()
    rows.append((k, e / max(m, 1e-30), e, m))
rows.sort(key=lambda r: -(r[1] if isinstance(r[1], float) else 1e9))
for r in rows[:25]:
    print(r)

# ---- part 2: compare PARAMETERS after K+1 updates (GPU trained from scratch vs the oracle above)
torch.nn.utils.clip_grad_norm_([q for q in params if q.grad is not None], 5.0)
opt.step()
torch.manual_seed(0)
net2 = Net(regression=False).to(DEV).train()
nm2 = dict(net2.named_parameters())
params2 = [nm2[k] for k in names]
opt2 = torch.optim.AdamW(params2, lr=1e-3, weight_decay=1e-4)
for t in range(K + 1):
    dg = synth.batch_to(batch(t), DEV)
    opt2.zero_grad()
    o2 = net2(dg)
    ds, gt, n1, n2 = o2["ds_mat"], dg["gt_perm_mat"], dg["ns"][0], dg["ns"][1]
    l2 = (torch.nn.functional.binary_cross_entropy(ds, gt, reduction="none") * mask).sum() / n1.sum().float()
    l2.backward()
    gn = torch.nn.utils.clip_grad_norm_([q for q in params2 if q.grad is not None], 5.0)
    opt2.step()
print("gpu grad norm at last step", float(gn))
rows = []
for k in names:
    a = nm2[k].detach().cpu(); b = p[k].detach()
    rows.append((k, (a - b).abs().max().item(), b.abs().max().item(), int(((a - b).abs() > 1e-4).sum())))
rows.sort(key=lambda r: -r[1])
print("parameter differences after", K + 1, "updates (name, max abs diff, max |param|, #entries off by > 1e-4)")
for r in rows[:20]:
    print(r)
