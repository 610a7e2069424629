"""Three stage-1 training steps (forward + backward + AdamW) at 100 keypoints for `ncu --metrics gpu__time_duration.sum`."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import synth
from src.loss_func import PermutationLoss
from src.model.ngm import Net
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
net = Net(regression=False).to("cuda").train()
frozen = ("encoder_k.", "final_row.", "final_col.", "match_cls.", "node_layers.", "edge_layers.")
params = [p for k, p in net.named_parameters() if not k.startswith(frozen)]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)
data = synth.make_batch(B, 100, seed=7, imposter_every=0, with_kron=False, with_dense_gh=False)
data.pop("label")
dev = synth.batch_to(data, "cuda")
crit = PermutationLoss()
for _ in range(3):
    d = dict(dev); d["pyg_graphs"] = [g.to("cuda") for g in dev["pyg_graphs"]]
    opt.zero_grad(set_to_none=True)
    out = net(d)
    loss = crit(out["ds_mat"], d["gt_perm_mat"], *d["ns"])
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
