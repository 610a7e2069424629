"""Host-side cost of one stage-1 training step at the 8-pair per-GPU share of config 3: cProfile over 20 steps
(no synchronisation inside the loop) + wall time per step with and without a trailing synchronise."""
import cProfile
import pstats
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import synth
from src.loss_func import PermutationLoss
from src.model.ngm import Net
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
net = Net(regression=False).to("cuda").train()
frozen = ("encoder_k.", "final_row.", "final_col.", "match_cls.", "node_layers.", "edge_layers.")
params = [p for k, p in net.named_parameters() if not k.startswith(frozen)]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)
data = synth.make_batch(B, 100, seed=7, imposter_every=0, with_kron=False, with_dense_gh=False)
data.pop("label")
dev = synth.batch_to(data, "cuda")
crit = PermutationLoss()


def step():
    d = dict(dev); d["pyg_graphs"] = [g.to("cuda") for g in dev["pyg_graphs"]]
    opt.zero_grad(set_to_none=True)
    out = net(d)
    loss = crit(out["ds_mat"], d["gt_perm_mat"], *d["ns"])
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t_enqueue = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
t_total = (time.perf_counter() - t0) / 20 * 1e3
print(f"B={B}: host enqueue {t_enqueue:.2f} ms/step, with final sync {t_total:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
