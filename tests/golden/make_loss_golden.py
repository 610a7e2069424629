"""Golden vectors for PermutationLoss / matching_recall / matching_precision, produced by the REFERENCE's own modules
(/root/reference/src/loss_func.py, /root/reference/src/evaluation_metric.py) imported in this container.
Run from the repo root:  python tests/golden/make_loss_golden.py   -> tests/golden/loss_metric.pt"""
import importlib.util
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, str(REF))
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(str(REF))
    return mod


loss_mod = load("ref_loss_func", REF / "src" / "loss_func.py")
met_mod = load("ref_evaluation_metric", REF / "src" / "evaluation_metric.py")

g = torch.Generator().manual_seed(5)
B, R, C = 5, 12, 14
n1 = torch.tensor([12, 7, 9, 12, 1]); n2 = torch.tensor([14, 14, 5, 12, 3])
pred = torch.rand(B, R, C, generator=g)
pred[0, 0, 0] = 0.0; pred[0, 1, 1] = 1.0; pred[1, 2, 3] = 1e-30      # exercise the log clamps
gt = torch.zeros(B, R, C)
for b in range(B):
    k = int(min(n1[b], n2[b]))
    perm = torch.randperm(int(n2[b]), generator=g)[:k]
    if b != 3:                                                       # pair 3: imposter, all-zero ground truth
        gt[b, torch.arange(k), perm] = 1
p = pred.clone().requires_grad_(True)
loss = loss_mod.PermutationLoss()(p, gt, n1, n2)
loss.backward()
hard = torch.zeros(B, R, C)
for b in range(B):
    k = int(min(n1[b], n2[b]))
    perm = torch.randperm(int(n2[b]), generator=g)[:k]
    keep = torch.rand(k, generator=g) < 0.7
    hard[b, torch.arange(k)[keep], perm[keep]] = 1
hard[4] = 0                                                          # nothing predicted for the last pair
hard[:, :, :] = torch.where(gt.bool() & (torch.rand(B, R, C, generator=g) < 0.5), torch.ones(()), hard) * \
    (hard.sum(1, keepdim=True) <= 1)
# keep it a partial permutation
for b in range(B):
    seen_r, seen_c = set(), set()
    for i, j in hard[b].nonzero().tolist():
        if i in seen_r or j in seen_c:
            hard[b, i, j] = 0
        else:
            seen_r.add(i); seen_c.add(j)
rec = met_mod.matching_recall(hard, gt, n1)
prec = met_mod.matching_precision(hard, gt, n1)
acc = met_mod.matching_accuracy(hard, gt, [n1, n2], 0)
# second case: pairs WITHOUT any predicted match (imposter pairs with k = 0 predict nothing): 0/0 -> the reference's
# NaN rule (matching_precision sets 1, evaluation_metric.py:123)
hard2 = hard.clone()
hard2[1] = 0; hard2[3] = 0; hard2[4] = 0
rec2 = met_mod.matching_recall(hard2, gt, n1)
prec2 = met_mod.matching_precision(hard2, gt, n1)
torch.save({"pred": pred, "gt": gt, "n1": n1, "n2": n2, "loss": loss.detach(), "grad": p.grad, "hard": hard,
            "recall": rec, "precision": prec, "accuracy": acc, "hard_empty": hard2, "recall_empty": rec2,
            "precision_empty": prec2}, ROOT / "tests" / "golden" / "loss_metric.pt")
print("empty-prediction case: recall", rec2.tolist(), "precision", prec2.tolist())
print("loss", float(loss), "recall", rec.tolist(), "precision", prec.tolist())
