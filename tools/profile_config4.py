"""One inference forward of config 4 (400 keypoints, 32 pairs) for `ncu --metrics gpu__time_duration.sum`."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import synth
from src.model.ngm import Net
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(0)
net = Net(regression=True).to("cuda").eval()
dev = synth.batch_to(synth.make_batch(B, n, seed=9, with_kron=False, with_dense_gh=False), "cuda")
for _ in range(3):
    with torch.no_grad():
        net(dict(dev))
torch.cuda.synchronize()
print("done")
