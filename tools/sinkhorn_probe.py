"""Times ops.sinkhorn_log alone at the headline shape (256 pairs x 100 x 100) and the ragged / transposed variants.
Run once per kernel choice: FPMATCH_SINKHORN_REG=0 keeps the shared-memory kernel.  Prints one JSON line per case."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fingerprint-matching-code_b200"))
from fpmatch import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(7)
for tag, B, R, C, it, ragged in [("n100_it20", 256, 100, 100, 20, False), ("n100_it10", 256, 100, 100, 10, False),
                                 ("n100_it20_ragged", 256, 100, 100, 20, True), ("n128_it20", 256, 128, 128, 20, False),
                                 ("n50_it20", 256, 50, 50, 20, False)]:
    s = torch.randn(B, R, C, generator=g).to(dev)
    if ragged:
        n1 = torch.randint(R // 2, R + 1, (B,), generator=g).to(dev)
        n2 = torch.randint(C // 2, C + 1, (B,), generator=g).to(dev)
    else:
        n1 = torch.full((B,), R, dtype=torch.int64, device=dev)
        n2 = torch.full((B,), C, dtype=torch.int64, device=dev)
    for _ in range(5):
        out, out_t = ops.sinkhorn_log(s, n1, n2, it, 0.05, True, want_t=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        out, out_t = ops.sinkhorn_log(s, n1, n2, it, 0.05, True, want_t=True)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"case": tag, "reg": os.environ.get("FPMATCH_SINKHORN_REG", "1"), "us_per_call": 1e3 * e0.elapsed_time(e1) / reps,
                      "checksum": float(out.double().sum()), "rowsum_err": float((out.sum(2) - 1).abs().max())}))
