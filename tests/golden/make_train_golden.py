"""Generates tests/golden/train_trajectory.json: the CPU oracle's stage-1 loss trajectory (BASELINE.json config 3 at
a size the CPU finishes in minutes).  Run from the repo root:  python tests/golden/make_train_golden.py [--fp64]

Set-up (mirrored by tests/test_gpu_train.py::test_stage1_loss_trajectory_100_steps):
  Net(regression=False), torch.manual_seed(0) default initialisation, train mode;
  step t uses a FRESH batch  synth.make_batch(B, n, seed=1000+t, imposter_every=0)  without labels (genuine pairs
  from get_pair(): no cls_loss) whose two images carry INDEPENDENT random feature maps.  On any learnable synthetic
  task (a repeated batch, or image-2 maps = image-1 maps + noise) AdamW at the reference's lr = 1e-3 drives the
  loss to < 1e-25 within 15-35 steps (measured), after which a relative comparison is meaningless; on a stream of
  unrelated pairs the loss stays O(1) for all 100 steps while every update still moves every parameter, so the
  loss at step t is a well-conditioned function of the whole update history;
  AdamW(weight_decay 1e-4) on the stage-1 parameter group, clip_grad_norm_ 5.0 (train.py:157-181,
  training_loop.py:59-61), with the reference's learning-rate warm-up: LR = 1e-3 (stage1.yml) is scaled by
  (epoch + 1) / 10 during the first 10 epochs (train.py:246-252,297; utils/scheduler.py:11-13) and an epoch is
  3 x num_iterations = 75 steps (training_loop.py:21-22, stage1.yml), so steps 0-74 run at 1e-4 and 75-99 at 2e-4.
  (A first version of this file used a constant 1e-3: the loss then collapses chaotically between steps 60 and 95
  and even the fp32 oracle is 46 % away from its own fp64 evaluation at step 69.)
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch  # noqa: E402

from fpmatch import synth  # noqa: E402
from oracle import head, train as otrain  # noqa: E402
from src.model.ngm import Net  # noqa: E402

import argparse

ap = argparse.ArgumentParser()
ap.add_argument("--fp64", action="store_true")
ap.add_argument("--config", default="n14", choices=["n14", "n100"],
                help="n14: 3 pairs x 14 keypoints, fresh unrelated pairs (flat loss, long history); "
                     "n100: BASELINE.json config 3's per-GPU share, 8 genuine pairs x 100 keypoints, image-2 maps = "
                     "image-1 maps + noise on a fixed cycle of 4 batches (the loss falls)")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--noise", type=float, default=None)
ap.add_argument("--cycle", type=int, default=None)
ap.add_argument("--out", default=None)
ARGS = ap.parse_args()
if ARGS.config == "n14":
    B, N_KPTS, FMAP_NOISE, CYCLE = 3, 14, None, 0
else:
    B, N_KPTS, FMAP_NOISE, CYCLE = 8, 100, 1.0, 4
if ARGS.noise is not None:
    FMAP_NOISE = ARGS.noise
if ARGS.cycle is not None:
    CYCLE = ARGS.cycle
STEPS = ARGS.steps
BASE_LR, WARMUP_EPOCHS, STEPS_PER_EPOCH = 1e-3, 10, 75


def lr_at(t):
    return BASE_LR * float(t // STEPS_PER_EPOCH + 1) / WARMUP_EPOCHS


def batch(t):
    """Step t's batch: a fresh one (CYCLE = 0) or batch t % CYCLE of a fixed cycle."""
    seed = 1000 + (t % CYCLE if CYCLE else t)
    d = synth.make_batch(B, N_KPTS, seed=seed, imposter_every=0, with_kron=True, with_dense_gh=False,
                         fmap_noise=FMAP_NOISE)
    d.pop("label")
    return d


def run(dtype):
    torch.manual_seed(0)
    net = Net(regression=False)
    p = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone())
         for k, v in net.state_dict().items()}
    names = otrain.trainable_names(p)
    params = [p[k].requires_grad_(True) for k in names]
    opt = torch.optim.AdamW(params, lr=lr_at(0), weight_decay=1e-4)
    losses = []
    cache = {}
    for t in range(STEPS):
        key = t % CYCLE if CYCLE else t
        d = cache[key] if key in cache else batch(t)
        if CYCLE:
            cache[key] = d
        d = synth.clone_batch(d)
        for gp in opt.param_groups:
            gp["lr"] = lr_at(t)
        opt.zero_grad()
        fm = [(a.to(dtype), b.to(dtype)) for a, b in d["fmaps"]]
        out = head.forward_head(p, d, fm, regression=False, training=True, keep_graph=True, dtype=dtype)
        loss = otrain.permutation_loss(out["ds_mat"], d["gt_perm_mat"], d["ns"][0], d["ns"][1], dtype)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([q for q in params if q.grad is not None], max_norm=5.0)
        opt.step()
        losses.append(float(loss.detach()))
        if t % 10 == 0 or STEPS <= 30:
            print(t, losses[-1], flush=True)
    return losses


if __name__ == "__main__":
    name = "train_trajectory.json" if ARGS.config == "n14" else "train_trajectory_n100.json"
    out = Path(ARGS.out) if ARGS.out else ROOT / "tests" / "golden" / name
    rec = json.loads(out.read_text()) if out.exists() else {}
    rec.update({"B": B, "n": N_KPTS, "steps": STEPS, "seed_base": 1000, "fmap_noise": FMAP_NOISE, "cycle": CYCLE,
                "lr": BASE_LR,
                "warmup_epochs": WARMUP_EPOCHS, "steps_per_epoch": STEPS_PER_EPOCH,
                "weight_decay": 1e-4, "clip": 5.0, "torch": torch.__version__})
    t0 = time.time()
    if ARGS.fp64:
        rec["loss_fp64"] = run(torch.float64)
    else:
        rec["loss_fp32"] = run(torch.float32)
    rec["seconds_" + ("fp64" if ARGS.fp64 else "fp32")] = time.time() - t0
    out.write_text(json.dumps(rec))
    print("wrote", out)
