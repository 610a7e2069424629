"""2-rank check of the training path's gradient all-reduce (run under torchrun on 2 GPUs):
every rank trains its shard of a global batch under DistributedDataParallel for a few AdamW steps; rank 0 also
trains a replica on the WHOLE batch in a single process and compares the parameters.  With equal keypoint counts
per pair the mean over ranks of the per-shard PermutationLoss equals the whole-batch loss, so the two runs must
agree to fp32 rounding.   torchrun --nproc-per-node 2 tools/ddp_train_check.py"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fpmatch import dist as fdist, synth  # noqa: E402
from src.model.ngm import Net  # noqa: E402


def loss_fn(out, d):
    ds, gt, n1, n2 = out["ds_mat"], d["gt_perm_mat"], d["ns"][0], d["ns"][1]
    B, R, C = ds.shape
    mask = (torch.arange(R, device=ds.device)[None, :, None] < n1.view(B, 1, 1)) & \
           (torch.arange(C, device=ds.device)[None, None, :] < n2.view(B, 1, 1))
    return (torch.nn.functional.binary_cross_entropy(ds, gt, reduction="none") * mask).sum() / n1.sum().float()


def trainable(net):
    frozen = ("encoder_k.", "final_row.", "final_col.", "match_cls.", "node_layers.", "edge_layers.")
    return [p for k, p in net.named_parameters() if not k.startswith(frozen)]


def run(net, batches, dev, steps):
    params = trainable(net.module if hasattr(net, "module") else net)
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)      # stage-1 warm-up LR (train.py:246-252,297)
    losses = []
    for t in range(steps):
        d = synth.batch_to(synth.clone_batch(batches[t % len(batches)]), dev)
        opt.zero_grad()
        out = net(d)
        loss = loss_fn(out, d)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
        opt.step()
        losses.append(loss.item())
    return losses


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    steps = 5
    glob = []
    for i in range(2):
        b = synth.make_batch(4, 16, seed=50 + i, imposter_every=0, with_kron=False)
        b.pop("label")
        glob.append(b)
    torch.manual_seed(0)
    net = Net(regression=False).to(dev).train()
    for k, p in net.named_parameters():
        if k.startswith(("encoder_k.", "final_row.", "final_col.", "match_cls.")):
            p.requires_grad_(False)
    ddp = fdist.wrap_ddp(net, dev)
    shard = [fdist.shard_batch(b, rank, world) for b in glob]
    l_ddp = run(ddp, shard, dev, steps)
    t = torch.tensor(l_ddp, device=dev)
    dist.all_reduce(t)
    l_mean = (t / world).tolist()
    ok = True
    if rank == 0:
        torch.manual_seed(0)
        ref = Net(regression=False).to(dev).train()
        l_ref = run(ref, glob, dev, steps)
        worst = 0.0        # weights only: the GNN biases have zero true gradient, Adam turns their fp32 noise into +-lr steps
        for (k, a), (_, b) in zip(net.named_parameters(), ref.named_parameters()):
            if not k.endswith("bias"):
                worst = max(worst, (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12))
        rel = max(abs(a - b) / abs(b) for a, b in zip(l_mean, l_ref))
        rec = {"world": world, "steps": steps, "loss_ddp_mean": l_mean, "loss_single": l_ref, "loss_rel_max": rel,
               "param_rel_max": worst}
        print(json.dumps(rec))
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "ddp_train_check.json").write_text(json.dumps(rec))
        ok = rel < 1e-3
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
