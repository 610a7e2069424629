"""N-rank check of the training path's gradient all-reduce (run under torchrun on 2..8 GPUs):
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


def run(net, batches, dev, steps, grad_probe=None):
    params = trainable(net.module if hasattr(net, "module") else net)
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)      # stage-1 warm-up LR (train.py:246-252,297)
    losses = []
    for t in range(steps):
        d = synth.batch_to(synth.clone_batch(batches[t % len(batches)]), dev)
        opt.zero_grad()
        out = net(d)
        loss = loss_fn(out, d)
        loss.backward()
        if grad_probe is not None and t == 0:        # first-step gradient magnitude of every tensor (see main)
            mod = net.module if hasattr(net, "module") else net
            for k, p in mod.named_parameters():
                if p.grad is not None:
                    grad_probe[k] = p.grad.abs().max().item()
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
        b = synth.make_batch(max(4, 2 * world), 16, seed=50 + i, imposter_every=0, with_kron=False)
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
        gmax = {}
        l_ref = run(ref, glob, dev, steps, grad_probe=gmax)
        # Parameter comparison.  AdamW moves every coordinate by ~lr per step whatever the size of its gradient, so a
        # coordinate whose TRUE gradient is zero (the loss sees the GNN scores only through shift-invariant Sinkhorn
        # layers: biases, and the weight directions that only shift a layer's scores) takes +-lr steps along fp32
        # ROUNDING NOISE, which differs between a sharded and a whole-batch run.  Such tensors can differ by up to
        # steps * 2 * lr in absolute terms - a large fraction of a small tensor - without any error in the all-reduce.
        # Reported: the worst relative difference over all weight tensors (with its name and absolute size), and the
        # worst over the well-conditioned ones (first-step gradient above 1e-4 of the largest gradient).
        gtop = max(gmax.values()) if gmax else 1.0
        worst, worst_name, worst_abs, worst_cond, worst_cond_name = 0.0, None, 0.0, 0.0, None
        for (k, a), (_, b) in zip(net.named_parameters(), ref.named_parameters()):
            if k.endswith("bias") or k not in gmax:
                continue
            d_abs = (a - b).abs().max().item()
            r = d_abs / max(b.abs().max().item(), 1e-12)
            if r > worst:
                worst, worst_name, worst_abs = r, k, d_abs
            if gmax[k] > 1e-4 * gtop and r > worst_cond:
                worst_cond, worst_cond_name = r, k
        rel = max(abs(a - b) / abs(b) for a, b in zip(l_mean, l_ref))
        rec = {"world": world, "steps": steps, "global_batch": glob[0]["gt_perm_mat"].shape[0],
               "loss_ddp_mean": l_mean, "loss_single": l_ref, "loss_rel_max": rel,
               "param_rel_max": worst, "param_rel_max_tensor": worst_name, "param_abs_diff_of_that_tensor": worst_abs,
               "adam_noise_bound_abs (steps * 2 * lr)": steps * 2 * 1e-4,
               "first_step_grad_max_of_that_tensor": gmax.get(worst_name), "largest_first_step_grad": gtop,
               "param_rel_max_well_conditioned": worst_cond, "param_rel_max_well_conditioned_tensor": worst_cond_name}
        print(json.dumps(rec))
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "ddp_train_check.json").write_text(json.dumps(rec))
        ok = rel < 1e-3
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
