"""Data-parallel training check (GPU, torchrun): N ranks x per-rank batch vs ONE process on the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_dp_train.py [--per-rank-batch 1] [--steps 2]

Every rank builds the same seed-0 `cmfsm`, takes its shard of a seeded global batch (256x512 crops, synthetic
disparity), runs `cmf_b200.parallel.dp_train_step` (forward + backward + NCCL gradient all-reduce + Adam).  Rank 0
then repeats the same steps alone on the WHOLE batch and compares losses and updated weights: the sharded run must
reproduce the single-process DataParallel semantics of the reference (train.py:78-81,159-181).  Also prints the
step time (CUDA events, max over ranks).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf.models import get_model  # noqa: E402
from cmf_b200 import parallel as par  # noqa: E402


def make_batch(n, seed=3):
    g = torch.Generator().manual_seed(seed)
    left, right = torch.rand(n, 3, 256, 512, generator=g), torch.rand(n, 3, 256, 512, generator=g)
    disp = torch.rand(n, 256, 512, generator=g) * 230 - 10  # ~17 % of the pixels fall outside (0, 192)
    return left, right, disp


def run_steps(model, opt, left, right, disp, steps):
    losses = []
    for _ in range(steps):
        losses.append(par.dp_train_step(model, opt, left, right, disp))
    return losses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-rank-batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nb = args.per_rank_batch
    left, right, disp = make_batch(nb * world)

    torch.manual_seed(0)
    model = get_model("cmfsm").to(dev).train()
    par.broadcast_parameters(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))  # train.py:85-86
    sl = slice(rank * nb, (rank + 1) * nb)
    shard = [t[sl].to(dev) for t in (left, right, disp)]
    run_steps(model, opt, *shard, 1)  # warm-up (also exercises a second optimiser state)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    losses = run_steps(model, opt, *shard, args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    report = {"world": world, "per_rank_batch": nb, "ms_per_step": float(ms), "pairs_per_s": nb * world / float(ms) * 1e3,
              "losses": losses}

    # ---- gradient equivalence: sharded (all-reduced) gradients vs one process on the whole batch, same weights
    def grads_of(net, l, r, d):
        net.zero_grad(set_to_none=True)
        loss = par.masked_smooth_l1_dp(net(l, r), d, 192)
        loss.backward()
        return loss.detach()

    torch.manual_seed(0)
    probe = get_model("cmfsm").to(dev).train()
    loss_s = grads_of(probe, *shard)
    par.allreduce_gradients(list(probe.parameters()), average=False)
    if world > 1:
        dist.all_reduce(loss_s)
    g_sharded = torch.cat([p.grad.reshape(-1) for p in probe.parameters()]).clone()
    if rank == 0 and world > 1:
        dist_world = par.world
        par.world = lambda group=None: 1  # single-process semantics for the reference pass
        try:
            loss_r = grads_of(probe, *[t.to(dev) for t in (left, right, disp)])
        finally:
            par.world = dist_world
        g_ref = torch.cat([p.grad.reshape(-1) for p in probe.parameters()])
        rel = float((g_sharded - g_ref).norm() / g_ref.norm())
        report.update(grad_rel_l2=rel, loss_sharded=float(loss_s), loss_single=float(loss_r))
        # same arithmetic per sample; differences come from fp32 summation order only (atomics, batch-size
        # dependent cuDNN algorithms), amplified by this network's random-init chaos
        assert abs(float(loss_s) - float(loss_r)) <= 1e-3 * abs(float(loss_r)), report
        assert rel < 2e-2, report
    if rank == 0:
        print(json.dumps(report), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
