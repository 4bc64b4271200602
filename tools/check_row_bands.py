"""Row-band sharded inference check (GPU, torchrun): N ranks cooperate on ONE pair; the gathered result must match
the un-sharded forward of the same model (BASELINE config 5 mechanics at a test size).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tools/check_row_bands.py [--height 512 --width 512 --maxdisp 192]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf.models.cmfsm import cmfsm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--maxdisp", type=int, default=192)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the band forward as one CUDA graph per rank")
    ap.add_argument("--torch-profile", action="store_true", help="rank 0: kernel table of one band forward (torch.profiler)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = cmfsm(maxdisp=args.maxdisp).to(dev).eval()
    g = torch.Generator().manual_seed(5)
    left = torch.rand(1, 3, args.height, args.width, generator=g).to(dev)
    right = torch.rand(1, 3, args.height, args.width, generator=g).to(dev)
    if args.graph:
        model.enable_cuda_graph(True)
    model.forward_row_bands(left, right)  # warm-up (kernel loading, NCCL connections)
    outs = model.forward_row_bands(left, right)  # second warm-up + result
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(args.steps):
        model.forward_row_bands(left, right)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    per_rank = [float(ms)]
    if world > 1:
        gathered = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(gathered, ms)
        per_rank = [float(t) for t in gathered]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if args.torch_profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            model.forward_row_bands(left, right)
            torch.cuda.synchronize()
        if rank == 0:
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
            # idle time on the device: gaps between consecutive kernels / copies of the replayed forward
            ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                         if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
            busy = sum(b - a for a, b, _ in ev)
            gaps, t_end = [], ev[0][1]
            for i in range(1, len(ev)):
                if ev[i][0] > t_end:
                    gaps.append((ev[i][0] - t_end, ev[i - 1][2][:60], ev[i][2][:60]))
                t_end = max(t_end, ev[i][1])
            print("device span %.2f ms, busy %.2f ms, %d gaps totalling %.2f ms" %
                  ((t_end - ev[0][0]) / 1e3, busy / 1e3, len(gaps), sum(g[0] for g in gaps) / 1e3))
            by_next = {}
            for g, prev, nxt in gaps:
                k = (prev[:40], nxt[:40])
                by_next[k] = (by_next.get(k, (0, 0))[0] + g, by_next.get(k, (0, 0))[1] + 1)
            for k, (g, n) in sorted(by_next.items(), key=lambda kv: -kv[1][0])[:12]:
                print("   %8.1f us in %4d gaps   %s  ->  %s" % (g, n, k[0], k[1]))
    if args.profile:  # where does a band forward spend its time?  (our kernels by event pairs vs the elapsed time)
        from cmf_b200 import ops
        ops.enable_event_timing(True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        model.forward_row_bands(left, right)
        p1.record()
        kern = ops.drain_event_timing()
        ops.enable_event_timing(False)
        if rank == 0:
            tot = sum(t for _, t in kern.values())
            print("band forward with events: %.2f ms elapsed, %.2f ms inside libcmfb200 kernels (%d launches); rest = NCCL, "
                  "torch.cat / contiguous copies, idle" % (p0.elapsed_time(p1), tot, sum(n for n, _ in kern.values())))
            for k, (n, t) in sorted(kern.items(), key=lambda kv: -kv[1][1])[:8]:
                print("   %-28s %4d %.2f ms" % (k, n, t))
    if rank == 0:
        with torch.no_grad():
            ref = model(left, right)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(args.steps):
                model(left, right)
            t1.record()
            torch.cuda.synchronize()
        rep = {"world": world, "shape": [args.height, args.width], "maxdisp": args.maxdisp, "ms_sharded": float(ms), "ms_per_rank": per_rank,
               "ms_single_gpu": t0.elapsed_time(t1) / args.steps}
        for i, (a, b) in enumerate(zip(outs, ref), 1):
            d = (a - b).abs()
            rep["pred%d_max_abs" % i], rep["pred%d_mean_abs" % i] = float(d.max()), float(d.mean())
            # same fp32 arithmetic per voxel; only GroupNorm summation order differs -> fp32 noise floor of the net
            assert float(d.max()) < 3e-2 and float(d.mean()) < 2e-3, rep
        print(json.dumps(rep), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
