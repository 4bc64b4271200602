"""bench.py -- headline benchmark of the cmfsm hot path (BASELINE.json: pairs/s @540x960 D=192; cost-volume
HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one forward of the network over one synthetic 540x960 stereo pair (BASELINE config 2: fed as
576x960 exactly as the reference loader pads it, cmf/loader/Flying3d.py:67-72; fp32, maxdisp 192, B=1 per
GPU).  N>1: one process per GPU (torchrun), every rank runs its own pair -- the path shards by independent
pairs with no data-path collective ("weak" scaling); the timed region is bracketed by barrier + synchronize
and the max over ranks is reported.  Prints ONE JSON line on rank 0.

  value     whole-job pairs/s with the padded inputs already resident in HBM
  e2e       same metric through the public API `model(left, right)` from pinned HOST tensors, including the
            H2D copy of both images and the D2H read of the cropped disparity every step
  roofline  K1 cost-volume kernel (the kernel BASELINE.json's metric names): algorithmic bytes / CUDA-event
            launch duration vs the measured HBM peak of MEASURED_PEAKS.json
  kernels   per-kernel share of the step: the same K steps are repeated kernel-by-kernel (no graph) with a CUDA
            event pair around every launch on the launching stream; `roofline` uses those durations
  cpu_baseline  the CPU oracle port of the reference forward timed on this box's host cores (rank 0, N=1)

`--impl reference` times the reference's own CPU path (the oracle port: the reference is pure Python and
cannot travel to the GPU box) on the same config/metric.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200")
for _p in (PKG, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

H_IMG, W_IMG, H_PAD, MAXDISP = 540, 960, 576, 192
METRIC = "pairs/s @540x960 D=192"
WORKLOAD = "cmfsm inference, synthetic 540x960 pair fed as 576x960 (BASELINE config 2), maxdisp 192, B=1 per GPU"


def synthetic_pair(seed):
    """Uniform-random 540x960 pair padded to 576 rows the way the reference test loader does (last 36 rows)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    imgs = []
    for _ in range(2):
        x = torch.rand(1, 3, H_IMG, W_IMG, generator=g)
        imgs.append(torch.cat([x, x[:, :, -(H_PAD - H_IMG):]], 2).contiguous())
    return imgs


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.reasons.update(k for k, bit in names.items() if r & bit)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report it instead of failing the run
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def time_cpu_oracle(n_timed, budget_s=None):
    """Reference CPU path (oracle port) on the bench workload; returns (seconds per pair list, cores)."""
    import torch

    import cmfsm_oracle as orc
    from cmf.models import get_model

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in get_model("cmfsm").state_dict().items()}
    left, right = synthetic_pair(1)
    times, t_begin = [], time.perf_counter()
    for i in range(n_timed):
        t0 = time.perf_counter()
        out = orc.forward(sd, left, right, MAXDISP)[2][:, :, :H_IMG]
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    assert out.shape[-2:] == (H_IMG, W_IMG)
    return times, cores


def run_reference(args, rank, world):
    if rank != 0:
        return  # rank 0 alone runs the CPU reference arm
    times, cores = time_cpu_oracle(args.warmup + args.steps, budget_s=240.0)
    warm = min(args.warmup, max(0, len(times) - 1))
    timed = times[warm:]
    sec = sum(timed) / len(timed)
    value = 1.0 / sec
    sample = "%d full 576x960 pairs (after %d warm-up) through the CPU oracle port of cmfsm.forward" % (len(timed), warm)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": len(timed), "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "device": "host CPU, %d threads" % cores},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from cmf.models import get_model
    from cmf_b200 import lib, ops

    lib.load()  # fail loudly before anything else if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the cmfsm hot path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(0)
    model = get_model("cmfsm").to(dev).eval()
    model.aggregation = args.aggregation
    use_graph = not args.no_cuda_graph
    model.enable_cuda_graph(use_graph)
    left_h, right_h = (t.pin_memory() for t in synthetic_pair(1 + rank))
    left_d, right_d = left_h.to(dev), right_h.to(dev)

    def step_device():
        with torch.no_grad():
            return model(left_d, right_d)[2][:, :, :H_IMG]

    def step_e2e():
        with torch.no_grad():
            l = left_h.to(dev, non_blocking=True)
            r = right_h.to(dev, non_blocking=True)
            return model(l, r)[2][:, :, :H_IMG].contiguous().cpu()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ---------------- timed region: K steps, device-resident inputs (one CUDA-graph replay per step by default)
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert tuple(out.shape) == (1, 1, H_IMG, W_IMG) and bool(torch.isfinite(out).all())

    # ---------------- same K steps launched kernel by kernel with a CUDA event pair around every launch (on the
    # launching stream): per-kernel durations for the roofline / share report and the launch count
    model.enable_cuda_graph(False)
    for _ in range(3):  # the eager path uses the regular allocator pool: populate it before timing
        step_device()
    ops.enable_event_timing(True)
    n0 = lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    launches = (lib.launch_count() - n0) // args.steps
    ms_eager = ev0.elapsed_time(ev1)
    kernels = ops.drain_event_timing()
    ops.enable_event_timing(False)
    model.enable_cuda_graph(use_graph)

    # ---------------- end-to-end: pinned host inputs -> H2D -> forward -> D2H, every step
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert res.shape[-2:] == (H_IMG, W_IMG)

    t = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        hbm_peak, peak_src = peaks()
        h, w, D = H_PAD // 4, W_IMG // 4, MAXDISP // 4
        # read both fp32 feature maps once + write every voxel (4 B fp32 NCDHW, or 2 B in the C8/bf16 mode)
        s_out = 4 if args.aggregation == "fp32" else 2
        k1_name = "cost_volume_concat_fwd" if args.aggregation == "fp32" else "cost_volume_concat_c8_bf16"
        k1_bytes = 2 * 32 * h * w * 4 + 64 * D * h * w * s_out
        k1_n, k1_ms = kernels.get(k1_name, (0, 0.0))
        k1_traffic = None  # dram read+write bytes per launch of the fp32 kernel, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "k1_ncu_traffic.json")
        if args.aggregation == "fp32" and os.path.exists(tpath):
            k1_traffic = json.load(open(tpath))["traffic_bytes_per_launch"]
        k1_gbs = (k1_bytes * k1_n / (k1_ms * 1e-3) / 1e9) if k1_ms > 0 else None
        share = {k: {"launches": n // args.steps, "ms_per_step": ms / args.steps,
                     "share": ms / ms_eager} for k, (n, ms) in sorted(kernels.items())}
        precision = ("fp32 FMA 3-D aggregation (parity mode)" if args.aggregation == "fp32" else
                     "bf16-operand/fp32-accumulate tcgen05 implicit-GEMM 3-D aggregation (all 28 layers; 2-D features fp32 FMA)")
        ig_n, ig_ms = 0, 0.0
        for name in ("conv3d_igemm_bf16_fwd", "conv3d_s2_igemm_bf16_fwd", "deconv3d_igemm_bf16_fwd"):
            n_, ms_ = kernels.get(name, (0, 0.0))
            ig_n, ig_ms = ig_n + n_, ig_ms + ms_
        vox = D * h * w
        # stride-1 convs (dres0/1, conv2, conv4, classif.0) + per hourglass conv1/conv3 (stride 2) and
        # conv5/conv6 (transposed) = every 3-D layer except the three 32->1 classifier convs
        ig_macs = 27 * (vox * (64 * 32 + 6 * 32 * 32) + 3 * (vox // 8) * 64 * 64 + 3 * (vox // 64) * 64 * 64
                        + 3 * ((vox // 8) * 32 * 64 + (vox // 64) * 64 * 64)      # conv1, conv3
                        + 3 * ((vox // 64) * 64 * 64 + (vox // 8) * 64 * 32))     # conv5, conv6
        line = {"metric": METRIC, "value": world * args.steps / (ms_total * 1e-3), "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.aggregation == "fp32" else "bf16",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": "independent pairs, %d rank(s), no collective" % world,
                           "l2": "working set per step (425 MB cost volume, 212 MB activations) exceeds the 126 MB L2",
                           "precision": precision + "; fp32 FMA 2-D features, K1/K4/K5 fp32"},
                "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                        "h2d_bytes_per_step": 2 * 3 * H_PAD * W_IMG * 4, "d2h_bytes_per_step": H_IMG * W_IMG * 4},
                "gpu_launches": int(launches) * args.steps,
                "launch_mode": {"timed_region": "one CUDA-graph replay per step (%d kernel nodes of libcmfb200 + ATen "
                                                "cat / zero-fill / copy nodes)" % launches if use_graph else "eager launches",
                                "ms_per_step_eager_with_events": ms_eager / args.steps},
                "roofline": {"kernel": k1_name + " (K1)", "bound": "hbm", "achieved": k1_gbs,
                             "peak": hbm_peak, "unit": "GB/s", "frac": (k1_gbs / hbm_peak) if k1_gbs else None,
                             "traffic": k1_traffic, "traffic_source": "profiles/k1_ncu_traffic.json (ncu --set full)",
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes,
                             "us_per_launch": (k1_ms / k1_n * 1e3) if k1_n else None},
                "kernels": share, "clocks": sampler.summary()}
        if ig_n:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
                os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            tpeak = float(pk.get("bf16_tflops_sustained", 1400.0))
            tf = 2.0 * ig_macs * args.steps / (ig_ms * 1e-3) / 1e12
            line["roofline_k2"] = {"kernel": "tcgen05 implicit-GEMM kernels conv3d/conv3d_s2/deconv3d_igemm_bf16 (K2, %d launches/step)" % (ig_n // args.steps),
                                   "bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                                   "frac": tf / tpeak, "traffic": None,
                                   "peak_source": "bf16_tflops_sustained of MEASURED_PEAKS.json" if pk else "fallback",
                                   "algorithmic_flops_per_step": 2.0 * ig_macs}
            c1_n, c1_ms = kernels.get("conv3d_igemm_cout1_bf16_fwd", (0, 0.0))
            if c1_n:  # the three 32->1 classifier convs run the 32->32 schedule on zero-padded weights (1/32 useful)
                all_macs = ig_macs + 3 * 27 * vox * 32
                tf28 = 2.0 * all_macs * args.steps / ((ig_ms + c1_ms) * 1e-3) / 1e12
                line["roofline_k2"]["all_28_layers"] = {"achieved": tf28, "frac": tf28 / tpeak,
                                                        "algorithmic_flops_per_step": 2.0 * all_macs,
                                                        "launches": (ig_n + c1_n) // args.steps}
        if world == 1 and not args.no_cpu_baseline:
            times, cores = time_cpu_oracle(3)
            sec = statistics.median(times[1:]) if len(times) > 1 else times[0]
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": "median of %d full 576x960 pairs after 1 warm-up (CPU oracle port)"
                                              % (len(times) - 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--aggregation", default="fp32", choices=("fp32", "bf16"),
                    help="3-D aggregation arithmetic: fp32 FMA (BASELINE config 2, default) or bf16 tcgen05")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="launch the ~900 kernels of a forward one by one instead of replaying one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~30 s CPU oracle timing")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
