"""Times the non-headline BASELINE configs on one GPU (CUDA events, 3 warm-up + N timed forwards):
  C1  256x512  B1  fp32           (the reference's CPU-runnable case)
  C4  384x1248 B16 bf16 aggregation (KITTI-2015 shape)
  C5  2048x3072 B1 maxdisp 384 fp32 / bf16 (un-sharded; row bands: tools/check_row_bands.py)
  C2-sub_8  576x960 B1, the cmfsm_sub_8 variant
Prints one JSON line per config."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf.models.cmfsm import cmfsm  # noqa: E402
from cmf.models.cmfsm_sub_8 import cmfsm_sub_8  # noqa: E402
from cmf.models.cmfsm_sub_16 import cmfsm_sub_16  # noqa: E402


def run(name, B, H, W, maxdisp, agg, steps, cls=cmfsm):
    torch.manual_seed(0)
    model = cls(maxdisp=maxdisp).cuda().eval()
    model.aggregation = agg
    g = torch.Generator().manual_seed(1)
    left = torch.rand(B, 3, H, W, generator=g).cuda()
    right = torch.rand(B, 3, H, W, generator=g).cuda()
    with torch.no_grad():
        for _ in range(3):
            out = model(left, right)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = model(left, right)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert bool(torch.isfinite(out[2]).all()) and tuple(out[2].shape) in ((B, 1, H, W), (B, H, W))
    print(json.dumps({"config": name, "B": B, "H": H, "W": W, "maxdisp": maxdisp, "aggregation": agg,
                      "ms_per_forward": ms, "pairs_per_s": B / ms * 1e3,
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    del model, left, right, out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("C1", 1, 256, 512, 192, "fp32", 10)
    run("C1-bf16", 1, 256, 512, 192, "bf16", 10)
    run("C4", 16, 384, 1248, 192, "bf16", 3)
    run("C4-fp32", 16, 384, 1248, 192, "fp32", 2)
    run("C5", 1, 2048, 3072, 384, "fp32", 2)
    run("C5-bf16", 1, 2048, 3072, 384, "bf16", 2)
    run("C2-sub_8", 1, 576, 960, 192, "fp32", 10, cmfsm_sub_8)  # the 1/8-resolution variant on the config-2 pair
    run("C2-sub_8-bf16", 1, 576, 960, 192, "bf16", 10, cmfsm_sub_8)
    run("C2-sub_16", 1, 576, 960, 192, "fp32", 10, cmfsm_sub_16)  # 576x960 is a multiple of 64 as well
    run("C2-sub_16-bf16", 1, 576, 960, 192, "bf16", 10, cmfsm_sub_16)
