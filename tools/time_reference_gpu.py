"""GPU-eager baseline: the UNMODIFIED reference `cmfsm` (staged under oracle/_ref/reference) on the B200 through
PyTorch/cuDNN, strict fp32 (TF32 off) and PyTorch-default TF32, BASELINE config 2 (576x960).  TEST INFRASTRUCTURE /
measurement only: this process imports the reference's `cmf`, never ours.  Prints one JSON line.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ref_harness import import_reference  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    get_model, _ = import_reference()
    dev = "cuda:0"
    torch.manual_seed(0)
    model = get_model("cmfsm").to(dev).eval()
    g = torch.Generator().manual_seed(1)
    imgs = []
    for _ in range(2):
        x = torch.rand(1, 3, 540, 960, generator=g)
        imgs.append(torch.cat([x, x[:, :, -36:]], 2).contiguous().to(dev))
    out = {"what": "reference cmfsm.forward, PyTorch eager + cuDNN on cuda:0, 576x960 B=1", "steps": steps}
    preds = {}
    for label, tf32 in (("fp32_tf32_off", False), ("tf32_default", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            for _ in range(2):
                p = model(*imgs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                p = model(*imgs)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        preds[label] = p[2].float()
        out[label] = {"ms_per_pair": ms, "pairs_per_s": 1e3 / ms}
    d = (preds["fp32_tf32_off"] - preds["tf32_default"]).abs()
    out["tf32_vs_fp32_pred3_px"] = {"max": float(d.max()), "mean": float(d.mean())}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
