"""Golden outputs of the UNMODIFIED reference at the shapes the bench numbers are quoted on.

TEST INFRASTRUCTURE.  Run from the repo root in the build container (needs the reference tree, CPU, ~5 min):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_configs.py

  * config 2: 540x960 fed as 576x960 (cmf/loader/Flying3d.py:67-72), B=1, maxdisp 192, seed-0 weights,
    `golden_common.flying_padded_pair(1)` (uniform random, the bench input) ;
  * config 4 shape: 384x1248 (KITTI pad, cmf/loader/KITTI.py:100-108), two structured pairs (true disparity
    20 px), each run at B=1 (the reference's per-replica batch).
For every pair it stores the reference's three fp32 outputs and the fp64 oracle's outputs, sub-sampled ::8, and
asserts that the fp32 oracle restatement equals the reference exactly at that shape too.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import golden_common as gc  # noqa: E402
from ref_harness import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SUB = 8


def main():
    torch.set_num_threads(gc.GOLDEN_THREADS)
    get_model, _ = import_reference()
    import cmfsm_oracle as orc

    torch.manual_seed(gc.WEIGHT_SEED)
    ref = get_model("cmfsm").eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    sd64 = {k: v.double() for k, v in sd.items()}
    cases = {"c2": gc.flying_padded_pair(1), "c4_s1": gc.kitti_padded_pair(1), "c4_s2": gc.kitti_padded_pair(2)}
    out, meta = {}, {"sub": SUB, "torch": torch.__version__, "threads": gc.GOLDEN_THREADS, "cases": {}}
    for name, (left, right) in cases.items():
        with torch.no_grad():
            preds = ref(left, right)
            mine = orc.forward(sd, left, right, 192)
            p64 = orc.forward(sd64, left.double(), right.double(), 192)
        exact = all(torch.equal(a.reshape(b.shape), b) for a, b in zip(preds, mine))
        assert exact, "oracle restatement deviates from the reference at %s" % name
        dist = [[float((a.reshape(b.shape).double() - b).abs().max()), float((a.reshape(b.shape).double() - b).abs().mean())]
                for a, b in zip(preds, p64)]
        meta["cases"][name] = {"shape": list(left.shape), "oracle_equals_reference": exact,
                               "ref32_vs_fp64_max_mean": dist,
                               "pred3_mean": float(preds[2].double().mean())}
        print(name, tuple(left.shape), "oracle==reference:", exact, "ref32-vs-fp64 (max, mean):", dist, flush=True)
        for i, (a, b) in enumerate(zip(preds, p64), 1):
            out["%s_pred%d_ref32" % (name, i)] = a.reshape(b.shape)[0, 0, ::SUB, ::SUB].contiguous().numpy()
            out["%s_pred%d_fp64" % (name, i)] = b[0, 0, ::SUB, ::SUB].contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, "cmfsm_configs.npz"), **out)
    with open(os.path.join(OUT, "cmfsm_configs_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(os.path.getsize(os.path.join(OUT, "cmfsm_configs.npz")), "bytes")


if __name__ == "__main__":
    main()
