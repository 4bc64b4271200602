"""Pins oracle/cmf_oracle.py against the UNMODIFIED reference `cmf` (run here) and writes
tests/golden/cmf_c1.npz + cmf_state_dict.json.  TEST INFRASTRUCTURE.
Usage:  PYTHONPATH=oracle python oracle/gen_golden_cmf.py
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


def main():
    torch.set_num_threads(gc.GOLDEN_THREADS)
    get_model, _ = import_reference()
    import cmf_oracle as orc

    torch.manual_seed(gc.WEIGHT_SEED)
    ref = get_model("cmf").eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    left, right = gc.seeded_pair(1, 256, 512)
    hooked = {}
    ref.classif3.register_forward_hook(lambda m, i, o: hooked.__setitem__("c3", o.detach().clone().squeeze(1)))
    with torch.no_grad():
        ps = ref(left, right)
    stages = {}
    os_ = orc.forward(sd, left, right, 192, stages)
    checks = {"c3": (stages["c3"], hooked["c3"])}
    checks.update({"pred%d" % (i + 1): (a, b) for i, (a, b) in enumerate(zip(os_, ps))})
    for k, (mine, theirs) in checks.items():
        assert mine.shape == theirs.shape, (k, mine.shape, theirs.shape)
        d = (mine - theirs).abs().max().item()
        print("cmf oracle vs reference  %-6s max|diff| = %.3e  %s" % (k, d, "EXACT" if torch.equal(mine, theirs) else ""))
        assert d == 0.0
    contract = {"n_tensors": len(sd), "n_params": int(sum(v.numel() for v in sd.values())), "weight_seed": gc.WEIGHT_SEED,
                "tensors": [{"key": k, "shape": list(v.shape), "sum": float(v.double().sum()),
                             "abssum": float(v.double().abs().sum())} for k, v in sd.items()]}
    with open(os.path.join(OUT, "cmf_state_dict.json"), "w") as f:
        json.dump(contract, f, indent=0)
    np.savez_compressed(os.path.join(OUT, "cmf_c1.npz"),
                        **{"pred%d_sub" % (i + 1): p[0, 0, ::4, ::4].numpy() for i, p in enumerate(ps)})
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
