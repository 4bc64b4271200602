"""Pins oracle/cm_sub_oracle.py against the UNMODIFIED reference `cm_sub_4` / `cm_sub_8` / `cm_sub_16` (run here) and
writes tests/golden/cm_sub{4,8,16}_c1.npz + cm_sub{4,8,16}_state_dict.json.  TEST INFRASTRUCTURE.
Usage:  PYTHONPATH=oracle python oracle/gen_golden_cm_sub.py
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
    import cm_sub_oracle as orc

    for variant in ("4", "8", "16"):
        name = "cm_sub_" + variant
        torch.manual_seed(gc.WEIGHT_SEED)
        ref = get_model(name).eval()
        sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        left, right = gc.seeded_pair(1, 256, 512)
        hooked = {}
        ref.classif1.register_forward_hook(lambda m, i, o: hooked.__setitem__("c1", o.detach().clone().squeeze(1)))
        with torch.no_grad():
            p1, p2, p3 = ref(left, right)
        assert p1 is p2 and p2 is p3
        stages = {}
        o1, _, _ = orc.forward(sd, left, right, variant, 192, stages)
        for k, (mine, theirs) in {"c1": (stages["c1"], hooked["c1"]), "pred1": (o1, p1)}.items():
            d = (mine - theirs).abs().max().item()
            print("%s oracle vs reference  %-6s max|diff| = %.3e  %s" % (name, k, d, "EXACT" if torch.equal(mine, theirs) else ""))
            assert d == 0.0
        contract = {"n_tensors": len(sd), "n_params": int(sum(v.numel() for v in sd.values())), "weight_seed": gc.WEIGHT_SEED,
                    "tensors": [{"key": k, "shape": list(v.shape), "sum": float(v.double().sum()),
                                 "abssum": float(v.double().abs().sum())} for k, v in sd.items()]}
        with open(os.path.join(OUT, "cm_sub%s_state_dict.json" % variant), "w") as f:
            json.dump(contract, f, indent=0)
        np.savez_compressed(os.path.join(OUT, "cm_sub%s_c1.npz" % variant), pred1_sub=p1[0, ::4, ::4].numpy(),
                            c1_sub=stages["c1"][0, ::2, ::2, ::2].numpy())
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
