"""Pins oracle/cmfsm_sub16_oracle.py against the UNMODIFIED reference `cmfsm_sub_16` (run here, in the build container)
and writes the fixtures of the 1/8-resolution variant:
    tests/golden/cmfsm_sub16_c1.npz            sub-sampled stages + outputs of the real reference at 256x512, seed-0 init
    tests/golden/cmfsm_sub16_state_dict.json   keys / shapes / checksums of the seed-0 state_dict (init + key contract)
TEST INFRASTRUCTURE.  Usage:  PYTHONPATH=oracle python oracle/gen_golden_sub16.py
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
    import cmfsm_sub16_oracle as orc8

    torch.manual_seed(gc.WEIGHT_SEED)
    ref = get_model("cmfsm_sub_16").eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    left, right = gc.seeded_pair(1, 256, 512)

    hooked, fe_calls = {}, []
    ref.feature_extraction.register_forward_hook(lambda m, i, o: fe_calls.append(o))
    ref.mapping_matrix.register_forward_hook(lambda m, i, o: hooked.update(weights=torch.cat(o[0], 1), w3=torch.cat(o[1], 1)))
    ref.dres0.register_forward_hook(lambda m, i, o: hooked.__setitem__("cost", i[0].detach().clone()))
    for name in ("classif1", "classif2", "classif3"):
        getattr(ref, name).register_forward_hook(
            lambda m, i, o, name=name: hooked.__setitem__(name, o.detach().clone().squeeze(1)))
    with torch.no_grad():
        p1, p2, p3 = ref(left, right)

    stages = {}
    o1, o2, o3 = orc8.forward(sd, left, right, 192, stages)
    checks = {"L": (stages["L"], fe_calls[0][0]), "all_l": (stages["all_l"], fe_calls[0][2]),
              "R": (stages["R"], fe_calls[1][0]), "weights": (stages["w5"], hooked["weights"]), "w3": (stages["w3"], hooked["w3"]),
              "cost": (stages["cost"], hooked["cost"]), "c1": (stages["c1"], hooked["classif1"]),
              "c2": (stages["c2"], hooked["classif2"]), "c3": (stages["c3"], hooked["classif3"]),
              "pred1": (o1, p1), "pred2": (o2, p2), "pred3": (o3, p3)}  # [B,H,W] each
    report = {}
    for k, (mine, theirs) in checks.items():
        assert mine.shape == theirs.shape, (k, mine.shape, theirs.shape)
        d = (mine - theirs).abs().max().item()
        report[k] = d
        print("sub_16 oracle vs reference  %-8s max|diff| = %.3e  %s" % (k, d, "EXACT" if torch.equal(mine, theirs) else ""))
    assert all(v == 0.0 for v in report.values()), "sub_16 oracle deviates from the reference: %r" % report

    contract = {"n_tensors": len(sd), "n_params": int(sum(v.numel() for v in sd.values())), "weight_seed": gc.WEIGHT_SEED,
                "tensors": [{"key": k, "shape": list(v.shape), "sum": float(v.double().sum()),
                             "abssum": float(v.double().abs().sum())} for k, v in sd.items()]}
    with open(os.path.join(OUT, "cmfsm_sub16_state_dict.json"), "w") as f:
        json.dump(contract, f, indent=0)
    arrays = {"pred1_sub": p1[0, ::4, ::4], "pred2_sub": p2[0, ::4, ::4], "pred3_sub": p3[0, ::4, ::4],
              "L_sub": stages["L"][0, ::4], "all_l_sub": stages["all_l"][0, ::8, ::8, ::8],
              "weights_sub": stages["w5"][0, :, ::8, ::8], "w3_sub": stages["w3"][0, :, ::8, ::8],
              "c1_sub": stages["c1"][0], "c3_sub": stages["c3"][0]}
    np.savez_compressed(os.path.join(OUT, "cmfsm_sub16_c1.npz"), **{k: v.numpy() for k, v in arrays.items()})
    meta = {"oracle_vs_reference_max_abs": report,
            "stats": {k: [float(v.double().mean()), float(v.double().abs().mean()), float(v.min()), float(v.max())]
                      for k, v in dict(pred1=p1, pred2=p2, pred3=p3, weights=stages["w5"], w3=stages["w3"], c1=stages["c1"]).items()}}
    with open(os.path.join(OUT, "cmfsm_sub16_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
