"""Generate tests/golden/* by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE.  Run from the repo root (needs /root/reference, CPU only, ~2 min):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

What it does
  1. imports the reference through oracle/ref_harness.py (two shims);
  2. runs `get_model('cmfsm')` (seed 0 weights, seed 1 inputs, 256x512, B=1) with forward hooks and
     asserts that oracle/cmfsm_oracle.py reproduces every hooked stage EXACTLY (same ATen ops, same
     order) -- this is the pinning of the oracle against the reference itself;
  3. runs individual reference modules (hourglass, convbn_3d, eight_related_context_mapping, ...) on
     small seeded inputs with weights from `golden_common.seeded_weights` and stores their outputs;
  4. writes small fixtures (sub-sampled tensors, crops, sha256 digests, per-tensor weight checksums).

The fixtures are committed; tests never need the reference tree.
"""
import hashlib
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


def sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(gc.GOLDEN_THREADS)
    get_model, ref_mod = import_reference()
    import cmfsm_oracle as orc

    # ------------------------------------------------------------------ full network, config C1
    torch.manual_seed(gc.WEIGHT_SEED)
    ref = get_model("cmfsm").eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    left, right = gc.seeded_pair(1, 256, 512)

    hooked = {}
    fe_calls = []
    ref.feature_extraction.register_forward_hook(lambda m, i, o: fe_calls.append(o))
    ref.mapping_matrix.register_forward_hook(lambda m, i, o: hooked.__setitem__("weights", torch.cat(o, 1)))
    ref.dres0.register_forward_hook(lambda m, i, o: hooked.__setitem__("cost", i[0].detach().clone()))
    for name in ("classif1", "classif2", "classif3"):
        getattr(ref, name).register_forward_hook(
            lambda m, i, o, name=name: hooked.__setitem__(name, o.detach().clone().squeeze(1)))
    for name in ("dres2", "dres3", "dres4"):
        getattr(ref, name).register_forward_hook(
            lambda m, i, o, name=name: hooked.__setitem__(name, [t.detach().clone() for t in o]))
    with torch.no_grad():
        p1, p2, p3 = ref(left, right)

    stages = {}
    o1, o2, o3 = orc.forward(sd, left, right, 192, stages)
    checks = {
        "L": (stages["L"], fe_calls[0][0]), "all_l": (stages["all_l"], fe_calls[0][2]),
        "R": (stages["R"], fe_calls[1][0]), "weights": (stages["weights"], hooked["weights"]),
        "cost": (stages["cost"], hooked["cost"]),
        "c1": (stages["c1"], hooked["classif1"]), "c2": (stages["c2"], hooked["classif2"]),
        "c3": (stages["c3"], hooked["classif3"]),
        "pre1": (stages["pre1"], hooked["dres2"][1]), "post2": (stages["post2"], hooked["dres3"][2]),
        "pred1": (o1, p1), "pred2": (o2, p2), "pred3": (o3, p3),
    }
    report = {}
    for k, (mine, theirs) in checks.items():
        d = (mine - theirs).abs().max().item()
        report[k] = d
        print("oracle vs reference  %-8s max|diff| = %.3e  %s" % (k, d, "EXACT" if torch.equal(mine, theirs) else ""))
    assert torch.equal(stages["cost"], hooked["cost"]), "cost volume restatement is not bit-exact"
    assert all(v == 0.0 for v in report.values()), "oracle restatement deviates from the reference: %r" % report

    # state-dict contract + per-tensor checksums of the seed-0 initialisation
    contract = {"n_tensors": len(sd), "n_params": int(sum(v.numel() for v in sd.values())),
                "weight_seed": gc.WEIGHT_SEED,
                "tensors": [{"key": k, "shape": list(v.shape), "sum": float(v.double().sum()),
                             "abssum": float(v.double().abs().sum())} for k, v in sd.items()]}
    with open(os.path.join(OUT, "cmfsm_state_dict.json"), "w") as f:
        json.dump(contract, f, indent=0)

    full = {"pred1_sub": p1[0, 0, ::4, ::4], "pred2_sub": p2[0, 0, ::4, ::4], "pred3_sub": p3[0, 0, ::4, ::4],
            "L_sub": stages["L"][0, ::4, ::4, ::4], "R_sub": stages["R"][0, ::4, ::4, ::4],
            "all_l_sub": stages["all_l"][0, ::8, ::8, ::8], "weights_sub": stages["weights"][0, :, ::8, ::8],
            "c1_sub": stages["c1"][0, ::4, ::4, ::4], "c3_sub": stages["c3"][0, ::4, ::4, ::4],
            "cost0_sub": stages["cost0"][0, ::4, ::4, ::4, ::8], "out3_sub": stages["out3"][0, ::4, ::4, ::4, ::8]}
    stats = {k: [float(v.double().mean()), float(v.double().abs().mean()), float(v.min()), float(v.max())]
             for k, v in dict(pred1=p1, pred2=p2, pred3=p3, L=stages["L"], R=stages["R"], weights=stages["weights"],
                              cost=stages["cost"], cost0=stages["cost0"], out1=stages["out1"], out3=stages["out3"],
                              c1=stages["c1"], c2=stages["c2"], c3=stages["c3"]).items()}
    # K1: crop of the real features and the matching crop of the real cost volume + digest of the whole volume
    Lr, Rr, cost = fe_calls[0][0], fe_calls[1][0], hooked["cost"]
    cs, rs = gc.K1_CROP_CH, gc.K1_CROP_ROWS
    k1 = {"k1_L_crop": Lr[0, :cs, :rs], "k1_R_crop": Rr[0, :cs, :rs],
          "k1_cost_crop": torch.cat([cost[0, :cs, :, :rs], cost[0, 32:32 + cs, :, :rs]], 0)}
    # K4: crops of the real classifier volumes / weights, and the real outputs on the crop interior
    y0, y1, x0, x1 = gc.K4_CROP  # low-res cell window
    k4 = {"k4_c1": hooked["classif1"][0, :, y0:y1, x0:x1], "k4_c2": hooked["classif2"][0, :, y0:y1, x0:x1],
          "k4_c3": hooked["classif3"][0, :, y0:y1, x0:x1],
          "k4_w": hooked["weights"][0, :, 4 * y0:4 * y1, 4 * x0:4 * x1]}
    for i, p in enumerate((p1, p2, p3)):
        k4["k4_out%d_interior" % (i + 1)] = p[0, 0, 4 * (y0 + 1):4 * (y1 - 1), 4 * (x0 + 1):4 * (x1 - 1)]
    np.savez_compressed(os.path.join(OUT, "cmfsm_c1_full.npz"),
                        **{k: v.contiguous().numpy() for k, v in {**full, **k1, **k4}.items()})
    meta = {"config": "cmfsm 256x512 B1 maxdisp192, weights seed %d (reference init), inputs seeded_pair(1,256,512)" % gc.WEIGHT_SEED,
            "torch": torch.__version__, "threads": gc.GOLDEN_THREADS, "stats": stats,
            "sha256": {"cost": sha(cost), "L": sha(Lr), "R": sha(Rr)},
            "oracle_vs_reference_maxabs": report}

    # ------------------------------------------------------------------ per-module fixtures (small)
    mods = {}
    with torch.no_grad():
        # K5: the reference mapping module on small features, B=2
        mm = ref_mod.eight_related_context_mapping().eval()
        w = gc.seeded_weights(mm.state_dict(), gc.SEED_K5_W)
        mm.load_state_dict(w)
        lr, hr = gc.k5_inputs()
        out = torch.cat(mm(lr, hr, lr, hr), 1)
        mine = orc.context_mapping_weights({"mapping_matrix." + k: v for k, v in w.items()}, lr, hr)
        assert torch.equal(out, mine), (out - mine).abs().max()
        mods["k5_weights"] = out

        # K2/K3: hourglass with and without skip inputs
        hg = ref_mod.hourglass(32).eval()
        w = gc.seeded_weights(hg.state_dict(), gc.SEED_HG_W, gn_affine=True)
        hg.load_state_dict(w)
        x, presqu, postsqu = gc.hourglass_inputs()
        a = hg(x, None, None)
        b = hg(x, presqu, postsqu)
        sdh = {"hg." + k: v for k, v in w.items()}
        a2 = orc.hourglass(sdh, "hg", x, None, None)
        b2 = orc.hourglass(sdh, "hg", x, presqu, postsqu)
        for u, v in zip(a + b, a2 + b2):
            assert torch.equal(u, v)
        mods.update(hg_out_a=a[0], hg_pre_a=a[1], hg_post_a=a[2], hg_out_b=b[0], hg_pre_b=b[1], hg_post_b=b[2])

        # K2/K3: dres0-style head (64->32 +GN+ReLU, 32->32 +GN+ReLU) and classifier (32->32+GN+ReLU, 32->1)
        head = torch.nn.Sequential(ref_mod.convbn_3d(64, 32, 3, 1, 1), torch.nn.ReLU(inplace=True),
                                   ref_mod.convbn_3d(32, 32, 3, 1, 1), torch.nn.ReLU(inplace=True)).eval()
        w = gc.seeded_weights(head.state_dict(), gc.SEED_HEAD_W, gn_affine=True)
        head.load_state_dict(w)
        xc = gc.head_input()
        mods["head_out"] = head(xc)
        cls = torch.nn.Sequential(ref_mod.convbn_3d(32, 32, 3, 1, 1), torch.nn.ReLU(inplace=True),
                                  torch.nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False)).eval()
        w = gc.seeded_weights(cls.state_dict(), gc.SEED_CLS_W, gn_affine=True)
        cls.load_state_dict(w)
        mods["cls_out"] = cls(gc.classif_input())
    np.savez_compressed(os.path.join(OUT, "cmfsm_modules.npz"), **{k: v.contiguous().numpy() for k, v in mods.items()})
    with open(os.path.join(OUT, "cmfsm_c1_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for fn in sorted(os.listdir(OUT)):
        print("%9d  %s" % (os.path.getsize(os.path.join(OUT, fn)), fn))


if __name__ == "__main__":
    main()
