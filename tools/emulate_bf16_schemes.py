"""CPU experiment (test infrastructure): which bf16 storage/rounding scheme of the 3-D aggregation meets the
0.02 px |EPE_bf16 - EPE_fp32| bar?  Emulates operand rounding on the CPU oracle (fp32 accumulate).

    python tools/emulate_bf16_schemes.py [H W]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
import cmfsm_oracle as orc  # noqa: E402
import golden_common as gc  # noqa: E402


def r(x):
    return x.to(torch.bfloat16).to(torch.float32)


def r2(x):  # two-term bf16 split (hi + mid) = 16 significand bits
    hi = r(x)
    return hi + r(x - hi)


class Scheme:
    def __init__(self, name, raw_round=True, hi_layers=(), act=r, wt=r):
        self.name, self.raw_round, self.hi_layers, self.act, self.wt = name, raw_round, hi_layers, act, wt


def make_patched(s):
    def is_hi(key):
        return any(key.startswith(h) for h in s.hi_layers)

    def convgn3d(sd, key, x, stride=1):
        hi = is_hi(key)
        a = r2 if hi else s.act
        w = r2 if hi else s.wt
        y = F.conv3d(a(x), w(sd[key + ".0.weight"]), None, stride, 1)
        if s.raw_round and not hi:
            y = r(y)
        return orc._gn(sd, key + ".1", y)

    def deconvgn3d(sd, key, x):
        hi = is_hi(key)
        a = r2 if hi else s.act
        w = r2 if hi else s.wt
        y = F.conv_transpose3d(a(x), w(sd[key + ".0.weight"]), None, stride=2, padding=1, output_padding=1)
        if s.raw_round and not hi:
            y = r(y)
        return orc._gn(sd, key + ".1", y)

    def classif(sd, key, x):
        x = F.relu(convgn3d(sd, key + ".0", x))
        hi = is_hi(key + ".2")
        a = (lambda t: t) if hi and os.environ.get("ONLY_C2") else r2 if hi else s.act
        w = (lambda t: t) if hi and os.environ.get("ONLY_C2") else r2 if hi else s.wt
        return F.conv3d(a(x), w(sd[key + ".2.weight"]), None, 1, 1).squeeze(1)

    return convgn3d, deconvgn3d, classif


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 512)
    torch.set_num_threads(os.cpu_count())
    from cmf.models.cmfsm import cmfsm  # parameter container only (CPU init)

    torch.manual_seed(gc.WEIGHT_SEED)
    sd = {k: v.detach() for k, v in cmfsm().state_dict().items()}
    seeds = [int(v) for v in os.environ.get("SEEDS", "1,2").split(",")]
    totals = {}
    for seed in seeds:
        left, right = gc.structured_pair(H, W, delta=20, seed=seed)
        with torch.no_grad():
            L, all_l = orc.feature_extraction(sd, left)
            R, _ = orc.feature_extraction(sd, right)
            weights = orc.context_mapping_weights(sd, L, all_l)
            cost = orc.cost_volume_concat(L, R, 48)
            ref = orc.softargmin_ctxmap(*orc.aggregation3d(sd, cost), weights, 4)
            epe_ref = [float((o - 20.0).abs().mean()) for o in ref]
            ident = lambda t: t  # noqa: E731
            schemes = [Scheme("double + classif.2 fp32 (FFMA, GN out of classif.0 kept fp32)", True, ("classif1.2", "classif2.2", "classif3.2")),
                       Scheme("double (raw conv out bf16, GN out bf16) = round 1", True)] if os.environ.get("ONLY_C2") else [
                Scheme("double (raw conv out bf16, GN out bf16) = round 1", True),
                Scheme("single (raw fp32, GN out bf16)", False),
                Scheme("single + classif.2 hi", False, ("classif1.2", "classif2.2", "classif3.2")),
                Scheme("single + classif hi", False, ("classif",)),
                Scheme("single + dres0/1 hi", False, ("dres0", "dres1")),
                Scheme("single + dres0/1 + classif hi", False, ("dres0", "dres1", "classif")),
                Scheme("double + classif hi", True, ("classif",)),
            ]
            saved = (orc._convgn3d, orc._deconvgn3d, orc.classif)
            for s in schemes:
                orc._convgn3d, orc._deconvgn3d, orc.classif = make_patched(s)
                got = orc.softargmin_ctxmap(*orc.aggregation3d(sd, r(cost)), weights, 4)
                orc._convgn3d, orc._deconvgn3d, orc.classif = saved
                line = []
                for i, (a, b, e) in enumerate(zip(got, ref, epe_ref)):
                    ea = float((a - 20.0).abs().mean())
                    line.append("mean|d| %.3f dEPE %+.4f" % (float((a - b).abs().mean()), ea - e))
                    totals.setdefault(s.name, [0.0, 0.0, 0.0])[i] += (ea - e) / len(seeds)
                print("seed %d %-52s %s" % (seed, s.name, " | ".join(line)), flush=True)
    for name, t in totals.items():
        print("DATASET dEPE over %d pairs  %-52s %+.4f %+.4f %+.4f" % (len(seeds), name, *t), flush=True)


if __name__ == "__main__":
    main()
