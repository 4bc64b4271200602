"""CPU experiment (test infrastructure): GPU-faithful emulation of the bf16 aggregation mode (every stored activation
bf16, raw conv outputs bf16, fp32 accumulate, fp32 classifier tail) and of candidate changes, reporting the dataset
EPE delta vs the fp32 oracle.   SEEDS=1,2,.. python tools/emulate_bf16_faithful.py"""
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


def r2(x):
    hi = r(x)
    return hi + r(x - hi)


def h(x):  # fp16 storage
    return x.to(torch.float16).to(torch.float32)


IDENT = lambda t: t  # noqa: E731


class Cfg:
    def __init__(self, name, act=r, wt=r, raw=r, store=r):
        self.name, self.act, self.wt, self.raw, self.store = name, act, wt, raw, store


def run(sd, cost, cfg):
    """Mirror of cmfsm._aggregate_bf16: `store` = rounding of every stored activation (also the residual streams),
    `raw` = rounding of the stored raw conv output, `act`/`wt` = operand rounding inside the conv."""
    def cg(key, x, stride=1, res=None, relu=False, transposed=False):
        w = cfg.wt(sd[key + ".0.weight"])
        if transposed:
            y = F.conv_transpose3d(cfg.act(x), w, None, stride=2, padding=1, output_padding=1)
        else:
            y = F.conv3d(cfg.act(x), w, None, stride, 1)
        y = orc._gn(sd, key + ".1", cfg.raw(y))
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return cfg.store(y)

    def hourglass(key, x, presqu, postsqu, resid):
        out = cg(key + ".conv1.0", x, 2, relu=True)
        pre = cg(key + ".conv2", out, res=postsqu, relu=True)
        out = cg(key + ".conv3.0", pre, 2, relu=True)
        out = cg(key + ".conv4.0", out, relu=True)
        post = cg(key + ".conv5", out, res=presqu if presqu is not None else pre, relu=True, transposed=True)
        out = cg(key + ".conv6", post, res=resid, transposed=True)
        return out, pre, post

    def classif(key, x):
        y = F.conv3d(cfg.act(x), cfg.wt(sd[key + ".0.0.weight"]), None, 1, 1)
        t = F.relu(orc._gn(sd, key + ".0.1", cfg.raw(y)))  # kept fp32
        return F.conv3d(t, sd[key + ".2.weight"], None, 1, 1).squeeze(1)

    cost = cfg.store(cost)
    c0 = cg("dres0.0", cost, relu=True)
    c0 = cg("dres0.2", c0, relu=True)
    t = cg("dres1.0", c0, relu=True)
    cost0 = cg("dres1.2", t, res=c0)
    out1, pre1, post1 = hourglass("dres2", cost0, None, None, cost0)
    out2, _p2, post2 = hourglass("dres3", out1, pre1, post1, cost0)
    out3, _p3, _q3 = hourglass("dres4", out2, pre1, post2, cost0)
    return classif("classif1", out1), classif("classif2", out2), classif("classif3", out3)


def main():
    torch.set_num_threads(os.cpu_count())
    from cmf.models.cmfsm import cmfsm

    torch.manual_seed(gc.WEIGHT_SEED)
    sd = {k: v.detach() for k, v in cmfsm().state_dict().items()}
    seeds = [int(v) for v in os.environ.get("SEEDS", "1,2,3,4").split(",")]
    cfgs = [Cfg("A faithful bf16 (round 2 GPU path)"),
            Cfg("B raw conv out fp32", raw=IDENT),
            Cfg("C weights fp32 (activation rounding only)", wt=IDENT),
            Cfg("D activations 2-term, weights bf16", act=r2, raw=IDENT, store=r2),
            Cfg("E fp16 storage + fp16 operands", act=h, wt=h, raw=h, store=h),
            Cfg("F weights 2-term, activations bf16", wt=r2),
            Cfg("G both 2-term", act=r2, wt=r2, raw=IDENT, store=r2)]
    tot = {c.name: [0.0] * 3 for c in cfgs}
    dev = {c.name: [0.0] * 3 for c in cfgs}
    for seed in seeds:
        left, right = gc.structured_pair(256, 512, delta=20, seed=seed)
        with torch.no_grad():
            L, all_l = orc.feature_extraction(sd, left)
            R, _ = orc.feature_extraction(sd, right)
            weights = orc.context_mapping_weights(sd, L, all_l)
            cost = orc.cost_volume_concat(L, R, 48)
            ref = orc.softargmin_ctxmap(*orc.aggregation3d(sd, cost), weights, 4)
            for c in cfgs:
                got = orc.softargmin_ctxmap(*run(sd, cost, c), weights, 4)
                for i, (a, b) in enumerate(zip(got, ref)):
                    tot[c.name][i] += (float((a - 20).abs().mean()) - float((b - 20).abs().mean())) / len(seeds)
                    dev[c.name][i] += float((a - b).abs().mean()) / len(seeds)
        print("seed", seed, "done", flush=True)
    for c in cfgs:
        print("%-44s dEPE %+.4f %+.4f %+.4f   mean|d| %.3f %.3f %.3f" % (c.name, *tot[c.name], *dev[c.name]), flush=True)


if __name__ == "__main__":
    main()
