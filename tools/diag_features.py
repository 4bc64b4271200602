"""Dev diagnostic (GPU): per-stage error of the 2-D feature extractor, ours vs cuDNN-fp32 vs the fp64 CPU oracle,
with teacher forcing (every stage is fed the fp64 oracle's input rounded to fp32)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cmfsm_oracle as orc  # noqa: E402
import golden_common as gc  # noqa: E402
from cmf.models import get_model  # noqa: E402
from cmf_b200 import ops  # noqa: E402

torch.manual_seed(0)
model = get_model("cmfsm").cuda().eval()
fe = model.feature_extraction
left, _ = gc.seeded_pair(1, 256, 512)
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
st64, st32 = {}, {}
orc.feature_extraction({k: v.double() for k, v in sd.items()}, left.double(), stages=st64)
orc.feature_extraction(sd, left, stages=st32)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def run_layer(layer, x, ours):
    with torch.no_grad():
        if not ours:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return layer(x)
        o = x
        for unit in layer:
            t = model._cg2(unit.conv1[0], o, relu=True)
            skip = o if unit.downsample is None else model._cg2(unit.downsample, o)
            o = model._cg2(unit.conv2, t, residual=skip)
        return o


print("end-to-end (no teacher forcing):")
with torch.no_grad():
    feat, full = model._features(left.cuda())
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        cfeat, _, cfull = fe(left.cuda())
print("  feat: ours %.2e cudnn %.2e cpu32 %.2e | full: ours %.2e cudnn %.2e cpu32 %.2e" % (
    rel(feat, st64["feat"]), rel(cfeat, st64["feat"]), rel(st32["feat"], st64["feat"]),
    rel(full, st64["full"]), rel(cfull, st64["full"]), rel(st32["full"], st64["full"])))

print("teacher-forced per stage (input = fp64 oracle stage rounded to fp32): ours | cudnn | cpu32(not forced)")
prev = "second"
for name in ("layer1", "layer2", "layer3", "layer4"):
    x = st64[prev].float().cuda()
    a = run_layer(getattr(fe, name), x, True)
    b = run_layer(getattr(fe, name), x, False)
    print("  %-7s ours %.2e  cudnn %.2e  cpu32 %.2e" % (name, rel(a, st64[name]), rel(b, st64[name]), rel(st32[name], st64[name])))
    prev = name
# single units of layer3 / layer4 with forcing, conv by conv
x = st64["layer2"].float().cuda()
unit = fe.layer3[0]
with torch.no_grad():
    raw_o, sums = model._c2(unit.conv1[0][0], x, True)
    ref_raw = F.conv2d(st64["layer2"], sd["feature_extraction.layer3.0.conv1.0.0.weight"].double(), None, 1, 1)
    print("  layer3.0.conv1 raw conv: ours %.2e" % rel(raw_o, ref_raw))
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        print("  layer3.0.conv1 raw conv: cudnn %.2e" % rel(unit.conv1[0][0](x), ref_raw))
    g = unit.conv1[0][1]
    want = F.group_norm(ref_raw, 32, g.weight.double().cpu(), g.bias.double().cpu(), 1e-5)
    mine = ops.gn_apply(raw_o, sums, g.weight, g.bias)
    print("  layer3.0.conv1 +GN: ours %.2e ; torch GN on our raw %.2e" % (rel(mine, want), rel(g(raw_o), want)))
    mean = ref_raw.mean((2, 3)); std = ref_raw.std((2, 3))
    print("  |mean|/std per channel: max %.2f median %.2f" % (float((mean.abs() / std).max()), float((mean.abs() / std).median())))
# SPP branch + lastconv
x = st64["layer4"].float().cuda()
with torch.no_grad():
    for i, key in ((1, "b1"), (2, "b2"), (3, "b3"), (4, "b4")):
        branch = getattr(fe, "branch%d" % i)
        pooled = branch[0](x)
        a = F.interpolate(model._cg2(branch[1], pooled.contiguous(), relu=True), x.shape[2:], mode="bilinear", align_corners=False)
        b = F.interpolate(branch(x), x.shape[2:], mode="bilinear", align_corners=False)
        print("  %s ours %.2e cudnn %.2e cpu32 %.2e" % (key, rel(a, st64[key]), rel(b, st64[key]), rel(st32[key], st64[key])))
    cat = torch.cat([st64[k] for k in ("layer2", "layer4", "b4", "b3", "b2", "b1")], 1).float().cuda()
    a = model._cg2(fe.lastconv[0], cat, relu=True)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        b = F.relu(fe.lastconv[0](cat))
    print("  last0 ours %.2e cudnn %.2e cpu32 %.2e" % (rel(a, st64["last0"]), rel(b, st64["last0"]), rel(st32["last0"], st64["last0"])))
