"""GPU: the tensor-core (tc3) 2-D extractor vs the FFMA one vs the fp64 oracle (accuracy at 256x512, time at 576x960)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cmfsm_oracle as orc  # noqa: E402
import golden_common as gc  # noqa: E402
from cmf.models import get_model  # noqa: E402
from cmf_b200 import ops  # noqa: E402

dev = "cuda:0"


def rel(a, b):
    return float((a.detach().cpu().double() - b.double()).norm() / b.double().norm())


torch.manual_seed(gc.WEIGHT_SEED)
model = get_model("cmfsm").to(dev).eval()
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
for name, (left, _r) in (("uniform", gc.seeded_pair(1, 256, 512)), ("structured", gc.structured_pair(256, 512))):
    f32, a32 = orc.feature_extraction(sd, left)
    f64, a64 = orc.feature_extraction({k: v.double() for k, v in sd.items()}, left.double())
    with torch.no_grad(), ops.sums_pool():
        ft, fullt = model._features_tc3(left.to(dev))
        ff, fullf = model._features(left.to(dev))
    print("%-10s feat rel-L2 vs fp64: tc3 %.2e  ffma %.2e  ref32 %.2e ; full: tc3 %.2e ffma %.2e ref32 %.2e"
          % (name, rel(ft, f64), rel(ff, f64), rel(f32, f64), rel(fullt, a64), rel(fullf, a64), rel(a32, a64)), flush=True)

both = torch.rand(2, 3, 576, 960, device=dev)
for label, fn in (("tc3", model._features_tc3), ("ffma", model._features)):
    with torch.no_grad():
        for _ in range(3):
            with ops.sums_pool():
                fn(both)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            with ops.sums_pool():
                fn(both)
        e1.record()
        torch.cuda.synchronize()
    print("%s features, 2 x 576x960: %.3f ms" % (label, e0.elapsed_time(e1) / 10), flush=True)
ops.enable_event_timing(True)
with torch.no_grad(), ops.sums_pool():
    model._features_tc3(both)
for k, (n, ms) in sorted(ops.drain_event_timing().items()):
    print("  %-28s %3d launches %.3f ms" % (k, n, ms))

# ---- whole forward: both engines vs the fp64 oracle (256x512 structured pair), then time at 576x960
left, right = gc.structured_pair(256, 512, delta=20)
ref32 = orc.forward(sd, left, right, 192)
ref64 = orc.forward({k: v.double() for k, v in sd.items()}, left.double(), right.double(), 192)
for engine in ("tc3", "ffma"):
    model.conv_engine = engine
    with torch.no_grad():
        got = model(left.to(dev), right.to(dev))
    for i, (a, b32, b64) in enumerate(zip(got, ref32, ref64), 1):
        ours, theirs = (a.cpu().double() - b64).abs(), (b32.double() - b64).abs()
        print("%-4s pred%d |ours-fp64| max %.2e mean %.2e   |ref32-fp64| max %.2e mean %.2e"
              % (engine, i, ours.max(), ours.mean(), theirs.max(), theirs.mean()), flush=True)
l2, r2 = torch.rand(1, 3, 576, 960, device=dev), torch.rand(1, 3, 576, 960, device=dev)
for engine in ("tc3", "ffma"):
    model.conv_engine = engine
    with torch.no_grad():
        for _ in range(3):
            model(l2, r2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model(l2, r2)
        e1.record()
        torch.cuda.synchronize()
    print("%s forward 576x960 (eager): %.3f ms" % (engine, e0.elapsed_time(e1) / 10), flush=True)
model.conv_engine = "tc3"
ops.enable_event_timing(True)
with torch.no_grad():
    model(l2, r2)
for k, (n, ms) in sorted(ops.drain_event_timing().items()):
    print("  %-28s %3d launches %.3f ms" % (k, n, ms))
