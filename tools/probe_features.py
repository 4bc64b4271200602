"""Dev probe (GPU): where does the non-kernel time of a step go?  Times the cuDNN 2-D feature extractor with
cudnn.benchmark off/on (fp32, TF32 off) and the rest of the forward."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf.models import get_model  # noqa: E402

torch.manual_seed(0)
model = get_model("cmfsm").cuda().eval()
x = torch.rand(2, 3, 576, 960, device="cuda")


def timeit(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


with torch.no_grad():
    for bench in (False, True):
        torch.backends.cudnn.benchmark = bench
        with torch.backends.cudnn.flags(enabled=True, benchmark=bench, allow_tf32=False):
            ms = timeit(lambda: model.feature_extraction(x))
        print("feature_extraction [2,3,576,960] fp32 cudnn.benchmark=%s: %.2f ms" % (bench, ms))
    with torch.backends.cudnn.flags(enabled=True, benchmark=True, allow_tf32=True):
        print("  (tf32 allowed, benchmark) %.2f ms" % timeit(lambda: model.feature_extraction(x)))
    fe = model.feature_extraction
    with torch.backends.cudnn.flags(enabled=True, benchmark=True, allow_tf32=False):
        print("  firstconv only: %.2f ms" % timeit(lambda: fe.firstconv(x)))
        full = fe.firstconv(x)
        half_in = fe.secondconv(full)
        print("  secondconv: %.2f ms" % timeit(lambda: fe.secondconv(full)))
        print("  layer1: %.2f ms" % timeit(lambda: fe.layer1(half_in)))
        h = fe.layer1(half_in)
        print("  layer2: %.2f ms" % timeit(lambda: fe.layer2(h)))
        raw = fe.layer2(h)
        print("  layer3+4: %.2f ms" % timeit(lambda: fe.layer4(fe.layer3(raw))))
