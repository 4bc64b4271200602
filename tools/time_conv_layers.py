"""Dev probe (GPU): time every distinct conv2d / conv3d layer shape of BASELINE config 2 in isolation
(CUDA events, 10 repetitions) and print achieved TMAC/s per shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ONLY_BF16 = "--bf16" in sys.argv
print("conv2d (B=2 images)")
H, W = 576, 960
cases2d = [  # name, count, Cin, Cout, H, W, k, s, d
    ("stem 3->32", 1, 3, 32, H, W, 3, 1, 1), ("stem 32->32", 3, 32, 32, H, W, 3, 1, 1),
    ("second s2", 1, 32, 32, H, W, 3, 2, 1), ("half 32->32", 7, 32, 32, H // 2, W // 2, 3, 1, 1),
    ("layer2.0 s2 32->64", 1, 32, 64, H // 2, W // 2, 3, 2, 1), ("layer2 ds 1x1 s2", 1, 32, 64, H // 2, W // 2, 1, 2, 1),
    ("layer2 64->64", 31, 64, 64, H // 4, W // 4, 3, 1, 1), ("layer3.0 64->128", 1, 64, 128, H // 4, W // 4, 3, 1, 1),
    ("layer3 ds 1x1", 1, 64, 128, H // 4, W // 4, 1, 1, 1), ("layer3 128->128", 5, 128, 128, H // 4, W // 4, 3, 1, 1),
    ("layer4 128->128 d2", 6, 128, 128, H // 4, W // 4, 3, 1, 2), ("lastconv 320->128", 1, 320, 128, H // 4, W // 4, 3, 1, 1),
    ("lastconv 1x1 128->32", 1, 128, 32, H // 4, W // 4, 1, 1, 1),
]
tot = 0.0
for name, cnt, ci, co, h, w, k, s, d in ([] if ONLY_BF16 else cases2d):
    x = torch.randn(2, ci, h, w, device="cuda")
    wp = ops.pack_conv2d_weight(torch.randn(co, ci, k, k, device="cuda"))
    ms = timeit(lambda: ops.conv2d(x, wp, k, s, d, want_stats=True))
    ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
    gmac = 2 * ci * co * k * k * ho * wo / 1e9
    tot += ms * cnt
    print("  %-22s x%2d  %7.3f ms  %6.2f GMAC  %5.1f TMAC/s  (total %.2f ms)" % (name, cnt, ms, gmac, gmac / ms, ms * cnt))
print("  conv2d total %.2f ms" % tot)

print("conv3d fp32 (B=1)")
D, h, w = 48, 144, 240
cases3d = [("dres0.0 64->32", 1, 64, 32, D, h, w, 1, False), ("32->32 full", 6, 32, 32, D, h, w, 1, False),
           ("conv1 s2 32->64", 3, 32, 64, D, h, w, 2, False), ("conv2 64->64 1/8", 3, 64, 64, D // 2, h // 2, w // 2, 1, False),
           ("conv3 s2 64->64", 3, 64, 64, D // 2, h // 2, w // 2, 2, False), ("conv4 64->64 1/16", 3, 64, 64, D // 4, h // 4, w // 4, 1, False),
           ("conv5 deconv 64->64", 3, 64, 64, D // 4, h // 4, w // 4, 1, True), ("conv6 deconv 64->32", 3, 64, 32, D // 2, h // 2, w // 2, 1, True),
           ("classif 32->1", 3, 32, 1, D, h, w, 1, False)]
tot = 0.0
for name, cnt, ci, co, d_, h_, w_, s, tr in ([] if ONLY_BF16 else cases3d):
    x = torch.randn(1, ci, d_, h_, w_, device="cuda")
    wt = torch.randn(ci, co, 3, 3, 3, device="cuda") if tr else torch.randn(co, ci, 3, 3, 3, device="cuda")
    wp = ops.pack_conv3d_weight(wt, transposed=tr)
    ms = timeit(lambda: ops.conv3d_k3(x, wp, s, tr, want_stats=co > 1))
    vox_out = (8 * d_ * h_ * w_) if tr else (((d_ - 1) // s + 1) * ((h_ - 1) // s + 1) * ((w_ - 1) // s + 1))
    gmac = (d_ * h_ * w_ if tr else vox_out) * ci * co * 27 / 1e9
    tot += ms * cnt
    print("  %-22s x%2d  %7.3f ms  %6.2f GMAC  %5.1f TMAC/s  (total %.2f ms)" % (name, cnt, ms, gmac, gmac / ms, ms * cnt))
print("  conv3d total %.2f ms" % tot)

print("conv3d bf16 tcgen05 (B=1, C8 layout)")
BF = torch.bfloat16
cases_ig = [("dres0.0 64->32", 1, "s1", 64, 32, D, h, w), ("32->32 full", 8, "s1", 32, 32, D, h, w),
            ("conv1 s2 32->64", 3, "s2", 32, 64, D, h, w), ("conv2 64->64 1/8", 3, "s1", 64, 64, D // 2, h // 2, w // 2),
            ("conv3 s2 64->64", 3, "s2", 64, 64, D // 2, h // 2, w // 2), ("conv4 64->64 1/16", 3, "s1", 64, 64, D // 4, h // 4, w // 4),
            ("conv5 deconv 64->64", 3, "tr", 64, 64, D // 4, h // 4, w // 4), ("conv6 deconv 64->32", 3, "tr", 64, 32, D // 2, h // 2, w // 2)]
tot = totf = 0.0
for name, cnt, kind, ci, co, d_, h_, w_ in cases_ig:
    wt = torch.randn(ci, co, 3, 3, 3, device="cuda") if kind == "tr" else torch.randn(co, ci, 3, 3, 3, device="cuda")
    wp = ops.pack_igemm_weight(wt, transposed=(kind == "tr"))
    if kind == "s2":
        x = torch.randn(1, 8, ci // 8, d_ // 2, h_ // 2, w_ // 2, 8, device="cuda").to(BF)
        ms = timeit(lambda: ops.conv3d_s2_igemm(x, wp))
        gmac = (d_ // 2) * (h_ // 2) * (w_ // 2) * ci * co * 27 / 1e9
    elif kind == "tr":
        x = torch.randn(1, ci // 8, d_, h_, w_, 8, device="cuda").to(BF)
        ms = timeit(lambda: ops.deconv3d_igemm(x, wp))
        gmac = d_ * h_ * w_ * ci * co * 27 / 1e9
    else:
        x = torch.randn(1, ci // 8, d_, h_, w_, 8, device="cuda").to(BF)
        ms = timeit(lambda: ops.conv3d_igemm(x, wp))
        gmac = d_ * h_ * w_ * ci * co * 27 / 1e9
    tot += ms * cnt
    totf += 2 * gmac * cnt
    print("  %-22s x%2d  %7.3f ms  %6.2f GMAC  %6.1f TFLOP/s  (total %.3f ms)" % (name, cnt, ms, gmac, 2 * gmac / ms, ms * cnt))
print("  tcgen05 total %.3f ms, %.1f TFLOP/s" % (tot, totf / tot))
