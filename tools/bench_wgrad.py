"""Dev probe (GPU): cmfb200_conv_wgrad per layer shape of the training step (config 3: 256x512 crops, batch 8; the 2-D
feature layers see 16 images), CUDA-event time and fp32 MAC rate against the FFMA peak (148 SMs x 128 lanes x clock)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
# (name, count in the network, B, Cin, Cout, D, H, W, k, stride, dilation)
LAYERS = [
    ("3d 64->32 1/4", 1, 8, 64, 32, 48, 64, 128, 3, 1, 1),
    ("3d 32->32 1/4", 6, 8, 32, 32, 48, 64, 128, 3, 1, 1),
    ("3d 64->64 1/8", 3, 8, 64, 64, 24, 32, 64, 3, 1, 1),
    ("3d 64->64 1/16", 3, 8, 64, 64, 12, 16, 32, 3, 1, 1),
    ("3d 32->64 s2", 3, 8, 32, 64, 48, 64, 128, 3, 2, 1),
    ("3d 64->64 s2", 3, 8, 64, 64, 24, 32, 64, 3, 2, 1),
    ("2d 32->32 1/2", 8, 16, 32, 32, 1, 128, 256, 3, 1, 1),
    ("2d 64->64 1/4", 31, 16, 64, 64, 1, 64, 128, 3, 1, 1),
    ("2d 64->128 1/4", 1, 16, 64, 128, 1, 64, 128, 3, 1, 1),
    ("2d 128->128 1/4", 5, 16, 128, 128, 1, 64, 128, 3, 1, 1),
    ("2d 128->128 d2", 6, 16, 128, 128, 1, 64, 128, 3, 1, 2),
    ("2d 320->128 1/4", 1, 16, 320, 128, 1, 64, 128, 3, 1, 1),
    ("2d 32->64 s2", 1, 16, 32, 64, 1, 128, 256, 3, 2, 1),
    ("2d 3->32 s2", 1, 16, 3, 32, 1, 256, 512, 3, 2, 1),
    ("2d 128->32 1x1", 1, 16, 128, 32, 1, 64, 128, 1, 1, 1),
]
total = 0.0
for name, cnt, B, Cin, Cout, D, H, W, k, s, dil in LAYERS:
    three = D > 1
    x = torch.randn((B, Cin) + ((D, H, W) if three else (H, W)), device=dev)
    osz = tuple((v - 1) // s + 1 for v in x.shape[2:])
    dy = torch.randn((B, Cout) + osz, device=dev)
    for _ in range(2):
        ops.conv_wgrad(x, dy, k, s, dil)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        ops.conv_wgrad(x, dy, k, s, dil)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    macs = B * Cin * Cout * (27 if three else k * k) * float(torch.tensor(osz).prod())
    total += ms * cnt
    print("%-18s x%-2d %8.3f ms  %6.2f TMAC/s  (%.0f%% of 37.2)" % (name, cnt, ms, macs / ms * 1e-9, macs / ms * 1e-9 / 37.2 * 100))
print("sum over the network: %.1f ms" % total)
