"""GPU probe: how does the tcgen05 fp32 accumulator round?  (decides the design of the fp32-accurate split-bf16 mode)

A 32->1 3x3x3 conv with bf16-exact operands through (a) the tcgen05 kernel (54 MMAs of K=16 in three chains, fp32
TMEM accumulate, two fp32 adds), (b) the fp32 FFMA kernel (864 round-to-nearest FMAs), both against an fp64 conv of
the same operands.  Reports the rms error and the bias toward zero, each relative to rms(|y|).
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf_b200 import ops  # noqa: E402

dev = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator().manual_seed(0)
for label, positive in (("normal operands", False), ("positive operands", True)):
    x = torch.randn(1, 32, 8, 32, 64, generator=g)
    w = torch.randn(1, 32, 3, 3, 3, generator=g) * 0.05
    if positive:
        x, w = x.abs(), w.abs()
    x = x.to(torch.bfloat16).float().to(dev)
    w = w.to(torch.bfloat16).float().to(dev)
    ref = F.conv3d(x.double(), w.double(), None, 1, 1)[:, 0]
    padded = torch.zeros(32, 32, 3, 3, 3, device=dev)
    padded[:1] = w
    y_tc = ops.conv3d_igemm_cout1(ops.f32_to_c8(x), ops.pack_igemm_weight(padded))
    y_ff = ops.conv3d_k3(x, ops.pack_conv3d_weight(w), 1)[0][:, 0]
    y_cd = F.conv3d(x, w, None, 1, 1)[:, 0]
    scale = float(ref.pow(2).mean().sqrt())
    for name, y in (("tcgen05 kind::f16 fp32-acc", y_tc), ("FFMA fp32 (ours)", y_ff), ("cuDNN fp32 (TF32 off)", y_cd)):
        e = (y.double() - ref) / scale
        bias = float((e * torch.sign(ref)).mean())
        print("%-18s %-28s rms %.3e  bias-toward-zero %+.3e  max %.3e  (units of rms|y| = %.3g; 2^-24 = 5.96e-8)"
              % (label, name, float(e.pow(2).mean().sqrt()), bias, float(e.abs().max()), scale))
