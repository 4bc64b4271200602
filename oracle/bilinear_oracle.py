"""CPU oracle of the `bilinear_cmf`, `bilinear_cmf_sub_8`, `bilinear_cmf_sub_16` baselines -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/cmf/models/bilinear_cmf.py (+ _sub_8 / _sub_16): the feature extractor of cm_sub_4 /
cmfsm_sub_8 / cmfsm_sub_16 (the second feature-extractor output is all that is used), the cmfsm aggregation, cumulative
classifier volumes, `F.interpolate(mode="trilinear", align_corners=False)` to [maxdisp,H,W], softmax + regression
(bilinear_cmf.py:384-452).  No context mapping.  Pinned by oracle/gen_golden_bilinear.py.
"""
import torch
import torch.nn.functional as F

import cmfsm_oracle as base
import cmfsm_sub8_oracle as sub8
import cmfsm_sub16_oracle as sub16


def trilinear_softargmin(c1, c2, c3, maxdisp, H, W):
    outs, cost = [], None
    for c in (c1, c2, c3):
        cost = c if cost is None else c + cost
        up = F.interpolate(cost.unsqueeze(1), [maxdisp, H, W], mode="trilinear", align_corners=False).squeeze(1)
        outs.append(base.softargmin(up))
    return tuple(outs)


def forward(sd, left, right, variant, maxdisp=192, stages=None):
    """variant: "4" (bilinear_cmf), "8", "16".  Returns three [B,H,W] maps."""
    if variant == "4":
        def fe(sd_, x):
            return base.feature_extraction(sd_, x, dilations=(2, 4))
    else:
        fe = sub8.feature_extraction if variant == "8" else sub16.feature_extraction
    with torch.no_grad():
        sd = base.strip_module_prefix(sd)
        L, _ = fe(sd, left)
        R, _ = fe(sd, right)
        scale = left.shape[-1] // L.shape[-1]
        c1, c2, c3 = base.aggregation3d(sd, base.cost_volume_concat(L, R, maxdisp // scale))
        if stages is not None:
            stages.update(L=L, c1=c1, c3=c3)
        return trilinear_softargmin(c1, c2, c3, maxdisp, left.shape[2], left.shape[3])
