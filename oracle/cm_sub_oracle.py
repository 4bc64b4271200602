"""CPU oracle of the single-hourglass ablations `cm_sub_4` / `cm_sub_8` / `cm_sub_16` -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/cmf/models/cm_sub_8.py and cm_sub_16.py: feature extractor + six_related_context_mapping of
the corresponding cmfsm_sub_* variant, dres0 + dres1 + ONE hourglass + classif1, then the cost-volume mapping of
cmfsm_sub_16 on that single volume; the prediction is returned three times.  Pinned against the real reference
modules by oracle/gen_golden_cm_sub.py.
"""
import torch
import torch.nn.functional as F

import cmfsm_oracle as base
import cmfsm_sub8_oracle as sub8
import cmfsm_sub16_oracle as sub16


def aggregation_single(sd, cost):
    """dres0, dres1 (+skip), dres2 (+skip), classif1 -- cm_sub_8.py / cm_sub_16.py forward."""
    cost0 = F.relu(base._convgn3d(sd, "dres0.0", cost))
    cost0 = F.relu(base._convgn3d(sd, "dres0.2", cost0))
    t = F.relu(base._convgn3d(sd, "dres1.0", cost0))
    cost0 = base._convgn3d(sd, "dres1.2", t) + cost0
    out1, _pre, _post = base.hourglass(sd, "dres2", cost0, None, None)
    return base.classif(sd, "classif1", out1 + cost0)


def forward(sd, left, right, variant, maxdisp=192, stages=None):
    """variant: "4", "8" or "16".  Returns (pred1, pred1, pred1), each [B,H,W].  cm_sub_4 keeps cmfsm's 1/4-resolution
    extractor (with dilations 2 / 4 in layer3 / layer4, cm_sub_4.py:148-149) and the pre-GroupNorm stem output as hr."""
    if variant == "4":
        def fe(sd_, x):
            return base.feature_extraction(sd_, x, dilations=(2, 4))
    else:
        fe = sub8.feature_extraction if variant == "8" else sub16.feature_extraction
    with torch.no_grad():
        sd = base.strip_module_prefix(sd)
        L, all_l = fe(sd, left)
        R, all_r = fe(sd, right)
        scale = all_l.shape[-1] // L.shape[-1]
        w5, w3 = sub16.context_mapping_weights(sd, L, all_l, R, all_r)
        c1 = aggregation_single(sd, base.cost_volume_concat(L, R, maxdisp // scale))
        fused = sub16._disparity_mix(sub16._spatial_mix(sub16._up3(c1, scale), w5, scale), sub16._target_volumes(w3, maxdisp), scale)
        pred1 = base.softargmin(fused)
        if stages is not None:
            stages.update(L=L, w5=w5, w3=w3, c1=c1)
        return pred1, pred1, pred1
