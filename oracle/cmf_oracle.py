"""CPU oracle of the registry model `cmf` -- TEST INFRASTRUCTURE ONLY (same rules as cmfsm_oracle.py).

Restates /root/reference/cmf/models/cmf.py (cited `cmf.py`): PSMNet-style extractor with a stride-2 stem (features at 1/4,
`layer1` output at 1/2 kept for the refinement), the cmfsm 3-D aggregation with cumulative classifier volumes, soft-argmin
at 1/4 resolution over maxdisp/4 planes, and `super_resolution_refinement`: conv(1->64) on the low-resolution disparity,
two transposed-conv stages fed with cat([x, feature]) (1/4 features, then the 1/2 `layer1` output), cat with three
conv layers of the RGB image, conv(96->96), conv(96->1, bias), ReLU.  Pinned by oracle/gen_golden_cmf.py.
"""
import torch
import torch.nn.functional as F

import cmfsm_oracle as base


def feature_extraction(sd, x, prefix="feature_extraction"):
    """cmf.py feature_extraction.forward: returns (feature [B,32,H/4,W/4], layer1 output [B,32,H/2,W/2])."""
    p = prefix
    o = F.relu(base._convgn2d(sd, p + ".firstconv.0", x, stride=2))
    o = F.relu(base._convgn2d(sd, p + ".firstconv.2", o))
    o = F.relu(base._convgn2d(sd, p + ".firstconv.4", o))
    half = base._layer(sd, p + ".layer1", o, 3, 1, 1)
    raw = base._layer(sd, p + ".layer2", half, 16, 2, 1)
    l3 = base._layer(sd, p + ".layer3", raw, 3, 1, 1)
    skip = base._layer(sd, p + ".layer4", l3, 3, 1, 2)
    size = skip.shape[2:]
    branches = []
    for name, k in (("branch1", 64), ("branch2", 32), ("branch3", 16), ("branch4", 8)):
        b = F.avg_pool2d(skip, (k, k), (k, k))
        b = F.relu(base._convgn2d(sd, "%s.%s.1" % (p, name), b, 1, 0, 1))
        branches.append(F.interpolate(b, size, mode="bilinear", align_corners=False))
    b1, b2, b3, b4 = branches
    o = F.relu(base._convgn2d(sd, p + ".lastconv.0", torch.cat((raw, skip, b4, b3, b2, b1), 1)))
    return F.conv2d(o, sd[p + ".lastconv.2.weight"]), half


def srr(sd, pred_lr, rgb, feat, half, prefix="srr"):
    """super_resolution_refinement.forward (cmf.py): pred_lr [B,h,w] -> [B,1,H,W]."""
    p = prefix
    x = F.relu(base._convgn2d(sd, p + ".conv1.0", pred_lr.unsqueeze(1)))
    for i, z in enumerate((feat, half)):
        k = "%s.deconv_module_list.%d" % (p, i)
        x = F.conv_transpose2d(torch.cat([x, z], 1), sd[k + ".0.weight"], sd[k + ".0.bias"], stride=2, padding=1,
                               output_padding=1)
        x = F.relu(base._gn(sd, k + ".1", x))
    r = rgb
    for i in (0, 2, 4):
        r = F.relu(base._convgn2d(sd, "%s.rgb_fea.%d" % (p, i), r))
    x = F.relu(base._convgn2d(sd, p + ".conv2.0", torch.cat([x, r], 1)))
    return F.relu(F.conv2d(x, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"], 1, 1))


def forward(sd, left, right, maxdisp=192, stages=None):
    """cmf.forward: three [B,1,H,W] refined disparity maps."""
    with torch.no_grad():
        sd = base.strip_module_prefix(sd)
        L, half = feature_extraction(sd, left)
        R, _ = feature_extraction(sd, right)
        c1, c2, c3 = base.aggregation3d(sd, base.cost_volume_concat(L, R, maxdisp // 4))
        outs, cost, lows = [], None, []
        for c in (c1, c2, c3):
            cost = c if cost is None else c + cost
            lows.append(base.softargmin(cost))
            outs.append(srr(sd, lows[-1], left, L, half))
        if stages is not None:
            stages.update(L=L, half=half, c3=c3, low3=lows[-1])
        return tuple(outs)
