"""CPU oracle of the 1/16-resolution variant `cmfsm_sub_16` -- TEST INFRASTRUCTURE ONLY (same rules as cmfsm_oracle.py).

Functional restatement of /root/reference/cmf/models/cmfsm_sub_16.py (cited `sub16.py:line`).  It shares the 3-D
aggregation with cmfsm and the five reference-image mapping weights with cmfsm_sub_8; what is specific:
  * feature extractor: like sub_8 but layer3 has stride 2 (features at 1/16, D' = maxdisp/16) and dilation 1, SPP
    pools 4/2/16/8, and `lastconv_16` takes 384 channels = cat(layer3 out, layer4 out, 4 branches) (sub16.py:127-239);
  * six_related_context_mapping also returns three TARGET-image weights (centre, right, left of the right image,
    softmax*logit over the three) and they ARE used here (sub16.py:451-573);
  * the three classifier volumes are accumulated AFTER nearest upsampling to [B, maxdisp, H, W] (cost2 += cost1 etc.,
    sub16.py:768-773, 809-814, 829-834), every volume is mixed over the five spatial neighbours, then over the
    disparity axis with the target weights shifted by the disparity (sub16.py:774-805), and regressed with a softmax
    over all `maxdisp` planes.  Outputs are [B,H,W] (no channel dimension).
Pinned against the real reference module: oracle/gen_golden_sub16.py.
"""
import torch
import torch.nn.functional as F

import cmfsm_oracle as base
import cmfsm_sub8_oracle as sub8

TARGET3_DYDX = ((0, 0), (0, 1), (0, -1))  # centre, right, left (sub16.py:562)


def feature_extraction(sd, x, prefix="feature_extraction", stages=None):
    """sub16.py:199-239.  Returns (feature [B,32,H/16,W/16], all_feature [B,32,H,W])."""
    p = prefix
    o = x
    for i in (0, 2, 4, 6):
        o = F.relu(base._convgn2d(sd, "%s.firstconv.%d" % (p, i), o))
    all_feature = o
    o = F.relu(base._convgn2d(sd, p + ".secondconv.0", o, stride=2))
    o = F.relu(base._convgn2d(sd, p + ".secondconv.2", o))
    l1 = base._layer(sd, p + ".layer1", o, 3, 2, 1)
    l2 = base._layer(sd, p + ".layer2", l1, 16, 2, 1)
    raw = base._layer(sd, p + ".layer3", l2, 3, 2, 1)   # sub16.py:207 rebinds output_raw to the layer3 output
    skip = base._layer(sd, p + ".layer4", raw, 3, 1, 4)
    size = skip.shape[2:]
    branches = []
    for name, k in (("branch1", 4), ("branch2", 2), ("branch3", 16), ("branch4", 8)):
        b = F.avg_pool2d(skip, (k, k), (k, k))
        b = F.relu(base._convgn2d(sd, "%s.%s.1" % (p, name), b, 1, 0, 1))
        branches.append(F.interpolate(b, size, mode="bilinear", align_corners=False))
    b1, b2, b3, b4 = branches
    cat = torch.cat((raw, skip, b4, b3, b2, b1), 1)  # 128 + 128 + 4*32 = 384
    o = F.relu(base._convgn2d(sd, p + ".lastconv_16.0", cat))
    feat = F.conv2d(o, sd[p + ".lastconv_16.2.weight"])
    return feat, all_feature


def _logits(sd, lr, hr, neighbours, codes):
    B, C, h, w = lr.shape
    H, W = hr.shape[2:]
    s = W // w
    lr_up = lr.repeat_interleave(s, 2).repeat_interleave(s, 3)
    out = []
    for k, (dy, dx) in enumerate(neighbours):
        y0, y1 = max(0, -dy) * s, H - max(0, dy) * s
        x0, x1 = max(0, -dx) * s, W - max(0, dx) * s
        code = codes[k].repeat(1, H // s, W // s).unsqueeze(0).expand(B, -1, -1, -1)
        rep = torch.cat([lr_up[:, :, y0 + dy * s:y1 + dy * s, x0 + dx * s:x1 + dx * s],
                         hr[:, :, y0:y1, x0:x1], code[:, :, y0:y1, x0:x1]], 1)
        lg = hr.new_zeros((B, 1, H, W))
        lg[:, :, y0:y1, x0:x1] = sub8.similarity_mlp(sd, rep)
        out.append(lg)
    return torch.cat(out, 1)


def context_mapping_weights(sd, lr, hr, lr_r, hr_r):
    """six_related_context_mapping.forward (sub16.py:451-573): ([B,5,H,W], [B,3,H,W]), both softmax(l) * l."""
    s = hr.shape[-1] // lr.shape[-1]
    if s % 2 != 0:
        raise ValueError("odd scale (reference calls exit(), sub16.py:464)")
    codes = sub8.position_code5(s, lr.dtype)
    l5 = _logits(sd, lr, hr, sub8.NEIGHBOUR5_DYDX, codes)
    l3 = _logits(sd, lr_r, hr_r, TARGET3_DYDX, codes[:3])
    return F.softmax(l5, 1) * l5, F.softmax(l3, 1) * l3


def _up3(c, s):
    """nearest upsampling of [B,D,h,w] by s along all three axes (sub16.py:768-773)."""
    return c.repeat_interleave(s, 1).repeat_interleave(s, 2).repeat_interleave(s, 3)


def _spatial_mix(cost, w5, s):
    """sub16.py:774-778: cost*w_c + the right/left/top/bottom neighbours' (cell-shifted) costs times their weights."""
    H, W = cost.shape[2:]
    out = cost * w5[:, 0:1]
    for k in range(1, 5):
        dy, dx = sub8.NEIGHBOUR5_DYDX[k]
        y0, y1 = max(0, -dy) * s, H - max(0, dy) * s
        x0, x1 = max(0, -dx) * s, W - max(0, dx) * s
        add = torch.zeros_like(out)
        add[:, :, y0:y1, x0:x1] = cost[:, :, y0 + dy * s:y1 + dy * s, x0 + dx * s:x1 + dx * s] * w5[:, k:k + 1, y0:y1, x0:x1]
        out = out + add
    return out


def _target_volumes(w3, maxdisp):
    """sub16.py:782-794: volume[b,d,y,x] = w[b,y,x-d] for x >= d, 1 elsewhere, for the three target weights."""
    B, _, H, W = w3.shape
    vols = [w3.new_ones((B, maxdisp, H, W)) for _ in range(3)]
    for d in range(min(maxdisp, W)):
        for v, k in zip(vols, range(3)):
            v[:, d, :, d:] = w3[:, k, :, :W - d]
    return vols


def _disparity_mix(fused, vols, s):
    """sub16.py:796-798."""
    vt, vr, vl = vols
    out = fused * vt
    add = torch.zeros_like(out)
    add[:, :-s] = fused[:, s:] * vl[:, :-s]
    out = out + add
    add = torch.zeros_like(out)
    add[:, s:] = fused[:, :-s] * vr[:, s:]
    return out + add


def volume_mapping(c1, c2, c3, w5, w3, scale, maxdisp):
    """sub16.py:760-850: three [B,H,W] disparity maps from the raw classifier volumes [B,D',h,w]."""
    vols = _target_volumes(w3, maxdisp)
    outs, cost = [], None
    for c in (c1, c2, c3):
        up = _up3(c, scale)
        cost = up if cost is None else up + cost
        fused_t = _disparity_mix(_spatial_mix(cost, w5, scale), vols, scale)
        outs.append(base.softargmin(fused_t))
    return tuple(outs)


def check_shapes(H, W, maxdisp, B=1):
    if H % 64 or W % 64:
        raise ValueError("H and W must be multiples of 64 (1/16 features, two more stride-2 levels in the hourglass)")
    if maxdisp % 64:
        raise ValueError("maxdisp must be a multiple of 64")
    if H < 256 or W < 256 or B * (H // 256) * (W // 256) < 2:
        raise ValueError("the 16x16 SPP pool on the 1/16 map + GroupNorm needs B*(H//256)*(W//256) >= 2")


def forward(sd, left, right, maxdisp=192, stages=None):
    """cmfsm_sub_16.forward -- sub16.py:722-852; returns three [B,H,W] maps."""
    with torch.no_grad():
        sd = base.strip_module_prefix(sd)
        L, all_l = feature_extraction(sd, left)
        R, all_r = feature_extraction(sd, right)
        scale = all_l.shape[-1] // L.shape[-1]
        w5, w3 = context_mapping_weights(sd, L, all_l, R, all_r)
        cost = base.cost_volume_concat(L, R, maxdisp // scale)
        c1, c2, c3 = base.aggregation3d(sd, cost)
        outs = volume_mapping(c1, c2, c3, w5, w3, scale, maxdisp)
        if stages is not None:
            stages.update(L=L, R=R, all_l=all_l, w5=w5, w3=w3, cost=cost, c1=c1, c2=c2, c3=c3)
        return outs
