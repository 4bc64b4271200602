"""CPU oracle of the 1/8-resolution variant `cmfsm_sub_8` -- TEST INFRASTRUCTURE ONLY (same rules as cmfsm_oracle.py:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it; the product path never does).

Functional restatement of /root/reference/cmf/models/cmfsm_sub_8.py (cited as `sub8.py:line`), reusing the blocks that
are identical to `cmfsm` (convbn, BasicBlock, hourglass stack, classifier, cost volume, soft-argmin) from
cmfsm_oracle.py.  What differs from cmfsm:
  * feature extractor: four convbn+ReLU in `firstconv` (its post-ReLU output is the full-resolution "hr" feature),
    `secondconv` = stride-2 convbn + convbn, layer1 AND layer2 stride 2 (features at 1/8), layer3/layer4 dilation 2/4,
    SPP pools 4/32/16/8 (sub8.py:126-236);
  * mapping = six_related_context_mapping (sub8.py:440-572): five reference-image neighbours (centre, right, left,
    top, bottom) with ZERO logits where the neighbour cell is outside the image, a trailing LeakyReLU on the MLP output
    (sub8.py:318,342), weights = softmax(logits) * logits; the three target-image weights it also computes are never
    used by the forward (sub8.py:768-802) and are not restated;
  * the three classifier volumes are NOT accumulated (sub8.py:757,774,790) and each prediction is mapped with the five
    weights in the order c, r, l, t, b (sub8.py:768-772).
Pinned against the real reference module run in the build container: oracle/gen_golden_sub8.py.
"""
import torch
import torch.nn.functional as F

import cmfsm_oracle as base

# neighbour order of the five returned maps (sub8.py:563): centre, right, left, top, bottom
NEIGHBOUR5_DYDX = ((0, 0), (0, 1), (0, -1), (-1, 0), (1, 0))
# (channel-0 kind over x, channel-1 kind over y): sub8.py:497 (right: dec), :516 (left: inc), :534 (top: dec), :546 (bottom: inc)
NEIGHBOUR5_CODES = (("off", "off"), ("dec", "off"), ("inc", "off"), ("off", "dec"), ("off", "inc"))


def feature_extraction(sd, x, prefix="feature_extraction", stages=None):
    """sub8.py:199-236.  Returns (feature [B,32,H/8,W/8], all_feature [B,32,H,W] = firstconv output, post-ReLU)."""
    p = prefix
    o = x
    for i in (0, 2, 4, 6):
        o = F.relu(base._convgn2d(sd, "%s.firstconv.%d" % (p, i), o))
    all_feature = o
    o = F.relu(base._convgn2d(sd, p + ".secondconv.0", o, stride=2))
    o = F.relu(base._convgn2d(sd, p + ".secondconv.2", o))
    l1 = base._layer(sd, p + ".layer1", o, 3, 2, 1)
    raw = base._layer(sd, p + ".layer2", l1, 16, 2, 1)
    l3 = base._layer(sd, p + ".layer3", raw, 3, 1, 2)
    skip = base._layer(sd, p + ".layer4", l3, 3, 1, 4)
    size = skip.shape[2:]
    branches = []
    for name, k in (("branch1", 4), ("branch2", 32), ("branch3", 16), ("branch4", 8)):
        b = F.avg_pool2d(skip, (k, k), (k, k))
        b = F.relu(base._convgn2d(sd, "%s.%s.1" % (p, name), b, 1, 0, 1))
        branches.append(F.interpolate(b, size, mode="bilinear", align_corners=False))
    b1, b2, b3, b4 = branches
    cat = torch.cat((raw, skip, b4, b3, b2, b1), 1)  # sub8.py:229-231
    o = F.relu(base._convgn2d(sd, p + ".lastconv.0", cat))
    feat = F.conv2d(o, sd[p + ".lastconv.2.weight"])
    if stages is not None:
        stages.update(full=all_feature, layer1=l1, layer2=raw, layer3=l3, layer4=skip, feat=feat)
    return feat, all_feature


def position_code5(scale, dtype=torch.float32):
    """[5,2,s,s] positional codes of the five neighbours (sub8.py:452-459, 491-548)."""
    s = scale
    half = s // 2
    off = torch.tensor([float(v) for v in list(range(-half, 0)) + list(range(1, half + 1))], dtype=dtype)
    inc = torch.arange(1, s + 1, dtype=dtype)
    dec = s - inc + 1
    ch0 = {"off": off.view(1, s).expand(s, s), "inc": inc.view(1, s).expand(s, s), "dec": dec.view(1, s).expand(s, s)}
    ch1 = {"off": off.view(s, 1).expand(s, s), "inc": inc.view(s, 1).expand(s, s), "dec": dec.view(s, 1).expand(s, s)}
    return torch.stack([torch.stack((ch0[a], ch1[b])) for a, b in NEIGHBOUR5_CODES])


def similarity_mlp(sd, x, prefix="mapping_matrix.similarity1"):
    """similarity_measure1.forward of the variant: LeakyReLU after EVERY conv, the last one included (sub8.py:334-343)."""
    for i in range(4):
        x = F.leaky_relu(F.conv2d(x, sd["%s.conv%d.weight" % (prefix, i)]), 0.01)
    return x


def context_mapping_weights5(sd, lr, hr):
    """six_related_context_mapping.forward, reference-image half (sub8.py:445-572): [B,5,H,W] = softmax(l) * l with
    l_k = MLP(...) where the neighbour cell exists and 0 elsewhere (the zero `padding1/2` blocks, sub8.py:502,519,537,549)."""
    B, C, h, w = lr.shape
    H, W = hr.shape[2:]
    s = W // w
    if s % 2 != 0:
        raise ValueError("odd scale (reference calls exit(), sub8.py:462)")
    codes = position_code5(s, lr.dtype)
    lr_up = lr.repeat_interleave(s, 2).repeat_interleave(s, 3)
    logits = []
    for k, (dy, dx) in enumerate(NEIGHBOUR5_DYDX):
        y0, y1 = max(0, -dy) * s, H - max(0, dy) * s
        x0, x1 = max(0, -dx) * s, W - max(0, dx) * s
        code = codes[k].repeat(1, H // s, W // s).unsqueeze(0).expand(B, -1, -1, -1)
        rep = torch.cat([lr_up[:, :, y0 + dy * s:y1 + dy * s, x0 + dx * s:x1 + dx * s],
                         hr[:, :, y0:y1, x0:x1], code[:, :, y0:y1, x0:x1]], 1)
        lg = hr.new_zeros((B, 1, H, W))
        lg[:, :, y0:y1, x0:x1] = similarity_mlp(sd, rep)
        logits.append(lg)
    logits = torch.cat(logits, 1)
    return F.softmax(logits, dim=1) * logits  # sub8.py:564-572


def mapped_upsample5(pred, weights5, scale):
    """sub8.py:762-772: refined = pred*w_c, then += the right / left / top / bottom neighbours' predictions."""
    s = scale
    up = s * pred.repeat_interleave(s, 1).repeat_interleave(s, 2)  # [B,H,W]
    H, W = up.shape[1:]
    out = up * weights5[:, 0]
    for k in range(1, 5):
        dy, dx = NEIGHBOUR5_DYDX[k]
        y0, y1 = max(0, -dy) * s, H - max(0, dy) * s
        x0, x1 = max(0, -dx) * s, W - max(0, dx) * s
        add = torch.zeros_like(out)
        add[:, y0:y1, x0:x1] = up[:, y0 + dy * s:y1 + dy * s, x0 + dx * s:x1 + dx * s] * weights5[:, k, y0:y1, x0:x1]
        out = out + add
    return out.unsqueeze(1)


def softargmin_ctxmap5(c1, c2, c3, weights5, scale):
    """Three independent soft-argmin regressions (no cumulative sums) + the five-neighbour mapping."""
    return tuple(mapped_upsample5(base.softargmin(c), weights5, scale) for c in (c1, c2, c3))


def check_shapes(H, W, maxdisp, B=1):
    if H % 32 or W % 32:
        raise ValueError("H and W must be multiples of 32 (1/8 features, two more stride-2 levels in the hourglass)")
    if maxdisp % 32:
        raise ValueError("maxdisp must be a multiple of 32")
    if H // 8 < 32 or W // 8 < 32:
        raise ValueError("the 32x32 SPP pool needs H, W >= 256")
    if B * (H // 256) * (W // 256) < 2:
        raise ValueError("GroupNorm over a single pooled value per channel (branch2) is undefined: need B*(H//256)*(W//256) >= 2")


def forward(sd, left, right, maxdisp=192, stages=None, grad=False):
    """cmfsm_sub_8.forward -- sub8.py:722-804, per-sample [B,1,H,W] outputs."""
    with torch.set_grad_enabled(grad):
        sd = base.strip_module_prefix(sd)
        L, all_l = feature_extraction(sd, left)
        R, _ = feature_extraction(sd, right)
        scale = all_l.shape[-1] // L.shape[-1]
        weights = context_mapping_weights5(sd, L, all_l)
        cost = base.cost_volume_concat(L, R, maxdisp // scale)
        c1, c2, c3 = base.aggregation3d(sd, cost)
        outs = softargmin_ctxmap5(c1, c2, c3, weights, scale)
        if stages is not None:
            stages.update(L=L, R=R, all_l=all_l, weights=weights, cost=cost, c1=c1, c2=c2, c3=c3)
        return outs
