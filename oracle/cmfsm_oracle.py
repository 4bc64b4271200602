"""CPU oracle for the `cmfsm` hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict driven) restatement of the reference network's forward pass,
`/root/reference/cmf/models/cmfsm.py:655-775`, in plain PyTorch CPU ops (the reference's own
arithmetic is ATen/oneDNN).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product path
(`explicit-context-mapping-for-stereo-matching_b200/`) never does.

Parity pinning: the reference ships no tests or golden vectors, so this oracle
is pinned against the reference ITSELF, executed in the build container by `oracle/gen_golden.py`
(via `oracle/ref_harness.py`).  That script asserts stage-by-stage equality oracle-vs-reference and
writes `tests/golden/cmfsm_golden.npz`; `tests/test_oracle_golden.py` re-checks the oracle against
those committed fixtures on any box (no reference tree needed).

Every function cites the reference lines it restates.  `sd` is a `state_dict` with the reference's
key names, optionally `module.`-prefixed keys already stripped.
"""
import torch
import torch.nn.functional as F

GN_GROUPS = 32  # cmfsm.py:33  group_norm_group_num
GN_EPS = 1e-5  # torch.nn.GroupNorm default, used everywhere in cmfsm.py


# ----------------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------------
def _gn(sd, key, x):
    """nn.GroupNorm(32, C) -- cmfsm.py:46,58."""
    return F.group_norm(x, GN_GROUPS, sd[key + ".weight"], sd[key + ".bias"], GN_EPS)


def _convgn2d(sd, key, x, stride=1, pad=1, dilation=1):
    """convbn(): Conv2d(bias=False) + GroupNorm -- cmfsm.py:36-46 (padding = dilation if dilation>1)."""
    p = dilation if dilation > 1 else pad
    x = F.conv2d(x, sd[key + ".0.weight"], None, stride, p, dilation)
    return _gn(sd, key + ".1", x)


def _convgn3d(sd, key, x, stride=1):
    """convbn_3d(): Conv3d(k3,p1,bias=False) + GroupNorm -- cmfsm.py:49-58."""
    x = F.conv3d(x, sd[key + ".0.weight"], None, stride, 1)
    return _gn(sd, key + ".1", x)


def _deconvgn3d(sd, key, x):
    """ConvTranspose3d(k3,s2,p1,op1,bias=False) + GroupNorm -- cmfsm.py:261-281."""
    x = F.conv_transpose3d(x, sd[key + ".0.weight"], None, stride=2, padding=1, output_padding=1)
    return _gn(sd, key + ".1", x)


def _basic_block(sd, key, x, stride, dilation):
    """BasicBlock.forward -- cmfsm.py:76-85 (no ReLU after the residual add)."""
    out = F.relu(_convgn2d(sd, key + ".conv1.0", x, stride, 1, dilation))
    out = _convgn2d(sd, key + ".conv2", out, 1, 1, dilation)
    if (key + ".downsample.0.weight") in sd:
        x = F.conv2d(x, sd[key + ".downsample.0.weight"], None, stride)
        x = _gn(sd, key + ".downsample.1", x)
    return out + x


def _layer(sd, key, x, blocks, stride, dilation):
    """feature_extraction._make_layer -- cmfsm.py:177-197."""
    for i in range(blocks):
        x = _basic_block(sd, "%s.%d" % (key, i), x, stride if i == 0 else 1, dilation)
    return x


# ----------------------------------------------------------------------------------------------
# a2: 2D feature extraction  (cmfsm.py:126-236)
# ----------------------------------------------------------------------------------------------
def feature_extraction(sd, x, prefix="feature_extraction", stages=None, dilations=(1, 2)):
    """Returns (feature [B,32,H/4,W/4], all_feature [B,32,H,W]).  cmfsm.py:199-236.

    `all_feature` is the output of firstconv (pre-GN, pre-ReLU) -- cmfsm.py:138,200.
    The reference's second return value (layer1 output) is unused by cmfsm.forward.
    """
    p = prefix
    o = F.relu(_convgn2d(sd, p + ".firstconv.0", x))
    o = F.relu(_convgn2d(sd, p + ".firstconv.2", o))
    o = F.relu(_convgn2d(sd, p + ".firstconv.4", o))
    all_feature = F.conv2d(o, sd[p + ".firstconv.6.weight"], None, 1, 1)
    o = F.relu(_gn(sd, p + ".secondconv.0", all_feature))
    o = F.relu(_convgn2d(sd, p + ".secondconv.2", o, stride=2))
    o = F.relu(_convgn2d(sd, p + ".secondconv.4", o))
    second = o
    l1 = _layer(sd, p + ".layer1", o, 3, 1, 1)
    raw = _layer(sd, p + ".layer2", l1, 16, 2, 1)
    l3 = _layer(sd, p + ".layer3", raw, 3, 1, dilations[0])  # cm_sub_4 uses dilations (2, 4), cm_sub_4.py:148-149
    skip = _layer(sd, p + ".layer4", l3, 3, 1, dilations[1])
    size = skip.shape[2:]
    branches = []
    for name, k in (("branch1", 64), ("branch2", 32), ("branch3", 16), ("branch4", 8)):
        b = F.avg_pool2d(skip, (k, k), (k, k))
        b = F.relu(_convgn2d(sd, "%s.%s.1" % (p, name), b, 1, 0, 1))
        branches.append(F.interpolate(b, size, mode="bilinear", align_corners=False))
    b1, b2, b3, b4 = branches
    cat = torch.cat((raw, skip, b4, b3, b2, b1), 1)  # cmfsm.py:231-233
    o = F.relu(_convgn2d(sd, p + ".lastconv.0", cat))
    feat = F.conv2d(o, sd[p + ".lastconv.2.weight"])
    if stages is not None:
        stages.update(full=all_feature, second=second, layer1=l1, layer2=raw, layer3=l3, layer4=skip, b1=b1, b2=b2,
                      b3=b3, b4=b4, last0=o, feat=feat)
    return feat, all_feature


# ----------------------------------------------------------------------------------------------
# a3-a5: context-mapping weights (cmfsm.py:391-593), closed form of SURVEY.md appendix A.3
# ----------------------------------------------------------------------------------------------
# neighbour order of the returned 9 maps (cmfsm.py:551): c, l, r, t, b, lt, rt, lb, rb
NEIGHBOUR_DYDX = ((0, 0), (0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1))


def position_code(scale, dtype=torch.float32):
    """The positional codes fed as channels 64,65 for each neighbour k: code[k, 0:2, y%s, x%s].

    matrix_generation (cmfsm.py:391-428): base tile channel0[y,x] = off[x], channel1[y,x] = off[y],
    off = [-s/2..-1, 1..s/2]; distance_matrix1/2 overwrite channel 0 with dec/inc over x,
    distance_matrix3/4 overwrite channel 1 with dec/inc over y.  forward() re-binds matrices 5..8 to
    matrices 1..4 (cmfsm.py:459-462), so the diagonal neighbours reuse the axis encodings.
    """
    s = scale
    half = s // 2
    off = torch.tensor([float(v) for v in list(range(-half, 0)) + list(range(1, half + 1))], dtype=dtype)
    inc = torch.arange(1, s + 1, dtype=dtype)
    dec = s - inc + 1
    ch0 = {"off": off.view(1, s).expand(s, s), "inc": inc.view(1, s).expand(s, s), "dec": dec.view(1, s).expand(s, s)}
    ch1 = {"off": off.view(s, 1).expand(s, s), "inc": inc.view(s, 1).expand(s, s), "dec": dec.view(s, 1).expand(s, s)}
    # (channel-0 kind over x, channel-1 kind over y) per neighbour, SURVEY.md A.3 table
    kinds = (("off", "off"), ("dec", "off"), ("inc", "off"), ("off", "dec"), ("off", "inc"),
             ("dec", "off"), ("inc", "off"), ("off", "dec"), ("off", "inc"))
    return torch.stack([torch.stack((ch0[a], ch1[b])) for a, b in kinds])  # [9,2,s,s]


def similarity_mlp(sd, x, prefix="mapping_matrix.similarity1"):
    """similarity_measure1.forward -- cmfsm.py:334-358: 1x1 convs 66-32-16-8-1, LeakyReLU(0.01) x3."""
    x = F.leaky_relu(F.conv2d(x, sd[prefix + ".conv0.weight"]), 0.01)
    x = F.leaky_relu(F.conv2d(x, sd[prefix + ".conv1.weight"]), 0.01)
    x = F.leaky_relu(F.conv2d(x, sd[prefix + ".conv2.weight"]), 0.01)
    return F.conv2d(x, sd[prefix + ".conv3.weight"])


def context_mapping_weights(sd, lr, hr):
    """eight_related_context_mapping.forward -- cmfsm.py:443-593.  Returns [B,9,H,W] softmax weights.

    logit_k[y,x] = MLP(cat[lr[:, :, y//s+dy, x//s+dx], hr[:, :, y, x], code_k[:, y%s, x%s]]) when the
    low-res cell is inside the image, else the constant -100 (cmfsm.py:451-452).
    """
    B, C, h, w = lr.shape
    H, W = hr.shape[2:]
    s = W // w
    if s % 2 != 0:
        raise ValueError("odd scale (reference calls exit(), cmfsm.py:448-449)")
    codes = position_code(s, lr.dtype)
    lr_up = lr.repeat_interleave(s, 2).repeat_interleave(s, 3)  # cmfsm.py:465-468
    logits = []
    for k, (dy, dx) in enumerate(NEIGHBOUR_DYDX):
        # destination window (pixels whose neighbour cell is in bounds) and matching source window
        y0, y1 = max(0, -dy) * s, H - max(0, dy) * s
        x0, x1 = max(0, -dx) * s, W - max(0, dx) * s
        code = codes[k].repeat(1, H // s, W // s).unsqueeze(0).expand(B, -1, -1, -1)
        # the reference slices the *code* with the lr window (cmfsm.py:484: distance_matrix1[:,:,:,:-scale]);
        # tiles repeat every s pixels and windows move by multiples of s, so both views agree.
        rep = torch.cat([lr_up[:, :, y0 + dy * s:y1 + dy * s, x0 + dx * s:x1 + dx * s],
                         hr[:, :, y0:y1, x0:x1], code[:, :, y0:y1, x0:x1]], 1)
        lg = hr.new_full((B, 1, H, W), -100.0)
        lg[:, :, y0:y1, x0:x1] = similarity_mlp(sd, rep)
        logits.append(lg)
    return F.softmax(torch.cat(logits, 1), dim=1)  # cmfsm.py:551-552


# ----------------------------------------------------------------------------------------------
# a6: concat cost volume (cmfsm.py:667-682), bit-exact contract of SURVEY.md appendix A.1
# ----------------------------------------------------------------------------------------------
def cost_volume_concat(L, R, D):
    """cost[b,c,d,y,x] = L[b,c,y,x] if x>=d else +0.0 ; cost[b,C+c,d,y,x] = R[b,c,y,x-d] if x>=d else +0.0."""
    B, C, h, w = L.shape
    cost = L.new_zeros((B, 2 * C, D, h, w))
    for d in range(D):
        if d >= w:
            break
        cost[:, :C, d, :, d:] = L[:, :, :, d:]
        cost[:, C:, d, :, d:] = R[:, :, :, :w - d]
    return cost


def cost_volume_corr(L, R, D, normalize=False):
    """Correlation form of K1 -- PARITY UNPINNED: the reference has no live correlation path.  Restates the cosine
    matching of the dead file "cmf/models/rstereo # dense volume match.py":309-311 (nn.CosineSimilarity(dim=1) of the
    left features with shifted right features) with the shift / mask convention of the concat volume above:
        corr[b,d,y,x] = mean_c L[b,c,y,x] * R[b,c,y,x-d]          (x >= d, +0.0 otherwise)
        normalize:      sum_c L R / max(|L| |R|, 1e-8)            (torch.nn.functional.cosine_similarity, eps 1e-8)"""
    B, C, h, w = L.shape
    out = L.new_zeros((B, D, h, w))
    for d in range(min(D, w)):
        l, r = L[:, :, :, d:], R[:, :, :, :w - d]
        if normalize:
            out[:, d, :, d:] = (l * r).sum(1) / (l.norm(dim=1) * r.norm(dim=1)).clamp_min(1e-8)
        else:
            out[:, d, :, d:] = (l * r).mean(1)
    return out


def masked_smooth_l1(outputs, disparity, maxdisp=192, weights=(0.5, 0.7, 1.0)):
    """The training loss of the reference drivers, train.py:162-174 (mask, three masked means, weighted sum)."""
    mask = (disparity < maxdisp) & (disparity > 0)
    loss = 0.0
    for wgt, o in zip(weights, outputs):
        o = torch.squeeze(o, 1)
        loss = loss + wgt * F.smooth_l1_loss(o[mask], disparity[mask], reduction="mean")
    return loss


def cost_volume_concat_bwd(g, C):
    """Adjoint of cost_volume_concat: dL[x] = sum_{d<=x} g[c,d,x]; dR[x] = sum_{d,x+d<w} g[C+c,d,x+d]."""
    B, _, D, h, w = g.shape
    dL = g.new_zeros((B, C, h, w))
    dR = g.new_zeros((B, C, h, w))
    for d in range(min(D, w)):
        dL[:, :, :, d:] += g[:, :C, d, :, d:]
        dR[:, :, :, :w - d] += g[:, C:, d, :, d:]
    return dL, dR


# ----------------------------------------------------------------------------------------------
# a7-a9: 3D aggregation (cmfsm.py:240-303, 604-634, 684-695, 724, 747), layer table SURVEY.md A.2
# ----------------------------------------------------------------------------------------------
def hourglass(sd, key, x, presqu, postsqu):
    """hourglass.forward -- cmfsm.py:283-303."""
    out = F.relu(_convgn3d(sd, key + ".conv1.0", x, 2))
    pre = _convgn3d(sd, key + ".conv2", out)
    pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
    out = F.relu(_convgn3d(sd, key + ".conv3.0", pre, 2))
    out = F.relu(_convgn3d(sd, key + ".conv4.0", out))
    c5 = _deconvgn3d(sd, key + ".conv5", out)
    post = F.relu(c5 + (presqu if presqu is not None else pre))
    out = _deconvgn3d(sd, key + ".conv6", post)
    return out, pre, post


def classif(sd, key, x):
    """classifN -- cmfsm.py:621-634: convbn_3d + ReLU + Conv3d(32,1).  Returns [B,D,h,w]."""
    x = F.relu(_convgn3d(sd, key + ".0", x))
    return F.conv3d(x, sd[key + ".2.weight"], None, 1, 1).squeeze(1)


def aggregation3d(sd, cost, stages=None):
    """cmfsm.py:684-695,724,747.  Returns the three RAW classifier volumes (c1,c2,c3), each [B,D,h,w].

    (The cumulative sums cost2=c2+cost1, cost3=c3+cost2 of cmfsm.py:725,748 happen in softargmin_ctxmap.)
    """
    cost0 = F.relu(_convgn3d(sd, "dres0.0", cost))
    cost0 = F.relu(_convgn3d(sd, "dres0.2", cost0))
    t = F.relu(_convgn3d(sd, "dres1.0", cost0))
    cost0 = _convgn3d(sd, "dres1.2", t) + cost0
    out1, pre1, post1 = hourglass(sd, "dres2", cost0, None, None)
    out1 = out1 + cost0
    out2, _pre2, post2 = hourglass(sd, "dres3", out1, pre1, post1)
    out2 = out2 + cost0
    out3, _pre3, _post3 = hourglass(sd, "dres4", out2, pre1, post2)  # note pre1 (cmfsm.py:692)
    out3 = out3 + cost0
    c1 = classif(sd, "classif1", out1)
    c2 = classif(sd, "classif2", out2)
    c3 = classif(sd, "classif3", out3)
    if stages is not None:
        stages.update(cost0=cost0, out1=out1, out2=out2, out3=out3, pre1=pre1, post1=post1, post2=post2)
    return c1, c2, c3


# ----------------------------------------------------------------------------------------------
# a10-a11: soft-argmin + x-scale upsample + 9-neighbour mapping (cmfsm.py:705-769), SURVEY.md A.4
# ----------------------------------------------------------------------------------------------
def softargmin(cost):
    """F.softmax(dim=1) + disparityregression -- cmfsm.py:705-706,111-123.  [B,D,h,w] -> [B,h,w]."""
    D = cost.shape[1]
    disp = torch.arange(D, dtype=cost.dtype).view(1, D, 1, 1)
    return torch.sum(F.softmax(cost, dim=1) * disp, 1)


def mapped_upsample(pred, weights, scale):
    """cmfsm.py:709-723 with per-sample semantics.  pred [B,h,w], weights [B,9,H,W].

    out[b,0,y,x] = sum_k w_k[b,y,x] * scale * pred[b, y//s+dy_k, x//s+dx_k] over in-bounds neighbours,
    accumulated in the reference's order (centre, l, r, t, b, lt, rt, lb, rb).
    """
    s = scale
    up = s * pred.repeat_interleave(s, 1).repeat_interleave(s, 2)  # [B,H,W]
    B, H, W = up.shape
    out = up * weights[:, 0]
    for k in range(1, 9):
        dy, dx = NEIGHBOUR_DYDX[k]
        y0, y1 = max(0, -dy) * s, H - max(0, dy) * s
        x0, x1 = max(0, -dx) * s, W - max(0, dx) * s
        out[:, y0:y1, x0:x1] += up[:, y0 + dy * s:y1 + dy * s, x0 + dx * s:x1 + dx * s] * weights[:, k, y0:y1, x0:x1]
    return out.unsqueeze(1)


def softargmin_ctxmap(c1, c2, c3, weights, scale):
    """Three outputs from the raw classifier volumes; cost2=c2+cost1, cost3=c3+cost2 (cmfsm.py:725,748)."""
    cost1 = c1
    cost2 = c2 + cost1
    cost3 = c3 + cost2
    return tuple(mapped_upsample(softargmin(c), weights, scale) for c in (cost1, cost2, cost3))


# ----------------------------------------------------------------------------------------------
# whole forward
# ----------------------------------------------------------------------------------------------
def strip_module_prefix(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def check_shapes(H, W, maxdisp, B=1):
    """SURVEY.md section 0.6."""
    if H % 16 or W % 16:
        raise ValueError("H and W must be multiples of 16, got %dx%d" % (H, W))
    if maxdisp % 16:
        raise ValueError("maxdisp must be a multiple of 16, got %d" % maxdisp)
    if H < 256 or W < 256 or (B * (H // 256) * (W // 256)) < 2:
        raise ValueError("image too small for the 64x64 SPP pooling branch followed by GroupNorm")


def forward(sd, left, right, maxdisp=192, stages=None, grad=False):
    """cmfsm.forward -- cmfsm.py:655-775, per-sample output semantics [B,1,H,W] for each of 3 outputs.

    For B>1 the reference broadcasts to [B,B,H,W]; the diagonal equals this result.
    `grad=True` keeps the autograd graph (every op above is a differentiable torch op), which makes this the
    gradient oracle of the training path as well.
    """
    with torch.set_grad_enabled(grad):
        sd = strip_module_prefix(sd)
        B, _, H, W = left.shape
        check_shapes(H, W, maxdisp, B)
        L, all_l = feature_extraction(sd, left)
        R, _all_r = feature_extraction(sd, right)
        scale = all_l.shape[-1] // L.shape[-1]
        weights = context_mapping_weights(sd, L, all_l)
        cost = cost_volume_concat(L, R, maxdisp // scale)
        c1, c2, c3 = aggregation3d(sd, cost, stages)
        outs = softargmin_ctxmap(c1, c2, c3, weights, scale)
        if stages is not None:
            stages.update(L=L, R=R, all_l=all_l, weights=weights, cost=cost, c1=c1, c2=c2, c3=c3)
        return outs
