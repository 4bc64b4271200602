"""Differentiable wrappers used when `cmfsm.forward` runs under autograd (training, train.py:166-181).

FORWARD always runs the libcmfb200 kernels (2-D extractor included).  BACKWARD: no cuDNN / ATen convolution anywhere:
  * cost volume           -- own kernel (`cmfb200_cost_volume_concat_bwd`);
  * GroupNorm (+ReLU mask, +residual gradient) -- own kernels (`cmfb200_gn_bwd`);
  * classifier 32->1 conv -- own dgrad/wgrad kernels (`cmfb200_conv3d_cout1_bwd`);
  * conv/deconv dgrad     -- the forward FFMA kernels with re-arranged weights (`ops.conv3d_dgrad`, `ops.conv2d_dgrad`:
    stride 1 = flipped taps, stride 2 = the transposed-conv kernel, transposed = the stride-2 kernel), strict fp32;
  * conv/deconv wgrad     -- own voxel-reduction kernel (`cmfb200_conv_wgrad`), strict fp32;
  * SPP upsample + concat -- gradient slices + two dense products per branch (adjoint of the bilinear map);
  * K5                    -- own kernel (`cmfb200_ctxmap_weights_bwd`) + two 1x1 GEMMs; the PyTorch closed form below
    is the test reference and the path for scales other than 4;
  * K4                    -- own kernel (`cmfb200_softargmin_ctxmap_bwd`); the PyTorch closed form below is its test
    reference.
None of this is reachable on CPU tensors (the forward kernels raise first).
"""
import os

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import ops

# "tc3": stride-1 convs of forward and dgrad on the tensor cores (fp32-accurate three-term split), "ffma": CUDA cores
ENGINE = os.environ.get("CMF_B200_CONV", "tc3")

_NEIGHBOURS = ((0, 0), (0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1))


class _CostVolume(Function):
    @staticmethod
    def forward(ctx, L, R, D):
        ctx.C = L.shape[1]
        return ops.cost_volume_concat(L.contiguous(), R.contiguous(), D)

    @staticmethod
    def backward(ctx, g):
        dL, dR = ops.cost_volume_concat_bwd(g.contiguous(), ctx.C)
        return dL, dR, None


def cost_volume_concat(L, R, D):
    return _CostVolume.apply(L, R, D)


class _ConvGN3d(Function):
    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, stride, transposed, relu):
        x = x.contiguous()
        if stride == 1 and not transposed:
            raw, sums = ops.conv_s1_nchw(x, weight, 1, True, ENGINE)
        else:
            raw, sums = ops.conv3d_k3(x, ops.pack_conv3d_weight(weight, transposed), stride, transposed, want_stats=True)
        res = residual.contiguous() if residual is not None else None
        out = ops.gn_apply(raw, sums, gamma, beta, res, relu)
        ctx.cfg = (stride, transposed, relu, residual is not None)
        ctx.save_for_backward(x, weight, gamma, raw, sums, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        stride, transposed, relu, has_res = ctx.cfg
        x, weight, gamma, raw, sums, out = ctx.saved_tensors
        d_raw, d_gamma, d_beta, d_res = ops.gn_backward(g, raw, sums, gamma, out if relu else None, has_res)
        dx = ops.conv3d_dgrad(d_raw, weight, stride, transposed, ENGINE) if ctx.needs_input_grad[0] else None
        if transposed:  # roles swapped: see cmfb200_conv_wgrad
            dw = ops.conv_wgrad(d_raw, x, 3, 2)
        else:
            dw = ops.conv_wgrad(x, d_raw, 3, stride)
        return dx, dw, d_gamma, d_beta, d_res, None, None, None


def conv3d_gn(x, weight, gamma, beta, stride=1, transposed=False, residual=None, relu=False):
    return _ConvGN3d.apply(x, weight, gamma, beta, residual, stride, transposed, relu)


class _ConvPlain3d(Function):
    @staticmethod
    def forward(ctx, x, weight):
        x = x.contiguous()
        y, _ = ops.conv3d_k3(x, ops.pack_conv3d_weight(weight, False), 1)
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        if weight.shape[0] == 1 and weight.shape[1] == 32:  # classifier tail: own kernels
            return ops.conv3d_cout1_backward(x, weight, g)
        return ops.conv3d_dgrad(g, weight, 1, False, ENGINE), ops.conv_wgrad(x, g, 3, 1)


def conv3d_plain(x, weight):
    return _ConvPlain3d.apply(x, weight)


class _ConvGN2d(Function):
    """2-D conv + GroupNorm (+residual) (+ReLU) of the feature extractor (convbn / BasicBlock, cmfsm.py:36-85)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, stride, dilation, relu):
        x = x.contiguous()
        k = weight.shape[-1]
        if stride == 1:
            raw, sums = ops.conv_s1_nchw(x, weight, dilation, True, ENGINE)
        else:
            raw, sums = ops.conv2d(x, ops.pack_conv2d_weight(weight), k, stride, dilation, want_stats=True)
        res = residual.contiguous() if residual is not None else None
        out = ops.gn_apply(raw, sums, gamma, beta, res, relu)
        ctx.cfg = (k, stride, dilation, relu, residual is not None)
        ctx.save_for_backward(x, weight, gamma, raw, sums, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        k, stride, dilation, relu, has_res = ctx.cfg
        x, weight, gamma, raw, sums, out = ctx.saved_tensors
        d_raw, d_gamma, d_beta, d_res = ops.gn_backward(g, raw, sums, gamma, out if relu else None, has_res)
        dx = ops.conv2d_dgrad(d_raw, weight, x.shape, stride, dilation, ENGINE) if ctx.needs_input_grad[0] else None
        dw = ops.conv_wgrad(x, d_raw, k, stride, dilation)
        return dx, dw, d_gamma, d_beta, d_res, None, None, None


def conv2d_gn(x, weight, gamma, beta, stride=1, dilation=1, residual=None, relu=False):
    return _ConvGN2d.apply(x, weight, gamma, beta, residual, stride, dilation, relu)


class _ConvPlain2d(Function):
    @staticmethod
    def forward(ctx, x, weight):
        x = x.contiguous()
        y, _ = ops.conv_s1_nchw(x, weight, 1, False, ENGINE)
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        return ops.conv2d_dgrad(g, weight, x.shape, 1, 1, ENGINE), ops.conv_wgrad(x, g.contiguous(), weight.shape[-1], 1, 1)


def conv2d_plain(x, weight):
    return _ConvPlain2d.apply(x, weight)


class _GroupNormAct(Function):
    """Stand-alone GroupNorm (+ReLU): feature_extraction.secondconv[0:2] on the kept full-resolution map."""

    @staticmethod
    def forward(ctx, x, gamma, beta, relu):
        x = x.contiguous()
        sums = ops.gn_stats(x)
        out = ops.gn_apply(x, sums, gamma, beta, None, relu)
        ctx.relu = relu
        ctx.save_for_backward(x, gamma, sums, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, gamma, sums, out = ctx.saved_tensors
        dx, d_gamma, d_beta, _ = ops.gn_backward(g, x, sums, gamma, out if ctx.relu else None, False)
        return dx, d_gamma, d_beta, None


def group_norm_act(x, gamma, beta, relu):
    return _GroupNormAct.apply(x, gamma, beta, relu)


_INTERP = {}


def _bilinear_matrix(n_in, n_out, device):
    """[n_out, n_in] matrix of F.interpolate(mode='bilinear', align_corners=False) along one axis."""
    key = (n_in, n_out, device)
    m = _INTERP.get(key)
    if m is None:
        src = ((torch.arange(n_out, dtype=torch.float32) + 0.5) * (float(n_in) / float(n_out)) - 0.5).clamp_min(0)
        i0 = src.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        l1 = src - i0.float()
        m = torch.zeros(n_out, n_in)
        m.scatter_add_(1, i0.view(-1, 1), (1 - l1).view(-1, 1))
        m.scatter_add_(1, i1.view(-1, 1), l1.view(-1, 1))
        m = m.to(device)
        _INTERP[key] = m
    return m


class _SppUpsampleConcat(Function):
    """cat([raw, skip, up(b4), up(b3), up(b2), up(b1)]) (cmfsm.py:207-233).  Backward: the channel slices of the
    incoming gradient; bilinear upsampling is the separable linear map Uy . b . Ux^T, so its adjoint is two small
    dense products per branch (ATen's upsample_bilinear2d_backward took 11 ms per step here)."""

    @staticmethod
    def forward(ctx, raw, skip, b4, b3, b2, b1):
        ctx.shapes = [tuple(t.shape[2:]) for t in (b4, b3, b2, b1)]
        return ops.spp_upsample_concat(raw.contiguous(), skip.contiguous(), b4.contiguous(), b3.contiguous(),
                                       b2.contiguous(), b1.contiguous())

    @staticmethod
    def backward(ctx, g):
        H, W = g.shape[2:]
        grads = [g[:, :64], g[:, 64:192]]
        for i, (hb, wb) in enumerate(ctx.shapes):
            gi = g[:, 192 + 32 * i:224 + 32 * i]
            uy, ux = _bilinear_matrix(hb, H, g.device), _bilinear_matrix(wb, W, g.device)
            grads.append(torch.matmul(torch.matmul(uy.t(), gi), ux))
        return tuple(grads)


def spp_upsample_concat(raw, skip, b4, b3, b2, b1):
    return _SppUpsampleConcat.apply(raw, skip, b4, b3, b2, b1)


# ---- backward-only closed forms (PyTorch CUDA ops; differentiated by autograd inside backward) -------
def _position_code(scale, device):
    half = scale // 2
    off = torch.tensor([float(v) for v in list(range(-half, 0)) + list(range(1, half + 1))], device=device)
    inc = torch.arange(1, scale + 1, device=device, dtype=torch.float32)
    dec = scale - inc + 1
    rowv = {"off": off, "inc": inc, "dec": dec}
    kinds = (("off", "off"), ("dec", "off"), ("inc", "off"), ("off", "dec"), ("off", "inc"),
             ("dec", "off"), ("inc", "off"), ("off", "dec"), ("off", "inc"))
    return torch.stack([torch.stack((rowv[a].view(1, scale).expand(scale, scale),
                                     rowv[b].view(scale, 1).expand(scale, scale))) for a, b in kinds])


def _ctxmap_weights_torch(lr, hr, w0, w1, w2, w3):
    B, C, h, w = lr.shape
    H, W = hr.shape[2:]
    s = W // w
    codes = _position_code(s, lr.device)
    lr_up = lr.repeat_interleave(s, 2).repeat_interleave(s, 3)
    logits = []
    for k, (dy, dx) in enumerate(_NEIGHBOURS):
        code = codes[k].repeat(1, h, w).unsqueeze(0).expand(B, -1, -1, -1)
        shifted = torch.roll(lr_up, shifts=(-dy * s, -dx * s), dims=(2, 3))
        t = torch.cat([shifted, hr, code], 1)
        t = F.leaky_relu(F.conv2d(t, w0), 0.01)
        t = F.leaky_relu(F.conv2d(t, w1), 0.01)
        t = F.leaky_relu(F.conv2d(t, w2), 0.01)
        t = F.conv2d(t, w3)
        valid = torch.ones((1, 1, H, W), device=lr.device, dtype=torch.bool)
        if dy < 0:
            valid[:, :, :s] = False
        if dy > 0:
            valid[:, :, H - s:] = False
        if dx < 0:
            valid[:, :, :, :s] = False
        if dx > 0:
            valid[:, :, :, W - s:] = False
        logits.append(torch.where(valid, t, torch.full_like(t, -100.0)))
    return F.softmax(torch.cat(logits, 1), dim=1)


def _softargmin_ctxmap_torch(c1, c2, c3, weights9, scale):
    s = scale
    outs = []
    cost = None
    for c in (c1, c2, c3):
        cost = c if cost is None else c + cost
        D = cost.shape[1]
        disp = torch.arange(D, device=cost.device, dtype=cost.dtype).view(1, D, 1, 1)
        p = (F.softmax(cost, 1) * disp).sum(1)
        up = s * p.repeat_interleave(s, 1).repeat_interleave(s, 2)
        H, W = up.shape[1:]
        out = up * weights9[:, 0]
        for k in range(1, 9):
            dy, dx = _NEIGHBOURS[k]
            y0, y1 = max(0, -dy) * s, H - max(0, dy) * s
            x0, x1 = max(0, -dx) * s, W - max(0, dx) * s
            pad = torch.zeros_like(up)
            pad[:, y0:y1, x0:x1] = up[:, y0 + dy * s:y1 + dy * s, x0 + dx * s:x1 + dx * s]
            out = out + pad * weights9[:, k]
        outs.append(out.unsqueeze(1))
    return tuple(outs)


class _CtxmapWeights(Function):
    @staticmethod
    def forward(ctx, lr, hr, w0, w1, w2, w3):
        lr, hr = lr.contiguous(), hr.contiguous()
        out = ops.ctxmap_weights(lr, hr, w0, w1, w2, w3)
        ctx.save_for_backward(lr, hr, w0, w1, w2, w3, out)
        return out

    @staticmethod
    def backward(ctx, g):
        lr, hr, w0, w1, w2, w3, out = ctx.saved_tensors
        if hr.shape[-1] // lr.shape[-1] == 4:  # own kernel (cmfsm: scale 4)
            return ops.ctxmap_weights_bwd(lr, hr, w0, w1, w2, w3, out, g)
        saved = [t.detach().requires_grad_(True) for t in (lr, hr, w0, w1, w2, w3)]
        with torch.enable_grad():
            ref = _ctxmap_weights_torch(*saved)
        return torch.autograd.grad(ref, saved, g, allow_unused=True)


def ctxmap_weights(lr, hr, w0, w1, w2, w3):
    return _CtxmapWeights.apply(lr, hr, w0, w1, w2, w3)


class _SoftargminCtxmap(Function):
    @staticmethod
    def forward(ctx, c1, c2, c3, weights9, scale):
        ctx.scale = scale
        c1, c2, c3, weights9 = c1.contiguous(), c2.contiguous(), c3.contiguous(), weights9.contiguous()
        ctx.save_for_backward(c1, c2, c3, weights9)
        return ops.softargmin_ctxmap(c1, c2, c3, weights9, scale)

    @staticmethod
    def backward(ctx, g1, g2, g3):
        c1, c2, c3, weights9 = ctx.saved_tensors
        gs = [g if g is not None else torch.zeros_like(weights9[:, :1]) for g in (g1, g2, g3)]
        return ops.softargmin_ctxmap_bwd(c1, c2, c3, weights9, gs[0], gs[1], gs[2], ctx.scale) + (None,)


def softargmin_ctxmap(c1, c2, c3, weights9, scale):
    return _SoftargminCtxmap.apply(c1, c2, c3, weights9, scale)


class _MaskedSmoothL1(Function):
    """The reference training loss (train.py:162-174) as two fused kernels: one pass for the three masked sums + the
    valid-pixel count, one pass for the three gradients.  `group`/`distributed`: the count is all-reduced so that every
    rank divides by the GLOBAL number of valid pixels (summing gradients over ranks then reproduces the reference's
    single-process DataParallel mean exactly)."""

    @staticmethod
    def forward(ctx, o1, o2, o3, disp, maxdisp, w1, w2, w3, distributed, group):
        o1, o2, o3, disp = o1.contiguous(), o2.contiguous(), o3.contiguous(), disp.contiguous()
        sums = ops.masked_smooth_l1_sums(o1, o2, o3, disp, maxdisp)
        count = sums[3:4].clone()
        if distributed:
            import torch.distributed as dist

            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
        count = count.clamp_min(1.0)
        wts = torch.tensor([w1, w2, w3], device=disp.device, dtype=torch.float64)
        ctx.maxdisp = maxdisp
        ctx.save_for_backward(o1, o2, o3, disp, (wts / count).float())
        return ((sums[:3] * wts).sum() / count[0]).float()

    @staticmethod
    def backward(ctx, g):
        o1, o2, o3, disp, scale = ctx.saved_tensors
        g1, g2, g3 = ops.masked_smooth_l1_grads(o1, o2, o3, disp, (scale * g).contiguous(), ctx.maxdisp)
        return g1, g2, g3, None, None, None, None, None, None, None


def masked_smooth_l1(outputs, disparity, maxdisp=192, weights=(0.5, 0.7, 1.0), distributed=False, group=None):
    o1, o2, o3 = outputs
    return _MaskedSmoothL1.apply(o1, o2, o3, disparity, float(maxdisp), float(weights[0]), float(weights[1]),
                                 float(weights[2]), bool(distributed), group)
