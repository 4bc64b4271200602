"""`cmf` -- the registry's first entry (PSMNet-style extractor + super-resolution refinement head) on the libcmfb200
kernels (inference).

Drop-in for the reference class `cmf.models.cmf` (reference cmf/models/cmf.py): same module tree (=> state_dict keys and
seeded initialisation), `forward(left, right)` returns three `[B,1,H,W]` maps.  Differences from `cmfsm`: the stem's
first conv has stride 2 (no full-resolution feature map), the 1/2-resolution `layer1` output and the 1/4 features feed
`super_resolution_refinement` (conv 1->64 on the low-resolution soft-argmin, two transposed-conv stages on
cat([x, feature]), cat with three conv layers of the RGB image, conv 96->96, conv 96->1 + bias, ReLU); no context
mapping.

Every convolution runs on an existing kernel; channel counts the kernels do not have are reached by zero padding:
  * Cin = 1 / 3-with-stride-2 -> input and weight padded to 8 input channels;
  * Cout = 96 / 1 -> weight padded to 128 / 32 output channels, the extra channels are dropped (GroupNorm statistics
    are then taken on the real channels with `gn_stats`);
  * ConvTranspose2d(96->64, k3 s2 p1 op1) -> the 3-D transposed-conv kernel on a depth-1 volume with the 2-D weights in
    the kd = 1 slice (output depth 0 is the 2-D result); its bias is added before the GroupNorm.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from cmf_b200 import ops
from cmf.models.cmfsm import GN_GROUPS, ResidualUnit, _conv_gn_2d, _conv_gn_3d, cmfsm, hourglass


class feature_extraction(nn.Module):
    """Parameter container with the reference's layout (cmf.py feature_extraction)."""

    def __init__(self):
        super().__init__()
        self._width = 32
        relu = lambda: nn.ReLU(inplace=True)  # noqa: E731
        self.firstconv = nn.Sequential(_conv_gn_2d(3, 32, 3, 2, 1, 1), relu(), _conv_gn_2d(32, 32, 3, 1, 1, 1), relu(),
                                       _conv_gn_2d(32, 32, 3, 1, 1, 1), relu())
        self.layer1 = self._stack(32, 3, 1, 1, 1)
        self.layer2 = self._stack(64, 16, 2, 1, 1)
        self.layer3 = self._stack(128, 3, 1, 1, 1)
        self.layer4 = self._stack(128, 3, 1, 1, 2)
        for i, k in enumerate((64, 32, 16, 8), 1):
            setattr(self, "branch%d" % i,
                    nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)), _conv_gn_2d(128, 32, 1, 1, 0, 1), relu()))
        self.lastconv = nn.Sequential(_conv_gn_2d(320, 128, 3, 1, 1, 1), relu(), nn.Conv2d(128, 32, 1, bias=False))

    def _stack(self, width, n, stride, pad, dilation):
        down = None
        if stride != 1 or self._width != width:
            down = nn.Sequential(nn.Conv2d(self._width, width, 1, stride, bias=False), nn.GroupNorm(GN_GROUPS, width))
        units = [ResidualUnit(self._width, width, stride, down, pad, dilation)]
        self._width = width
        units += [ResidualUnit(width, width, 1, None, pad, dilation) for _ in range(1, n)]
        return nn.Sequential(*units)


class super_resolution_refinement(nn.Module):
    """Parameter container of the refinement head (cmf.py super_resolution_refinement(32, 2)); run by cmf._srr."""

    def __init__(self, dis_planes, twice_times):
        super().__init__()
        self.twice_times = twice_times
        relu = lambda: nn.ReLU(inplace=True)  # noqa: E731
        self.conv1 = nn.Sequential(_conv_gn_2d(1, dis_planes * 2, 3, 1, 1, 1), relu())
        self.deconv_module_list = nn.ModuleList()
        for _ in range(twice_times):
            self.deconv_module_list.append(nn.Sequential(nn.ConvTranspose2d(dis_planes * 3, dis_planes * 2, 3, 2, 1, 1),
                                                         nn.GroupNorm(GN_GROUPS, dis_planes * 2), relu()))
        self.rgb_fea = nn.Sequential(_conv_gn_2d(3, dis_planes, 3, 1, 1, 1), relu(),
                                     _conv_gn_2d(dis_planes, dis_planes, 3, 1, 1, 1), relu(),
                                     _conv_gn_2d(dis_planes, dis_planes, 3, 1, 1, 1), relu())
        self.conv2 = nn.Sequential(_conv_gn_2d(dis_planes * 3, dis_planes * 3, 3, 1, 1, 1), relu())
        self.conv_out = nn.Conv2d(dis_planes * 3, 1, 3, 1, 1)
        self.crap = nn.ReLU(inplace=True)


class cmf(cmfsm):
    def __init__(self, maxdisp=192):
        nn.Module.__init__(self)
        self.maxdisp = maxdisp
        self.feature_extraction = feature_extraction()
        relu = lambda: nn.ReLU(inplace=True)  # noqa: E731
        self.dres0 = nn.Sequential(_conv_gn_3d(64, 32), relu(), _conv_gn_3d(32, 32), relu())
        self.dres1 = nn.Sequential(_conv_gn_3d(32, 32), relu(), _conv_gn_3d(32, 32))
        self.dres2 = hourglass(32)
        self.dres3 = hourglass(32)
        self.dres4 = hourglass(32)
        for i in (1, 2, 3):
            setattr(self, "classif%d" % i,
                    nn.Sequential(_conv_gn_3d(32, 32), relu(), nn.Conv3d(32, 1, 3, 1, 1, bias=False)))
        self.srr = super_resolution_refinement(32, 2)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                n = m.out_channels
                for k in m.kernel_size:
                    n *= k
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
        self._finish_init()

    # ---- packed weights with zero padding to a channel count the kernels have
    def _pack_padded(self, conv, cin_to=None, cout_to=None):
        w = conv.weight
        key = (conv._cmf_name, w.device.index, "pad", cin_to, cout_to)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == w._version and hit[1] == w.data_ptr():
            return hit[2]
        wd = w.detach()
        co, ci = wd.shape[:2]
        padded = torch.zeros((cout_to or co, cin_to or ci) + tuple(wd.shape[2:]), device=wd.device, dtype=wd.dtype)
        padded[:co, :ci] = wd
        packed = ops.pack_conv2d_weight(padded)
        self._packed[key] = (w._version, w.data_ptr(), packed)
        return packed

    def _pack_deconv2d(self, deconv):
        w = deconv.weight  # [Cin, Cout, 3, 3]
        key = (deconv._cmf_name, w.device.index, "dc2")
        hit = self._packed.get(key)
        if hit is not None and hit[0] == w._version and hit[1] == w.data_ptr():
            return hit[2]
        w3 = torch.zeros(tuple(w.shape[:2]) + (3, 3, 3), device=w.device, dtype=w.dtype)
        w3[:, :, 1] = w.detach()  # only the kd = 1 slice is populated: depth 0 of the output is the 2-D result
        packed = ops.pack_conv3d_weight(w3, transposed=True)
        self._packed[key] = (w._version, w.data_ptr(), packed)
        return packed

    @staticmethod
    def _pad_channels(x, c):
        if x.shape[1] == c:
            return x
        y = torch.zeros((x.shape[0], c) + tuple(x.shape[2:]), device=x.device, dtype=x.dtype)
        y[:, :x.shape[1]] = x
        return y

    def _gn_act(self, y, gn, relu=True):
        return ops.gn_apply(y, ops.gn_stats(y), gn.weight, gn.bias, None, relu, out=y)

    # ---- extractor: stride-2 stem (cmf.py feature_extraction.forward); returns (feature 1/4, layer1 output 1/2)
    def _features(self, x):
        fe = self.feature_extraction
        c0, g0 = fe.firstconv[0][0], fe.firstconv[0][1]
        y, sums = ops.conv2d(self._pad_channels(x, 8), self._pack_padded(c0, cin_to=8), 3, 2, 1, True)
        o = ops.gn_apply(y, sums, g0.weight, g0.bias, None, True, out=y)
        o = self._cg2(fe.firstconv[2], o, relu=True)
        o = self._cg2(fe.firstconv[4], o, relu=True)
        raw = half = None
        for name in ("layer1", "layer2", "layer3", "layer4"):
            for unit in getattr(fe, name):
                t = self._cg2(unit.conv1[0], o, relu=True)
                skip = o if unit.downsample is None else self._cg2(unit.downsample, o)
                o = self._cg2(unit.conv2, t, residual=skip)
            if name == "layer1":
                half = o
            if name == "layer2":
                raw = o
        skip = o
        pooled = ops.spp_pool(skip)
        b1, b2, b3, b4 = [self._cg2(getattr(fe, "branch%d" % (i + 1))[1], p, relu=True) for i, p in enumerate(pooled)]
        cat = ops.spp_upsample_concat(raw, skip, b4, b3, b2, b1)
        o = self._cg2(fe.lastconv[0], cat, relu=True)
        feat, _ = self._c2(fe.lastconv[2], o, False)
        return feat, half

    # ---- refinement head (cmf.py super_resolution_refinement.forward)
    def _rgb_features(self, rgb):
        s = self.srr
        o = rgb
        for i in (0, 2, 4):
            o = self._cg2(s.rgb_fea[i], o, relu=True)
        return o

    def _srr(self, pred_lr, rgb_fea, feat, half):
        s = self.srr
        c1, g1 = s.conv1[0][0], s.conv1[0][1]
        x, sums = ops.conv2d(self._pad_channels(pred_lr.unsqueeze(1), 8), self._pack_padded(c1, cin_to=8), 3, 1, 1, True)
        x = ops.gn_apply(x, sums, g1.weight, g1.bias, None, True, out=x)
        for stage, z in zip(s.deconv_module_list, (feat, half)):
            deconv, gn = stage[0], stage[1]
            inp = torch.cat([x, z], 1).unsqueeze(2).contiguous()  # [B,96,1,h,w]
            y, _ = ops.conv3d_k3(inp, self._pack_deconv2d(deconv), transposed=True)
            y = y[:, :, 0].contiguous()
            y += deconv.bias.detach().view(1, -1, 1, 1)
            x = self._gn_act(y, gn)
        c2, g2 = s.conv2[0][0], s.conv2[0][1]
        y, _ = ops.conv2d(torch.cat([x, rgb_fea], 1).contiguous(), self._pack_padded(c2, cout_to=128), 3, 1, 1, False)
        x = self._gn_act(y[:, :c2.out_channels].contiguous(), g2)
        y, _ = ops.conv2d(x, self._pack_padded(s.conv_out, cout_to=32), 3, 1, 1, False)
        return F.relu(y[:, :1] + s.conv_out.bias.detach().view(1, 1, 1, 1))

    def forward_row_bands(self, left, right, gather=True):
        raise NotImplementedError("row-band sharding is built for cmfsm only")

    def _forward_body(self, left, right):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("cmf: only inference is built (wrap the call in torch.no_grad())")
        B, _, H, W = left.shape
        left = left.float().contiguous()
        both = torch.cat([left, right.float()], 0).contiguous()
        feat, half = self._features(both)
        lfeat, rfeat = feat[:B].contiguous(), feat[B:].contiguous()
        D = self.maxdisp // 4
        if self.aggregation == "bf16":
            c1, c2, c3 = self._aggregate_bf16(lfeat, rfeat, D)
        elif self.aggregation == "fp32":
            c1, c2, c3 = self._aggregate_fp32(lfeat, rfeat, D)
        else:
            raise ValueError("aggregation must be 'fp32' or 'bf16', got %r" % (self.aggregation,))
        # cumulative soft-argmin at 1/4 resolution = phase 1 of K4 (the mapped outputs are not needed here)
        dummy = torch.zeros((B, 9, H, W), device=left.device, dtype=torch.float32)
        _, low = ops.softargmin_ctxmap(c1, c2, c3, dummy, 4, want_lowres=True)
        rgb_fea = self._rgb_features(left)
        half_l = half[:B].contiguous()
        return tuple(self._srr(low[i], rgb_fea, lfeat, half_l) for i in range(3))
