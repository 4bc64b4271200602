"""`cmfsm_sub_16` -- the 1/16-resolution variant of the context-mapping stereo network on the libcmfb200 kernels.

Drop-in for the reference class `cmf.models.cmfsm_sub_16` (reference cmf/models/cmfsm_sub_16.py, cited `sub16.py:line`):
same constructor, module tree (=> `state_dict` keys and seeded initialisation) and `forward(left, right)`, which
returns three `[B,H,W]` maps (this variant has no channel dimension).

Differences from `cmfsm_sub_8`: layer3 has stride 2 (features at 1/16, D' = maxdisp/16), SPP pools 4/2/16/8,
`lastconv_16` on 384 channels (sub16.py:127-239); the target-image weights of `six_related_context_mapping` are used;
the epilogue maps the upsampled cost VOLUME (five spatial neighbours, then three disparity-axis taps with the target
weights) and regresses over all `maxdisp` planes (sub16.py:760-850) -- one kernel, `cmfb200_volume_mapping_fwd`.
Inference only.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from cmf_b200 import ops
from cmf.models.cmfsm import GN_GROUPS, ResidualUnit, _conv_gn_2d, _conv_gn_3d, cmfsm, hourglass
from cmf.models.cmfsm_sub_8 import six_related_context_mapping


class feature_extraction(nn.Module):
    """Parameter container with the reference's layout (sub16.py:127-197); run by cmfsm_sub_16._features."""

    def __init__(self):
        super().__init__()
        self._width = 32
        relu = lambda: nn.ReLU(inplace=True)  # noqa: E731
        self.firstconv = nn.Sequential(_conv_gn_2d(3, 32, 3, 1, 1, 1), relu(), _conv_gn_2d(32, 32, 3, 1, 1, 1), relu(),
                                       _conv_gn_2d(32, 32, 3, 1, 1, 1), relu(), _conv_gn_2d(32, 32, 3, 1, 1, 1), relu())
        self.secondconv = nn.Sequential(_conv_gn_2d(32, 32, 3, 2, 1, 1), relu(), _conv_gn_2d(32, 32, 3, 1, 1, 1), relu())
        self.layer1 = self._stack(32, 3, 2, 1, 1)
        self.layer2 = self._stack(64, 16, 2, 1, 1)
        self.layer3 = self._stack(128, 3, 2, 1, 1)
        self.layer4 = self._stack(128, 3, 1, 1, 4)
        for i, k in enumerate((4, 2, 16, 8), 1):
            setattr(self, "branch%d" % i,
                    nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)), _conv_gn_2d(128, 32, 1, 1, 0, 1), relu()))
        self.lastconv_16 = nn.Sequential(_conv_gn_2d(384, 128, 3, 1, 1, 1), relu(), nn.Conv2d(128, 32, 1, bias=False))

    def _stack(self, width, n, stride, pad, dilation):
        down = None
        if stride != 1 or self._width != width:
            down = nn.Sequential(nn.Conv2d(self._width, width, 1, stride, bias=False), nn.GroupNorm(GN_GROUPS, width))
        units = [ResidualUnit(self._width, width, stride, down, pad, dilation)]
        self._width = width
        units += [ResidualUnit(width, width, 1, None, pad, dilation) for _ in range(1, n)]
        return nn.Sequential(*units)


class cmfsm_sub_16(cmfsm):
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
        self.mapping_matrix = six_related_context_mapping()
        for m in self.modules():  # sub16.py:705-712
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                n = m.out_channels
                for k in m.kernel_size:
                    n *= k
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
        self._finish_init()

    def _features(self, x):
        fe = self.feature_extraction
        o = x
        for i in (0, 2, 4, 6):
            o = self._cg2(fe.firstconv[i], o, relu=True)
        full = o
        o = self._cg2(fe.secondconv[0], o, relu=True)
        o = self._cg2(fe.secondconv[2], o, relu=True)
        raw = None
        for name in ("layer1", "layer2", "layer3", "layer4"):
            for unit in getattr(fe, name):
                t = self._cg2(unit.conv1[0], o, relu=True)
                skip = o if unit.downsample is None else self._cg2(unit.downsample, o)
                o = self._cg2(unit.conv2, t, residual=skip)
            if name == "layer3":
                raw = o  # sub16.py:207: `output_raw` is re-bound to the layer3 output (128 channels)
        skip = o
        b1, b2, b3, b4 = [self._cg2(getattr(fe, "branch%d" % (i + 1))[1], F.avg_pool2d(skip, k, k).contiguous(), relu=True)
                          for i, k in enumerate((4, 2, 16, 8))]
        cat = ops.spp_upsample_concat_sized(raw, skip, [b4, b3, b2, b1])
        o = self._cg2(fe.lastconv_16[0], cat, relu=True)
        feat, _ = self._c2(fe.lastconv_16[2], o, False)
        return feat, full

    @staticmethod
    def _check(left, right, maxdisp):
        if left.shape != right.shape or left.dim() != 4 or left.shape[1] != 3:
            raise ValueError("expected two [B,3,H,W] images, got %s and %s" % (tuple(left.shape), tuple(right.shape)))
        B, _, H, W = left.shape
        if H % 64 or W % 64:
            raise ValueError("H and W must be multiples of 64 (got %dx%d)" % (H, W))
        if maxdisp % 64:
            raise ValueError("maxdisp must be a multiple of 64 (got %d)" % maxdisp)
        if H < 256 or W < 256 or B * (H // 256) * (W // 256) < 2:
            raise ValueError("image %dx%d (B=%d) is too small for the 16x16 SPP branch + GroupNorm" % (H, W, B))
        if not left.is_cuda:
            raise ops._lib.CmfB200Error("cmfsm_sub_16 runs on CUDA (sm_100a) only; there is no CPU path")

    def forward_row_bands(self, left, right, gather=True):
        raise NotImplementedError("row-band sharding is built for cmfsm only")

    def _forward_body(self, left, right):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("cmfsm_sub_16: only inference is built (wrap the call in torch.no_grad())")
        B = left.shape[0]
        both = torch.cat([left.float(), right.float()], 0).contiguous()
        feat, full = self._features(both)
        lfeat, rfeat = feat[:B].contiguous(), feat[B:].contiguous()
        hr_l, hr_r = full[:B].contiguous(), full[B:].contiguous()
        scale = hr_l.shape[-1] // lfeat.shape[-1]
        D = self.maxdisp // scale
        sim = self.mapping_matrix.similarity1
        ws = (sim.conv0.weight, sim.conv1.weight, sim.conv2.weight, sim.conv3.weight)
        weights5 = ops.ctxmap_weights5(lfeat, hr_l, *ws)
        weights3 = ops.ctxmap_weights3(rfeat, hr_r, *ws)
        if self.aggregation == "bf16":
            c1, c2, c3 = self._aggregate_bf16(lfeat, rfeat, D)
        elif self.aggregation == "fp32":
            c1, c2, c3 = self._aggregate_fp32(lfeat, rfeat, D)
        else:
            raise ValueError("aggregation must be 'fp32' or 'bf16', got %r" % (self.aggregation,))
        return ops.volume_mapping(c1, c2, c3, weights5, weights3, scale)
