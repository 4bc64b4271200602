"""`cm_sub_4` -- single-hourglass ablation at 1/4 resolution on the libcmfb200 kernels (inference).

Drop-in for the reference class `cmf.models.cm_sub_4` (reference cmf/models/cm_sub_4.py): cmfsm's feature extractor with
dilations 2 / 4 in layer3 / layer4 (`:148-149`), `six_related_context_mapping` (five reference-image + three
target-image weights; its `similarity2` / `fuse` sub-modules exist but are unused), `dres0`, `dres1`, ONE hourglass,
`classif1`, then the cost-volume mapping of cmfsm_sub_16 at scale 4; the prediction is returned three times, `[B,H,W]`.
Same module tree => same state_dict keys and seeded initialisation as the reference.
"""
import math

import torch
import torch.nn as nn

from cmf_b200 import ops
from cmf.models.cmfsm import GN_GROUPS, _conv_gn_2d, _conv_gn_3d, cmfsm, hourglass
from cmf.models.cmfsm import feature_extraction as _cmfsm_feature_extraction
from cmf.models.cmfsm_sub_8 import similarity_measure1


class feature_extraction(_cmfsm_feature_extraction):
    """cmfsm's extractor, layer3 / layer4 with dilation 2 / 4 (built after the parent's layers, in the same order)."""

    def __init__(self):
        nn.Module.__init__(self)
        self._width = 32
        relu = lambda: nn.ReLU(inplace=True)  # noqa: E731
        c2 = _conv_gn_2d
        self.firstconv = nn.Sequential(c2(3, 32, 3, 1, 1, 1), relu(), c2(32, 32, 3, 1, 1, 1), relu(), c2(32, 32, 3, 1, 1, 1),
                                       relu(), nn.Conv2d(32, 32, 3, 1, 1, bias=False))
        self.secondconv = nn.Sequential(nn.GroupNorm(GN_GROUPS, 32), relu(), c2(32, 32, 3, 2, 1, 1), relu(),
                                        c2(32, 32, 3, 1, 1, 1), relu())
        self.layer1 = self._stack(32, 3, 1, 1, 1)
        self.layer2 = self._stack(64, 16, 2, 1, 1)
        self.layer3 = self._stack(128, 3, 1, 1, 2)
        self.layer4 = self._stack(128, 3, 1, 1, 4)
        for i, k in enumerate((64, 32, 16, 8), 1):
            setattr(self, "branch%d" % i, nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)), c2(128, 32, 1, 1, 0, 1), relu()))
        self.lastconv = nn.Sequential(c2(320, 128, 3, 1, 1, 1), relu(), nn.Conv2d(128, 32, 1, bias=False))


class similarity_measure2(nn.Module):
    """Unused 3-3-2-1 MLP of the reference's mapping module (cm_sub_4.py); kept for the state_dict / RNG contract."""

    def __init__(self):
        super().__init__()
        self.conv0 = nn.Conv2d(3, 3, 1, bias=False)
        self.conv1 = nn.Conv2d(3, 2, 1, bias=False)
        self.conv2 = nn.Conv2d(2, 1, 1, bias=False)
        for m in (self.conv0, self.conv1, self.conv2):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class six_related_context_mapping(nn.Module):
    def __init__(self):
        super().__init__()
        self.similarity1 = similarity_measure1()
        self.similarity2 = similarity_measure2()
        self.fuse = nn.Sequential(nn.Conv2d(2, 1, 1, bias=False), nn.LeakyReLU(inplace=True))


class cm_sub_4(cmfsm):
    def __init__(self, maxdisp=192):
        nn.Module.__init__(self)
        self.maxdisp = maxdisp
        self.feature_extraction = feature_extraction()
        relu = lambda: nn.ReLU(inplace=True)  # noqa: E731
        self.dres0 = nn.Sequential(_conv_gn_3d(64, 32), relu(), _conv_gn_3d(32, 32), relu())
        self.dres1 = nn.Sequential(_conv_gn_3d(32, 32), relu(), _conv_gn_3d(32, 32))
        self.dres2 = hourglass(32)
        self.classif1 = nn.Sequential(_conv_gn_3d(32, 32), relu(), nn.Conv3d(32, 1, 3, 1, 1, bias=False))
        self.mapping_matrix = six_related_context_mapping()
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                n = m.out_channels
                for k in m.kernel_size:
                    n *= k
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
        self._finish_init()

    def forward_row_bands(self, left, right, gather=True):
        raise NotImplementedError("row-band sharding is built for cmfsm only")

    def _forward_body(self, left, right):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("cm_sub_4: only inference is built (wrap the call in torch.no_grad())")
        B = left.shape[0]
        both = torch.cat([left.float(), right.float()], 0).contiguous()
        feat, full = self._features(both)
        lfeat, rfeat = feat[:B].contiguous(), feat[B:].contiguous()
        hr_l, hr_r = full[:B].contiguous(), full[B:].contiguous()
        scale = hr_l.shape[-1] // lfeat.shape[-1]
        sim = self.mapping_matrix.similarity1
        ws = (sim.conv0.weight, sim.conv1.weight, sim.conv2.weight, sim.conv3.weight)
        weights5 = ops.ctxmap_weights5(lfeat, hr_l, *ws)
        weights3 = ops.ctxmap_weights3(rfeat, hr_r, *ws)
        if self.aggregation == "bf16":
            (c1,) = self._aggregate_bf16(lfeat, rfeat, self.maxdisp // scale)
        elif self.aggregation == "fp32":
            (c1,) = self._aggregate_fp32(lfeat, rfeat, self.maxdisp // scale)
        else:
            raise ValueError("aggregation must be 'fp32' or 'bf16', got %r" % (self.aggregation,))
        zero = torch.zeros_like(c1)
        pred1 = ops.volume_mapping(c1, zero, zero, weights5, weights3, scale)[0]
        return pred1, pred1, pred1
