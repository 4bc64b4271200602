"""`cm_sub_16` -- single-hourglass ablation of `cmfsm_sub_16` on the libcmfb200 kernels (inference).

Drop-in for the reference class `cmf.models.cm_sub_16` (reference cmf/models/cm_sub_16.py): the feature extractor and
`six_related_context_mapping` of cmfsm_sub_16, `dres0`, `dres1`, ONE hourglass (`dres2`) and `classif1`; the single
classifier volume goes through the cost-volume mapping of cmfsm_sub_16 (spatial five-neighbour mix, disparity-axis mix
with the shifted target weights, softmax over maxdisp planes; reference cm_sub_16.py, end of `forward`) and the result
is returned three times (`return pred1, pred1, pred1`), each `[B,H,W]`.  Same module tree => same state_dict keys and
seeded initialisation as the reference.
"""
import math

import torch
import torch.nn as nn

from cmf_b200 import ops
from cmf.models.cmfsm import _conv_gn_3d, hourglass
from cmf.models.cmfsm_sub_8 import six_related_context_mapping
from cmf.models.cmfsm_sub_16 import cmfsm_sub_16, feature_extraction


class cm_sub_16(cmfsm_sub_16):
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

    def _forward_body(self, left, right):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("cm_sub_16: only inference is built (wrap the call in torch.no_grad())")
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
        zero = torch.zeros_like(c1)  # the mapping kernel accumulates c1, c1 + c2, c1 + c2 + c3
        pred1 = ops.volume_mapping(c1, zero, zero, weights5, weights3, scale)[0]
        return pred1, pred1, pred1
