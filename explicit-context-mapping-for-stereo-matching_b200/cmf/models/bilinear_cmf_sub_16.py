"""`bilinear_cmf_sub_16` -- the no-mapping baseline at 1/16 resolution on the libcmfb200 kernels (inference).

Drop-in for the reference class `cmf.models.bilinear_cmf_sub_16` (reference cmf/models/bilinear_cmf_sub_16.py): the feature extractor of
`cmfsm_sub_16`, the cmfsm 3-D aggregation, cumulative classifier volumes, trilinear upsampling to `[maxdisp, H, W]`,
softmax and soft-argmin regression -- no context mapping (the PSMNet-style baseline of the paper).  The epilogue is one
kernel, `cmfb200_trilinear_softargmin_fwd`.  Returns three `[B,H,W]` maps.  Same module tree => same state_dict keys and
seeded initialisation as the reference.
"""
import math

import torch
import torch.nn as nn

from cmf_b200 import ops
from cmf.models.cmfsm import _conv_gn_3d, hourglass
from cmf.models.cmfsm_sub_16 import cmfsm_sub_16, feature_extraction


class bilinear_cmf_sub_16(cmfsm_sub_16):
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
            raise NotImplementedError("bilinear_cmf_sub_16: only inference is built (wrap the call in torch.no_grad())")
        B, _, H, W = left.shape
        both = torch.cat([left.float(), right.float()], 0).contiguous()
        feat, _full = self._features(both)
        lfeat, rfeat = feat[:B].contiguous(), feat[B:].contiguous()
        D = self.maxdisp // (W // lfeat.shape[-1])
        if self.aggregation == "bf16":
            c1, c2, c3 = self._aggregate_bf16(lfeat, rfeat, D)
        elif self.aggregation == "fp32":
            c1, c2, c3 = self._aggregate_fp32(lfeat, rfeat, D)
        else:
            raise ValueError("aggregation must be 'fp32' or 'bf16', got %r" % (self.aggregation,))
        return ops.trilinear_softargmin(c1, c2, c3, self.maxdisp, H, W)
