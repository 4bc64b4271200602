"""GPU parity tests: every libcmfb200 kernel (through the C ABI via cmf_b200.ops) against the CPU oracle
on the same seeded inputs, against fixtures produced by the real reference, and -- at BASELINE sizes --
through size-independent properties."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cmfsm_oracle as orc
import golden_common as gc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _npz(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("B,C,h,w,D", [(1, 32, 64, 128, 48), (2, 32, 7, 20, 12), (1, 4, 3, 8, 12), (1, 32, 5, 36, 1),
                                       (3, 2, 9, 4, 3)])
def test_k1_bit_exact_vs_oracle(B, C, h, w, D):
    from cmf_b200 import ops

    L, R = _rand(B, C, h, w, seed=1), _rand(B, C, h, w, seed=2)
    want = orc.cost_volume_concat(L, R, D)
    got = ops.cost_volume_concat(L.to(DEV), R.to(DEV), D).cpu()
    assert torch.equal(got, want)
    assert not torch.signbit(got[got == 0]).any()  # +0.0 everywhere in the masked triangle


def test_k1_golden_crop_from_reference(golden_dir):
    from cmf_b200 import ops

    g = _npz(golden_dir, "cmfsm_c1_full.npz")
    L, R = g["k1_L_crop"].unsqueeze(0), g["k1_R_crop"].unsqueeze(0)
    got = ops.cost_volume_concat(L.to(DEV).contiguous(), R.to(DEV).contiguous(), 48).cpu()
    assert torch.equal(got[0], g["k1_cost_crop"])


def test_k1_full_size_properties():
    """BASELINE config 2 (576x960 -> 144x240, D'=48): slice identities of SURVEY.md A.1, checked on device."""
    from cmf_b200 import ops

    B, C, h, w, D = 1, 32, 144, 240, 48
    L = torch.randn(B, C, h, w, device=DEV)
    R = torch.randn(B, C, h, w, device=DEV)
    cost = ops.cost_volume_concat(L, R, D)
    assert cost.shape == (B, 2 * C, D, h, w)
    for d in (0, 1, 7, 47):
        assert torch.equal(cost[:, :C, d, :, d:], L[..., d:])
        assert torch.equal(cost[:, C:, d, :, d:], R[..., :w - d])
        assert int(cost[:, :, d, :, :d].count_nonzero()) == 0
    # checksum of checksums: every L element appears min(x+1, D) times
    mult = torch.clamp(torch.arange(w, device=DEV) + 1, max=D).double()
    assert abs(float(cost[:, :C].double().sum()) - float((L.double() * mult).sum())) < 1e-6 * cost[:, :C].numel()


def test_k1_backward_vs_oracle():
    from cmf_b200 import ops

    g = _rand(2, 8, 6, 5, 16, seed=3)
    dL, dR = ops.cost_volume_concat_bwd(g.to(DEV), 4)
    wL, wR = orc.cost_volume_concat_bwd(g, 4)
    torch.testing.assert_close(dL.cpu(), wL, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dR.cpu(), wR, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------ K2 / K3
CONV_CASES = [  # B, Cin, Cout, D, H, W, stride
    (1, 64, 32, 6, 8, 20, 1), (2, 32, 32, 5, 9, 36, 1), (1, 32, 64, 8, 8, 16, 2), (1, 32, 64, 7, 9, 18, 2),
    (1, 64, 64, 6, 10, 40, 2), (1, 64, 64, 3, 5, 34, 1), (2, 32, 1, 6, 8, 12, 1), (1, 32, 1, 9, 17, 70, 1), (2, 32, 1, 17, 19, 72, 1),
    (1, 32, 32, 4, 16, 64, 1),
]


@pytest.mark.parametrize("B,Cin,Cout,D,H,W,stride", CONV_CASES)
def test_conv3d_vs_fp64_oracle(B, Cin, Cout, D, H, W, stride):
    from cmf_b200 import ops

    x = _rand(B, Cin, D, H, W, seed=10)
    wgt = _rand(Cout, Cin, 3, 3, 3, seed=11) * (2.0 / (27 * Cin)) ** 0.5
    want = F.conv3d(x.double(), wgt.double(), None, stride, 1)
    y, sums = ops.conv3d_k3(x.to(DEV), ops.pack_conv3d_weight(wgt.to(DEV)), stride, want_stats=Cout > 1)
    assert y.shape == want.shape
    assert _rel_l2(y, want) < 1e-5  # fp32 FMA accumulation vs fp64 (tolerance: SURVEY.md 8c, K2 fp32 mode)
    if sums is not None:
        s = sums.cpu()
        torch.testing.assert_close(s[..., 0], want.sum((2, 3, 4)), rtol=1e-5, atol=1e-3)
        torch.testing.assert_close(s[..., 1], (want * want).sum((2, 3, 4)), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("B,Cin,Cout,D,H,W", [(1, 64, 64, 3, 4, 6), (1, 64, 32, 4, 5, 34), (2, 64, 32, 2, 8, 32)])
def test_deconv3d_vs_fp64_oracle(B, Cin, Cout, D, H, W):
    from cmf_b200 import ops

    x = _rand(B, Cin, D, H, W, seed=12)
    wgt = _rand(Cin, Cout, 3, 3, 3, seed=13) * 0.05
    want = F.conv_transpose3d(x.double(), wgt.double(), None, stride=2, padding=1, output_padding=1)
    y, sums = ops.conv3d_k3(x.to(DEV), ops.pack_conv3d_weight(wgt.to(DEV), transposed=True), transposed=True,
                            want_stats=True)
    assert y.shape == want.shape
    assert _rel_l2(y, want) < 1e-5
    torch.testing.assert_close(sums.cpu()[..., 0], want.sum((2, 3, 4)), rtol=1e-5, atol=1e-3)


CONV2D_CASES = [  # B, Cin, Cout, H, W, k, stride, dilation
    (2, 3, 32, 20, 40, 3, 1, 1), (1, 32, 32, 17, 70, 3, 1, 1), (1, 32, 32, 21, 50, 3, 2, 1), (1, 32, 64, 16, 64, 3, 2, 1),
    (1, 64, 64, 9, 33, 3, 1, 1), (1, 64, 128, 12, 36, 3, 1, 1), (2, 128, 128, 10, 44, 3, 1, 2), (1, 32, 64, 18, 30, 1, 2, 1),
    (1, 64, 128, 7, 12, 1, 1, 1), (2, 128, 32, 2, 3, 1, 1, 1), (1, 320, 128, 8, 36, 3, 1, 1), (1, 128, 32, 9, 60, 1, 1, 1),
]


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,stride,dil", CONV2D_CASES)
def test_conv2d_vs_fp64_oracle(B, Cin, Cout, H, W, k, stride, dil):
    from cmf_b200 import ops

    x = _rand(B, Cin, H, W, seed=50)
    wgt = _rand(Cout, Cin, k, k, seed=51) * (2.0 / (k * k * Cin)) ** 0.5
    pad = (k // 2) * dil
    want = F.conv2d(x.double(), wgt.double(), None, stride, pad, dil)
    y, sums = ops.conv2d(x.to(DEV), ops.pack_conv2d_weight(wgt.to(DEV)), k, stride, dil, want_stats=True)
    assert y.shape == want.shape
    assert _rel_l2(y, want) < 1e-5
    torch.testing.assert_close(sums.cpu()[..., 0], want.sum((2, 3)), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(sums.cpu()[..., 1], (want * want).sum((2, 3)), rtol=1e-5, atol=1e-3)


def test_feature_extractor_vs_oracle():
    """Own conv2d/GroupNorm kernels through the whole SPP feature extractor vs the CPU oracle (fp32 and fp64)."""
    from cmf.models import get_model

    torch.manual_seed(gc.WEIGHT_SEED)
    model = get_model("cmfsm").to(DEV).eval()
    left, _ = gc.seeded_pair(1, 256, 512)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    f32, a32 = orc.feature_extraction(sd, left)
    f64, a64 = orc.feature_extraction({k: v.double() for k, v in sd.items()}, left.double())
    with torch.no_grad():
        feat, full = model._features(left.to(DEV))
    assert torch.equal(torch.tensor(feat.shape), torch.tensor(f32.shape))
    print("feat rel-L2 ours/fp64 %.2e  ref32/fp64 %.2e ; full ours %.2e ref32 %.2e"
          % (_rel_l2(feat, f64), _rel_l2(f32, f64), _rel_l2(full, a64), _rel_l2(a32, a64)))
    assert _rel_l2(full, a64) < 1e-5
    assert _rel_l2(feat, f64) < max(2e-5, 3 * _rel_l2(f32, f64))


@pytest.mark.parametrize("C,relu,res", [(32, True, False), (64, True, True), (64, False, True), (32, False, False)])
def test_groupnorm_apply_vs_oracle(C, relu, res):
    from cmf_b200 import ops

    x = _rand(2, C, 5, 6, 12, seed=20) * 3 + 0.5
    gamma, beta = _rand(C, seed=21), _rand(C, seed=22)
    r = _rand(2, C, 5, 6, 12, seed=23) if res else None
    want = F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5)
    if res:
        want = want + r.double()
    if relu:
        want = want.relu()
    xs = x.to(DEV)
    got = ops.gn_apply(xs, ops.gn_stats(xs), gamma.to(DEV), beta.to(DEV), r.to(DEV) if res else None, relu)
    torch.testing.assert_close(got.cpu().double(), want, rtol=1e-5, atol=1e-5)


def test_groupnorm_odd_spatial():
    from cmf_b200 import ops

    x = _rand(1, 32, 3, 5, 7, seed=24)
    ones, zeros = torch.ones(32), torch.zeros(32)
    xs = x.to(DEV)
    got = ops.gn_apply(xs, ops.gn_stats(xs), ones.to(DEV), zeros.to(DEV))
    torch.testing.assert_close(got.cpu(), F.group_norm(x, 32, ones, zeros, 1e-5), rtol=1e-5, atol=1e-5)


def _model_with(sub, weights):
    from cmf.models import get_model

    model = get_model("cmfsm")
    getattr(model, sub).load_state_dict(weights)
    return model.to(DEV).eval()


def test_hourglass_golden_from_reference_module(golden_dir):
    """Product hourglass (conv/deconv + GN + skip + ReLU kernels) vs the reference module's outputs."""
    from test_oracle_golden import _hourglass_sd

    g = _npz(golden_dir, "cmfsm_modules.npz")
    model = _model_with("dres2", gc.seeded_weights(_hourglass_sd(), gc.SEED_HG_W, gn_affine=True))
    x, presqu, postsqu = (t.to(DEV) for t in gc.hourglass_inputs())
    with torch.no_grad():
        a = model._hourglass(model.dres2, x, None, None, None)
        b = model._hourglass(model.dres2, x, presqu, postsqu, None)
    for got, key in zip(a + b, ("hg_out_a", "hg_pre_a", "hg_post_a", "hg_out_b", "hg_pre_b", "hg_post_b")):
        assert _rel_l2(got, g[key]) < 2e-5, key
        torch.testing.assert_close(got.cpu(), g[key], rtol=1e-3, atol=2e-4)


def test_head_and_classifier_golden_from_reference_modules(golden_dir):
    g = _npz(golden_dir, "cmfsm_modules.npz")
    from cmf.models import get_model

    model = get_model("cmfsm")
    w = gc.seeded_weights(model.dres0.state_dict(), gc.SEED_HEAD_W, gn_affine=True)
    model.dres0.load_state_dict(w)
    w = gc.seeded_weights(model.classif1.state_dict(), gc.SEED_CLS_W, gn_affine=True)
    model.classif1.load_state_dict(w)
    model = model.to(DEV).eval()
    with torch.no_grad():
        t = model._cg(model.dres0[0], gc.head_input().to(DEV), relu=True)
        head = model._cg(model.dres0[2], t, relu=True)
        cls = model._classify(model.classif1, gc.classif_input().to(DEV))
    assert _rel_l2(head, g["head_out"]) < 2e-5
    assert _rel_l2(cls, g["cls_out"].squeeze(1)) < 2e-5


# ------------------------------------------------------------------------------------------ K5 / K4
def test_k5_golden_from_reference_module(golden_dir):
    from cmf_b200 import ops

    g = _npz(golden_dir, "cmfsm_modules.npz")
    proto = {"conv0.weight": torch.empty(32, 66, 1, 1), "conv1.weight": torch.empty(16, 32, 1, 1),
             "conv2.weight": torch.empty(8, 16, 1, 1), "conv3.weight": torch.empty(1, 8, 1, 1)}
    w = gc.seeded_weights({"similarity1." + k: v for k, v in proto.items()}, gc.SEED_K5_W)
    lr, hr = gc.k5_inputs()
    got = ops.ctxmap_weights(lr.to(DEV), hr.to(DEV), *[w["similarity1.conv%d.weight" % i].to(DEV) for i in range(4)])
    # first layer is evaluated as lr-part + hr-part + code-part: summation order differs from the
    # reference's single 66-wide dot product -> tolerance 1e-5 absolute on softmax weights
    torch.testing.assert_close(got.cpu(), g["k5_weights"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,h,w", [(1, 3, 5), (2, 8, 16), (1, 9, 21)])
def test_k5_vs_oracle_ragged(B, h, w):
    from cmf_b200 import ops

    proto = {"mapping_matrix.similarity1.conv0.weight": torch.empty(32, 66, 1, 1),
             "mapping_matrix.similarity1.conv1.weight": torch.empty(16, 32, 1, 1),
             "mapping_matrix.similarity1.conv2.weight": torch.empty(8, 16, 1, 1),
             "mapping_matrix.similarity1.conv3.weight": torch.empty(1, 8, 1, 1)}
    sd = gc.seeded_weights(proto, 77)
    lr, hr = _rand(B, 32, h, w, seed=30), _rand(B, 32, 4 * h, 4 * w, seed=31)
    want = orc.context_mapping_weights(sd, lr, hr)
    got = ops.ctxmap_weights(lr.to(DEV), hr.to(DEV), *[sd["mapping_matrix.similarity1.conv%d.weight" % i].to(DEV)
                                                        for i in range(4)])
    torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5)


def test_k4_golden_interior_from_reference(golden_dir):
    from cmf_b200 import ops

    g = _npz(golden_dir, "cmfsm_c1_full.npz")
    args = [g[k][None].contiguous().to(DEV) for k in ("k4_c1", "k4_c2", "k4_c3", "k4_w")]
    outs = ops.softargmin_ctxmap(*args, 4)
    for i, o in enumerate(outs, 1):
        torch.testing.assert_close(o[0, 0, 4:-4, 4:-4].cpu(), g["k4_out%d_interior" % i], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("B,D,h,w", [(1, 48, 8, 16), (2, 48, 11, 19), (1, 24, 3, 5), (1, 96, 17, 33)])
def test_k4_vs_oracle(B, D, h, w):
    from cmf_b200 import ops

    c = [_rand(B, D, h, w, seed=40 + i) * 4 for i in range(3)]
    wts = torch.softmax(_rand(B, 9, 4 * h, 4 * w, seed=44) * 2, 1)
    want = orc.softargmin_ctxmap(*c, wts, 4)
    got, low = ops.softargmin_ctxmap(*[t.to(DEV) for t in c], wts.to(DEV), 4, want_lowres=True)
    for a, b in zip(got, want):
        torch.testing.assert_close(a.cpu(), b, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(low[2].cpu(), orc.softargmin(c[0] + c[1] + c[2]), rtol=1e-5, atol=1e-4)
    # outputs are convex combinations of scale*p: bounded by scale*(D-1)
    assert float(got[2].max()) <= 4 * (D - 1) + 1e-3 and float(got[2].min()) >= -1e-3


# ------------------------------------------------------------------------------------------ bf16 / C8 / tcgen05
def _to_c8_ref(x):
    """torch reference of the C8 layout: [B,C,D,H,W] -> [B,C/8,D,H,W,8] bf16."""
    B, C, D, H, W = x.shape
    return x.view(B, C // 8, 8, D, H, W).permute(0, 1, 3, 4, 5, 2).contiguous().to(torch.bfloat16)


def test_c8_layout_converters_bit_exact():
    from cmf_b200 import ops

    x = _rand(2, 32, 3, 5, 12, seed=60)
    c8 = ops.f32_to_c8(x.to(DEV))
    assert torch.equal(c8.cpu(), _to_c8_ref(x))
    back = ops.c8_to_f32(c8)
    assert torch.equal(back.cpu(), x.to(torch.bfloat16).float())


def test_k1_c8_bf16_matches_fp32_cost_volume():
    from cmf_b200 import ops

    L, R = _rand(2, 32, 6, 20, seed=61), _rand(2, 32, 6, 20, seed=62)
    want = _to_c8_ref(orc.cost_volume_concat(L, R, 12))
    got = ops.cost_volume_concat_c8(L.to(DEV), R.to(DEV), 12)
    assert torch.equal(got.cpu(), want)


IGEMM_CASES = [  # B, Cin, Cout, D, H, W
    (1, 32, 32, 4, 16, 8), (1, 32, 32, 6, 20, 12), (2, 64, 32, 5, 16, 8), (1, 64, 64, 4, 18, 24), (1, 32, 32, 9, 37, 29),
    # depth-walking schedule: single plane, two planes, several work units per CTA / depth segments, sample changes
    (2, 32, 32, 1, 16, 8), (1, 32, 32, 2, 16, 16), (1, 64, 32, 12, 64, 80), (3, 32, 32, 16, 48, 200), (2, 64, 64, 6, 40, 72),
]


@pytest.mark.parametrize("B,Cin,Cout,D,H,W", IGEMM_CASES)
def test_conv3d_igemm_tcgen05_vs_oracle(B, Cin, Cout, D, H, W):
    """tcgen05/TMEM/TMA implicit GEMM vs fp64 conv of the SAME bf16-rounded operands; the only differences are
    fp32 accumulation order and the final bf16 rounding of the output (2^-9 relative per element)."""
    from cmf_b200 import ops

    x = _rand(B, Cin, D, H, W, seed=70)
    wgt = _rand(Cout, Cin, 3, 3, 3, seed=71) * (2.0 / (27 * Cin)) ** 0.5
    xq, wq = x.to(torch.bfloat16).double(), wgt.to(torch.bfloat16).double()
    want = F.conv3d(xq, wq, None, 1, 1)
    y, sums = ops.conv3d_igemm(ops.f32_to_c8(x.to(DEV)), ops.pack_igemm_weight(wgt.to(DEV)))
    torch.cuda.synchronize()
    got = ops.c8_to_f32(y).cpu().double()
    err = _rel_l2(got, want)
    print("igemm %s rel-L2 %.3e  max|d| %.3e" % ((B, Cin, Cout, D, H, W), err, float((got - want).abs().max())))
    assert err < 3e-3
    torch.testing.assert_close(sums.cpu()[..., 0], got.sum((2, 3, 4)), rtol=1e-6, atol=1e-3)
    torch.testing.assert_close(sums.cpu()[..., 1], (got * got).sum((2, 3, 4)), rtol=1e-6, atol=1e-3)


def test_gn_apply_c8_vs_oracle():
    from cmf_b200 import ops

    x = _rand(2, 64, 4, 6, 10, seed=80) * 2 + 0.3
    r = _rand(2, 64, 4, 6, 10, seed=81)
    gamma, beta = _rand(64, seed=82), _rand(64, seed=83)
    xq, rq = x.to(torch.bfloat16).float(), r.to(torch.bfloat16).float()
    want = (F.group_norm(xq.double(), 32, gamma.double(), beta.double(), 1e-5) + rq.double()).relu()
    xs = xq.to(DEV)
    got, split = ops.gn_apply_c8(ops.f32_to_c8(xs), ops.gn_stats(xs), gamma.to(DEV), beta.to(DEV),
                                 ops.f32_to_c8(rq.to(DEV)), True, want_split=True)
    assert _rel_l2(ops.c8_to_f32(got), want) < 3e-3
    assert torch.equal(split, ops.c8_parity_split(got))  # fused parity-split copy == the stand-alone kernel
    B, NC, D, H, W, _ = got.shape  # and == the definition: [B][(d&1)*4+(h&1)*2+(w&1)][C/8][D/2][H/2][W/2][8]
    ref = got.view(B, NC, D // 2, 2, H // 2, 2, W // 2, 2, 8).permute(0, 3, 5, 7, 1, 2, 4, 6, 8).reshape(B, 8, NC, D // 2, H // 2, W // 2, 8)
    assert torch.equal(split, ref)


@pytest.mark.parametrize("B,Cout,D,H,W", [(1, 64, 2, 16, 8), (2, 32, 3, 10, 12), (1, 64, 3, 18, 20), (1, 32, 5, 33, 9),
                                           (3, 32, 6, 40, 60), (2, 64, 4, 36, 60)])  # several tiles per persistent CTA
def test_deconv3d_igemm_tcgen05_vs_oracle(B, Cout, D, H, W):
    from cmf_b200 import ops

    x = _rand(B, 64, D, H, W, seed=90)
    wgt = _rand(64, Cout, 3, 3, 3, seed=91) * 0.05
    xq, wq = x.to(torch.bfloat16).double(), wgt.to(torch.bfloat16).double()
    want = F.conv_transpose3d(xq, wq, None, stride=2, padding=1, output_padding=1)
    y, sums = ops.deconv3d_igemm(ops.f32_to_c8(x.to(DEV)), ops.pack_igemm_weight(wgt.to(DEV), transposed=True))
    torch.cuda.synchronize()
    got = ops.c8_to_f32(y).cpu().double()
    assert got.shape == want.shape
    err = _rel_l2(got, want)
    print("deconv igemm %s rel-L2 %.3e" % ((B, Cout, D, H, W), err))
    assert err < 3e-3
    torch.testing.assert_close(sums.cpu()[..., 0], got.sum((2, 3, 4)), rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("B,Cin,D,H,W", [(1, 32, 4, 32, 16), (2, 64, 4, 20, 24), (1, 32, 6, 36, 44), (1, 64, 10, 66, 18),
                                          (3, 32, 8, 48, 80), (2, 64, 6, 40, 56)])
def test_conv3d_s2_igemm_tcgen05_vs_oracle(B, Cin, D, H, W):
    from cmf_b200 import ops

    x = _rand(B, Cin, D, H, W, seed=92)
    wgt = _rand(64, Cin, 3, 3, 3, seed=93) * (2.0 / (27 * Cin)) ** 0.5
    xq, wq = x.to(torch.bfloat16).double(), wgt.to(torch.bfloat16).double()
    want = F.conv3d(xq, wq, None, 2, 1)
    xs = ops.c8_parity_split(ops.f32_to_c8(x.to(DEV)))
    y, sums = ops.conv3d_s2_igemm(xs, ops.pack_igemm_weight(wgt.to(DEV)))
    torch.cuda.synchronize()
    got = ops.c8_to_f32(y).cpu().double()
    assert got.shape == want.shape
    err = _rel_l2(got, want)
    print("s2 igemm %s rel-L2 %.3e" % ((B, Cin, D, H, W), err))
    assert err < 3e-3
    torch.testing.assert_close(sums.cpu()[..., 1], (got * got).sum((2, 3, 4)), rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("B,D,H,W", [(1, 4, 6, 10), (2, 5, 9, 33), (1, 12, 48, 40), (1, 1, 16, 8), (3, 16, 48, 200)])
def test_conv3d_igemm_cout1_vs_oracle(B, D, H, W):
    """classifier tail on tensor cores: both operands bf16, fp32 accumulation, fp32 output (no output rounding)."""
    from cmf_b200 import ops

    x = _rand(B, 32, D, H, W, seed=95)
    wgt = _rand(1, 32, 3, 3, 3, seed=96) * 0.05
    want = F.conv3d(x.to(torch.bfloat16).double(), wgt.to(torch.bfloat16).double(), None, 1, 1).squeeze(1)
    padded = torch.zeros(32, 32, 3, 3, 3)
    padded[:1] = wgt
    got = ops.conv3d_igemm_cout1(ops.f32_to_c8(x.to(DEV)), ops.pack_igemm_weight(padded.to(DEV)))
    assert got.shape == want.shape
    assert _rel_l2(got, want) < 1e-5


@pytest.mark.parametrize("B,H,W", [(1, 64, 128), (2, 144, 240), (1, 96, 312)])
def test_spp_pool_and_upsample_concat_vs_oracle(B, H, W):
    """SPP pools (floor mode) and the bilinear upsample + concat vs the ATen ops the reference calls."""
    from cmf_b200 import ops

    skip = _rand(B, 128, H, W, seed=97)
    raw = _rand(B, 64, H, W, seed=98)
    got = ops.spp_pool(skip.to(DEV))
    for k, g in zip((64, 32, 16, 8), got):
        torch.testing.assert_close(g.cpu(), F.avg_pool2d(skip, (k, k), (k, k)), rtol=1e-5, atol=1e-6)
    bs = [_rand(B, 32, H // k, W // k, seed=100 + k) for k in (8, 16, 32, 64)]  # b4, b3, b2, b1
    cat = ops.spp_upsample_concat(raw.to(DEV), skip.to(DEV), *[t.to(DEV) for t in bs])
    want = torch.cat([raw, skip] + [F.interpolate(t, (H, W), mode="bilinear", align_corners=False) for t in bs], 1)
    torch.testing.assert_close(cat.cpu(), want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,h,w", [(1, 2, 8), (2, 5, 11), (1, 16, 24)])
def test_k5_backward_vs_autograd(B, h, w):
    """cmfb200_ctxmap_weights_bwd vs autograd through the PyTorch closed form of the same function (fp64 on CPU)."""
    from cmf_b200 import autograd_ops as aops
    from cmf_b200 import ops

    lr, hr = _rand(B, 32, h, w, seed=120), _rand(B, 32, 4 * h, 4 * w, seed=121)
    ws = [_rand(32, 66, 1, 1, seed=122) * 0.3, _rand(16, 32, 1, 1, seed=123) * 0.4, _rand(8, 16, 1, 1, seed=124) * 0.5,
          _rand(1, 8, 1, 1, seed=125)]
    g = _rand(B, 9, 4 * h, 4 * w, seed=126)
    ref_in = [t.double().requires_grad_(True) for t in [lr, hr] + ws]
    ref_out = aops._ctxmap_weights_torch(*ref_in)
    ref_grads = torch.autograd.grad(ref_out, ref_in, g.double())
    dev = [t.to(DEV) for t in [lr, hr] + ws]
    out = ops.ctxmap_weights(*dev)
    torch.testing.assert_close(out.cpu().double(), ref_out.detach(), rtol=1e-4, atol=1e-6)
    got = ops.ctxmap_weights_bwd(*dev, out, g.to(DEV))
    for name, a, b in zip(("d_lr", "d_hr", "d_w0", "d_w1", "d_w2", "d_w3"), got, ref_grads):
        assert a.shape == b.shape, name
        err = _rel_l2(a.cpu().double(), b)
        print("K5 bwd %s rel-L2 %.2e" % (name, err))
        assert err < 2e-5, (name, err)


@pytest.mark.parametrize("shape,relu,res", [((2, 64, 3, 5, 7), True, True), ((1, 32, 4, 6, 8), False, False),
                                             ((3, 128, 9, 12), True, False), ((2, 32, 10, 6), False, True)])
def test_gn_backward_vs_autograd(shape, relu, res):
    """cmfb200_gn_bwd vs autograd through F.group_norm (+residual)(+ReLU) in fp64."""
    from cmf_b200 import ops

    C = shape[1]
    x, g = _rand(*shape, seed=130) * 2 + 0.3, _rand(*shape, seed=131)
    r = _rand(*shape, seed=132) if res else None
    gamma, beta = _rand(C, seed=133) + 1.5, _rand(C, seed=134)
    xr, gr, br = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rr = r.double().requires_grad_(True) if res else None
    y = F.group_norm(xr, 32, gr, br, 1e-5)
    if res:
        y = y + rr
    if relu:
        y = F.relu(y)
    want = torch.autograd.grad(y, [xr, gr, br] + ([rr] if res else []), g.double())
    xd = x.to(DEV)
    sums = ops.gn_stats(xd)
    out = ops.gn_apply(xd, sums, gamma.to(DEV), beta.to(DEV), r.to(DEV) if res else None, relu)
    dx, dg, db, dres = ops.gn_backward(g.to(DEV), xd, sums, gamma.to(DEV), out if relu else None, res)
    assert _rel_l2(dx.cpu(), want[0]) < 1e-5
    assert _rel_l2(dg.cpu(), want[1]) < 1e-5 and _rel_l2(db.cpu(), want[2]) < 1e-5
    if res:
        assert _rel_l2(dres.cpu(), want[3]) < 1e-6
    else:
        assert dres is None


@pytest.mark.parametrize("B,D,H,W", [(1, 3, 5, 7), (2, 6, 9, 20), (1, 4, 7, 4), (2, 5, 16, 64)])
def test_conv3d_cout1_backward_vs_autograd(B, D, H, W):
    from cmf_b200 import ops

    x, wgt, g = _rand(B, 32, D, H, W, seed=140), _rand(1, 32, 3, 3, 3, seed=141) * 0.1, _rand(B, 1, D, H, W, seed=142)
    xr, wr = x.double().requires_grad_(True), wgt.double().requires_grad_(True)
    want = torch.autograd.grad(F.conv3d(xr, wr, None, 1, 1), [xr, wr], g.double())
    dx, dw = ops.conv3d_cout1_backward(x.to(DEV), wgt.to(DEV), g.to(DEV))
    assert _rel_l2(dx.cpu(), want[0]) < 1e-5 and _rel_l2(dw.cpu(), want[1]) < 1e-5


# ---- cmfsm_sub_8 variant kernels ------------------------------------------------------------------------
def _sub8_sd(seed):
    proto = {"mapping_matrix.similarity1.conv0.weight": torch.empty(32, 66, 1, 1),
             "mapping_matrix.similarity1.conv1.weight": torch.empty(16, 32, 1, 1),
             "mapping_matrix.similarity1.conv2.weight": torch.empty(8, 16, 1, 1),
             "mapping_matrix.similarity1.conv3.weight": torch.empty(1, 8, 1, 1)}
    return gc.seeded_weights(proto, seed)


@pytest.mark.parametrize("B,h,w,scale", [(1, 3, 5, 8), (2, 6, 10, 8), (1, 4, 9, 4)])
def test_k5_five_neighbour_variant_vs_oracle(B, h, w, scale):
    """cmfb200_ctxmap_weights5_fwd vs the restatement of six_related_context_mapping (pinned to the reference)."""
    import cmfsm_sub8_oracle as orc8
    from cmf_b200 import ops

    sd = _sub8_sd(150)
    lr, hr = _rand(B, 32, h, w, seed=151), _rand(B, 32, h * scale, w * scale, seed=152)
    want = orc8.context_mapping_weights5({k: v.double() for k, v in sd.items()}, lr.double(), hr.double())
    got = ops.ctxmap_weights5(lr.to(DEV), hr.to(DEV), *[sd["mapping_matrix.similarity1.conv%d.weight" % i].to(DEV) for i in range(4)])
    assert got.shape == want.shape
    torch.testing.assert_close(got.cpu().double(), want, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,D,h,w,scale", [(1, 6, 5, 7, 8), (2, 24, 9, 18, 8), (1, 12, 8, 16, 4)])
def test_k4_five_neighbour_variant_vs_oracle(B, D, h, w, scale):
    import cmfsm_sub8_oracle as orc8
    from cmf_b200 import ops

    cs = [_rand(B, D, h, w, seed=160 + i) * 3 for i in range(3)]
    w5 = _rand(B, 5, h * scale, w * scale, seed=164)
    want = orc8.softargmin_ctxmap5(*[c.double() for c in cs], w5.double(), scale)
    got = ops.softargmin_ctxmap5(*[c.to(DEV) for c in cs], w5.to(DEV), scale)
    for a, b in zip(got, want):
        torch.testing.assert_close(a.cpu().double(), b, rtol=1e-5, atol=1e-4)


def test_conv2d_dilation4_and_sized_spp_concat():
    from cmf_b200 import ops

    x = _rand(2, 128, 20, 36, seed=170)
    wgt = _rand(128, 128, 3, 3, seed=171) * (2.0 / (9 * 128)) ** 0.5
    want = F.conv2d(x.double(), wgt.double(), None, 1, 4, 4)
    y, sums = ops.conv2d(x.to(DEV), ops.pack_conv2d_weight(wgt.to(DEV)), 3, 1, 4, want_stats=True)
    assert _rel_l2(y, want) < 1e-5
    torch.testing.assert_close(sums.cpu()[..., 0], want.sum((2, 3)), rtol=1e-6, atol=1e-4)
    raw, skip = _rand(1, 64, 32, 64, seed=172), _rand(1, 128, 32, 64, seed=173)
    br = [_rand(1, 32, 32 // k, 64 // k, seed=174 + i) for i, k in enumerate((8, 16, 32, 4))]
    want = torch.cat([raw, skip] + [F.interpolate(b, (32, 64), mode="bilinear", align_corners=False) for b in br], 1)
    got = ops.spp_upsample_concat_sized(raw.to(DEV), skip.to(DEV), [b.to(DEV) for b in br])
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,Dl,h,w,scale", [(1, 3, 4, 6, 8), (2, 12, 5, 9, 16), (1, 4, 6, 20, 4)])
def test_volume_mapping_vs_oracle(B, Dl, h, w, scale):
    """cmfb200_volume_mapping_fwd (cmfsm_sub_16 epilogue) vs the restatement that materialises the volumes."""
    import cmfsm_sub16_oracle as orc16
    from cmf_b200 import ops

    cs = [_rand(B, Dl, h, w, seed=180 + i) * 2 for i in range(3)]
    w5 = _rand(B, 5, h * scale, w * scale, seed=184) * 0.5
    w3 = _rand(B, 3, h * scale, w * scale, seed=185) * 0.5 + 0.3
    want = orc16.volume_mapping(*[c.double() for c in cs], w5.double(), w3.double(), scale, Dl * scale)
    got = ops.volume_mapping(*[c.to(DEV) for c in cs], w5.to(DEV), w3.to(DEV), scale)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        torch.testing.assert_close(a.cpu().double(), b, rtol=1e-4, atol=1e-3)


def test_k5_target_three_neighbour_variant_vs_oracle():
    import cmfsm_sub16_oracle as orc16
    from cmf_b200 import ops

    sd = _sub8_sd(190)
    lr, hr = _rand(2, 32, 4, 7, seed=191), _rand(2, 32, 64, 112, seed=192)
    sd64 = {k: v.double() for k, v in sd.items()}
    w5, w3 = orc16.context_mapping_weights(sd64, lr.double(), hr.double(), lr.double(), hr.double())
    ws = [sd["mapping_matrix.similarity1.conv%d.weight" % i].to(DEV) for i in range(4)]
    torch.testing.assert_close(ops.ctxmap_weights3(lr.to(DEV), hr.to(DEV), *ws).cpu().double(), w3, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ops.ctxmap_weights5(lr.to(DEV), hr.to(DEV), *ws).cpu().double(), w5, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,D,h,w,scale", [(1, 6, 3, 5, 4), (2, 12, 9, 20, 4), (1, 8, 17, 33, 4)])
def test_k4_backward_vs_autograd(B, D, h, w, scale):
    """cmfb200_softargmin_ctxmap_bwd vs autograd through the PyTorch closed form of K4 in fp64."""
    from cmf_b200 import autograd_ops as aops
    from cmf_b200 import ops

    cs = [_rand(B, D, h, w, seed=200 + i) * 2 for i in range(3)]
    w9 = torch.softmax(_rand(B, 9, h * scale, w * scale, seed=204), 1)
    gs = [_rand(B, 1, h * scale, w * scale, seed=205 + i) for i in range(3)]
    ins = [t.double().requires_grad_(True) for t in cs + [w9]]
    outs = aops._softargmin_ctxmap_torch(*ins, scale)
    want = torch.autograd.grad(outs, ins, [g.double() for g in gs])
    got = ops.softargmin_ctxmap_bwd(*[c.to(DEV) for c in cs], w9.to(DEV), *[g.to(DEV) for g in gs], scale)
    for name, a, b in zip(("dc1", "dc2", "dc3", "dw9"), got, want):
        err = _rel_l2(a.cpu().double(), b)
        print("K4 bwd %s rel-L2 %.2e" % (name, err))
        assert a.shape == b.shape and err < 2e-5, (name, err)


@pytest.mark.parametrize("B,Dl,h,w,maxdisp,H,W", [(1, 6, 5, 7, 24, 20, 28), (2, 12, 9, 18, 192, 144, 288), (1, 5, 4, 9, 40, 32, 72)])
def test_trilinear_softargmin_vs_oracle(B, Dl, h, w, maxdisp, H, W):
    """cmfb200_trilinear_softargmin_fwd vs F.interpolate(trilinear) + softmax + regression in fp64."""
    import bilinear_oracle as orcb
    from cmf_b200 import ops

    cs = [_rand(B, Dl, h, w, seed=210 + i) * 2 for i in range(3)]
    want = orcb.trilinear_softargmin(*[c.double() for c in cs], maxdisp, H, W)
    got = ops.trilinear_softargmin(*[c.to(DEV) for c in cs], maxdisp, H, W)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        torch.testing.assert_close(a.cpu().double(), b, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("B,C,h,w,D,norm", [(1, 32, 6, 40, 12, False), (2, 32, 5, 77, 24, True), (1, 64, 3, 33, 48, False),
                                             (1, 8, 4, 20, 30, True)])
def test_cost_volume_corr_vs_oracle(B, C, h, w, D, norm):
    """Correlation form of K1 (warp-shuffle channel reduction) vs the oracle restatement (parity unpinned by the
    reference: it has no live correlation path, see oracle/cmfsm_oracle.py::cost_volume_corr)."""
    from cmf_b200 import ops

    L, R = _rand(B, C, h, w, seed=110), _rand(B, C, h, w, seed=111)
    want = orc.cost_volume_corr(L.double(), R.double(), D, norm)
    got = ops.cost_volume_corr(L.to(DEV), R.to(DEV), D, norm)
    assert got.shape == want.shape
    torch.testing.assert_close(got.cpu().double(), want, rtol=1e-5, atol=1e-6)
    assert not torch.signbit(got.cpu()[want == 0]).any()  # +0.0 in the masked triangle, like the concat volume


@pytest.mark.parametrize("B,H,W", [(1, 16, 24), (3, 37, 53)])
def test_masked_smooth_l1_fused_vs_reference_statement(B, H, W):
    """The two fused loss kernels vs the reference drivers' statement (train.py:162-174), value and gradients."""
    from cmf_b200 import autograd_ops as aops

    g = torch.Generator().manual_seed(120)
    disp = torch.rand(B, H, W, generator=g) * 260 - 30  # some targets outside (0, 192)
    outs = [(disp + torch.randn(B, H, W, generator=g) * s).unsqueeze(1) for s in (0.3, 2.0, 20.0)]
    ref_in = [o.clone().double().requires_grad_(True) for o in outs]
    want = orc.masked_smooth_l1(ref_in, disp.double(), 192)
    want.backward()
    ours_in = [o.clone().to(DEV).requires_grad_(True) for o in outs]
    got = aops.masked_smooth_l1(ours_in, disp.to(DEV), 192)
    (got * 3.0).backward()
    assert abs(float(got.detach()) - float(want.detach())) < 1e-5 * abs(float(want.detach()))
    for a, b in zip(ours_in, ref_in):
        torch.testing.assert_close(a.grad.cpu().double(), 3.0 * b.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_copy_rows_pitched(dtype):
    """cmfb200_copy_2d (halo rows of band activations): strided rows <- dense, dense <- strided rows, odd layouts fall
    back to copy_; bit-exact, nothing outside the rows is touched."""
    from cmf_b200 import lib, ops

    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 4, 3, 6, 10, 12, 8, generator=g).to(DEV).to(dtype)
    src = torch.randn(1, 4, 3, 6, 2, 12, 8, generator=g).to(DEV).to(dtype)
    want = x.clone()
    want.narrow(4, 8, 2).copy_(src)
    n0 = lib.launch_count()
    ops.copy_rows(x.narrow(4, 8, 2), src)
    assert lib.launch_count() == n0 + 1 and torch.equal(x, want)
    dense = torch.empty_like(src)
    ops.copy_rows(dense, x.narrow(4, 0, 2))
    assert lib.launch_count() == n0 + 2 and torch.equal(dense, x.narrow(4, 0, 2))
    # 4-byte-aligned only (one fp32 column): ATen path, same result
    y = torch.randn(2, 5, 7, generator=g).to(DEV)
    col = torch.randn(2, 5, 1, generator=g).to(DEV)
    ops.copy_rows(y.narrow(2, 3, 1), col)
    assert lib.launch_count() == n0 + 2 and torch.equal(y[:, :, 3:4], col)
