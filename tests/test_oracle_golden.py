"""CPU: the oracle restatement (oracle/cmfsm_oracle.py) against fixtures produced by the REAL reference
(oracle/gen_golden.py, run in the build container).  No reference tree needed."""
import json
import os

import numpy as np
import pytest
import torch

import cmfsm_oracle as orc
import golden_common as gc


def _npz(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def test_k1_cost_volume_crop_bit_exact(golden_dir):
    g = _npz(golden_dir, "cmfsm_c1_full.npz")
    L, R = g["k1_L_crop"].unsqueeze(0), g["k1_R_crop"].unsqueeze(0)  # [1,cs,rs,128]
    cost = orc.cost_volume_concat(L, R, 48)
    assert torch.equal(cost[0], g["k1_cost_crop"])
    # +0.0 (never -0.0) in the masked triangle
    assert not torch.signbit(cost[0, :, 5, :, :5]).any()


def test_k1_adjoint_is_transpose():
    g = torch.Generator().manual_seed(5)
    L, R = torch.randn(2, 3, 4, 12, generator=g), torch.randn(2, 3, 4, 12, generator=g)
    G = torch.randn(2, 6, 7, 4, 12, generator=g)
    dL, dR = orc.cost_volume_concat_bwd(G, 3)
    lhs = (orc.cost_volume_concat(L, R, 7).double() * G.double()).sum()
    rhs = (L.double() * dL.double()).sum() + (R.double() * dR.double()).sum()
    assert abs(lhs - rhs) < 1e-4 * abs(lhs).clamp_min(1)


def test_k5_mapping_weights(golden_dir):
    g = _npz(golden_dir, "cmfsm_modules.npz")
    proto = {"similarity1.conv0.weight": torch.empty(32, 66, 1, 1), "similarity1.conv1.weight": torch.empty(16, 32, 1, 1),
             "similarity1.conv2.weight": torch.empty(8, 16, 1, 1), "similarity1.conv3.weight": torch.empty(1, 8, 1, 1)}
    w = gc.seeded_weights(proto, gc.SEED_K5_W)
    lr, hr = gc.k5_inputs()
    out = orc.context_mapping_weights({"mapping_matrix." + k: v for k, v in w.items()}, lr, hr)
    torch.testing.assert_close(out, g["k5_weights"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out.sum(1), torch.ones_like(out[:, 0]), rtol=0, atol=1e-5)
    # out-of-image neighbours carry ~e^-100 weight: left neighbour on the first `scale` columns
    assert float(out[:, 1, :, :4].max()) < 1e-30


def _hourglass_sd():
    shapes = {}
    for name, (ci, co, tr) in {"conv1.0": (32, 64, 0), "conv2": (64, 64, 0), "conv3.0": (64, 64, 0),
                               "conv4.0": (64, 64, 0), "conv5": (64, 64, 1), "conv6": (64, 32, 1)}.items():
        shapes[name + ".0.weight"] = torch.empty((ci, co, 3, 3, 3) if tr else (co, ci, 3, 3, 3))
        shapes[name + ".1.weight"] = torch.empty(co)
        shapes[name + ".1.bias"] = torch.empty(co)
    return shapes


def test_hourglass_against_reference_module(golden_dir):
    g = _npz(golden_dir, "cmfsm_modules.npz")
    w = gc.seeded_weights(_hourglass_sd(), gc.SEED_HG_W, gn_affine=True)
    sd = {"hg." + k: v for k, v in w.items()}
    x, presqu, postsqu = gc.hourglass_inputs()
    a = orc.hourglass(sd, "hg", x, None, None)
    b = orc.hourglass(sd, "hg", x, presqu, postsqu)
    for got, key in zip(a + b, ("hg_out_a", "hg_pre_a", "hg_post_a", "hg_out_b", "hg_pre_b", "hg_post_b")):
        torch.testing.assert_close(got, g[key], rtol=1e-4, atol=1e-4)


def test_k4_softargmin_ctxmap_interior(golden_dir):
    g = _npz(golden_dir, "cmfsm_c1_full.npz")
    outs = orc.softargmin_ctxmap(g["k4_c1"][None], g["k4_c2"][None], g["k4_c3"][None], g["k4_w"][None], 4)
    for i, o in enumerate(outs, 1):
        torch.testing.assert_close(o[0, 0, 4:-4, 4:-4], g["k4_out%d_interior" % i], rtol=1e-5, atol=1e-4)


def test_full_forward_c1_against_reference(golden_dir):
    """Oracle forward at BASELINE config 1 (256x512, B=1, maxdisp 192) vs the reference's outputs."""
    from cmf.models import get_model

    g = _npz(golden_dir, "cmfsm_c1_full.npz")
    meta = json.load(open(os.path.join(golden_dir, "cmfsm_c1_meta.json")))
    torch.manual_seed(gc.WEIGHT_SEED)
    sd = get_model("cmfsm").state_dict()  # seeded init is pinned by test_boundary.py
    left, right = gc.seeded_pair(1, 256, 512)
    stages = {}
    p1, p2, p3 = orc.forward(sd, left, right, 192, stages)
    # fp32 reproducibility of the reference itself is 2e-3 px max between thread counts
    for got, key in ((p1, "pred1_sub"), (p2, "pred2_sub"), (p3, "pred3_sub")):
        d = (got[0, 0, ::4, ::4] - g[key]).abs()
        assert float(d.max()) < 2e-2 and float(d.mean()) < 1e-3, (key, float(d.max()), float(d.mean()))
    assert abs(float(p3.double().mean()) - meta["stats"]["pred3"][0]) < 1e-3
    torch.testing.assert_close(stages["weights"][0, :, ::8, ::8], g["weights_sub"], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(stages["L"][0, ::4, ::4, ::4], g["L_sub"], rtol=1e-3, atol=1e-4)


def test_sub8_full_forward_against_reference(golden_dir):
    """cmfsm_sub_8 oracle at 256x512 vs the outputs of the real reference module (oracle/gen_golden_sub8.py asserted
    exact equality when the fixtures were written; thread counts move fp32 results a little).  The variant's
    un-normalised softmax*logit weights make random-init outputs large, so the gates are relative."""
    import cmfsm_sub8_oracle as orc8
    from cmf.models import get_model

    g = _npz(golden_dir, "cmfsm_sub8_c1.npz")
    torch.manual_seed(gc.WEIGHT_SEED)
    sd = get_model("cmfsm_sub_8").state_dict()
    left, right = gc.seeded_pair(1, 256, 512)
    stages = {}
    p1, p2, p3 = orc8.forward(sd, left, right, 192, stages)
    for got, key in ((p1, "pred1_sub"), (p2, "pred2_sub"), (p3, "pred3_sub")):
        want = g[key]
        rel = float((got[0, 0, ::4, ::4] - want).norm() / want.norm())
        assert rel < 1e-3, (key, rel)
    torch.testing.assert_close(stages["weights"][0, :, ::8, ::8], g["weights_sub"], rtol=2e-3, atol=1e-3)
    torch.testing.assert_close(stages["L"][0, ::4, ::2, ::2], g["L_sub"], rtol=1e-3, atol=1e-4)


def test_sub16_full_forward_against_reference(golden_dir):
    """cmfsm_sub_16 oracle at 256x512 vs the outputs of the real reference module (exact when the fixtures were made)."""
    import cmfsm_sub16_oracle as orc16
    from cmf.models import get_model

    g = _npz(golden_dir, "cmfsm_sub16_c1.npz")
    torch.manual_seed(gc.WEIGHT_SEED)
    sd = get_model("cmfsm_sub_16").state_dict()
    left, right = gc.seeded_pair(1, 256, 512)
    stages = {}
    p1, p2, p3 = orc16.forward(sd, left, right, 192, stages)
    for got, key in ((p1, "pred1_sub"), (p2, "pred2_sub"), (p3, "pred3_sub")):
        assert tuple(got.shape) == (1, 256, 512)
        d = (got[0, ::4, ::4] - g[key]).abs()
        assert float(d.max()) < 0.5 and float(d.mean()) < 2e-2, (key, float(d.max()), float(d.mean()))
    torch.testing.assert_close(stages["w3"][0, :, ::8, ::8], g["w3_sub"], rtol=2e-3, atol=1e-3)
    torch.testing.assert_close(stages["c1"][0], g["c1_sub"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("variant", ["4", "8", "16"])
def test_cm_sub_forward_against_reference(golden_dir, variant):
    """Single-hourglass ablations cm_sub_8 / cm_sub_16: oracle vs the outputs of the real reference modules."""
    import cm_sub_oracle as orcs
    from cmf.models import get_model

    g = _npz(golden_dir, "cm_sub%s_c1.npz" % variant)
    torch.manual_seed(gc.WEIGHT_SEED)
    sd = get_model("cm_sub_" + variant).state_dict()
    left, right = gc.seeded_pair(1, 256, 512)
    stages = {}
    p1, p2, p3 = orcs.forward(sd, left, right, variant, 192, stages)
    assert p1 is p2 and p2 is p3 and tuple(p1.shape) == (1, 256, 512)
    d = (p1[0, ::4, ::4] - g["pred1_sub"]).abs()
    assert float(d.max()) < 0.5 and float(d.mean()) < 2e-2, (float(d.max()), float(d.mean()))
    torch.testing.assert_close(stages["c1"][0, ::2, ::2, ::2], g["c1_sub"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("variant,name", [("4", "bilinear_cmf"), ("8", "bilinear_cmf_sub_8"), ("16", "bilinear_cmf_sub_16")])
def test_bilinear_forward_against_reference(golden_dir, variant, name):
    """No-mapping baselines: oracle vs the outputs of the real reference modules."""
    import bilinear_oracle as orcb
    from cmf.models import get_model

    g = _npz(golden_dir, "bilinear%s_c1.npz" % variant)
    torch.manual_seed(gc.WEIGHT_SEED)
    sd = get_model(name).state_dict()
    left, right = gc.seeded_pair(1, 256, 512)
    for i, p in enumerate(orcb.forward(sd, left, right, variant, 192), 1):
        assert tuple(p.shape) == (1, 256, 512)
        d = (p[0, ::4, ::4] - g["pred%d_sub" % i]).abs()
        assert float(d.max()) < 0.5 and float(d.mean()) < 2e-2, (i, float(d.max()), float(d.mean()))


def test_cmf_forward_against_reference(golden_dir):
    """Registry model `cmf` (stride-2 stem + super-resolution refinement head): oracle vs the real reference outputs."""
    import cmf_oracle as orcc
    from cmf.models import get_model

    g = _npz(golden_dir, "cmf_c1.npz")
    torch.manual_seed(gc.WEIGHT_SEED)
    sd = get_model("cmf").state_dict()
    left, right = gc.seeded_pair(1, 256, 512)
    for i, p in enumerate(orcc.forward(sd, left, right, 192), 1):
        assert tuple(p.shape) == (1, 1, 256, 512)
        want = g["pred%d_sub" % i]
        rel = float((p[0, 0, ::4, ::4] - want).norm() / want.norm().clamp_min(1e-6))
        assert rel < 1e-3, (i, rel)


def test_shape_validation():
    import pytest

    with pytest.raises(ValueError):
        orc.check_shapes(540, 960, 192)
    with pytest.raises(ValueError):
        orc.check_shapes(256, 512, 100)
    with pytest.raises(ValueError):
        orc.check_shapes(256, 256, 192)
    orc.check_shapes(576, 960, 192)


def test_oracle_at_config4_shape_vs_reference_golden(golden_dir):
    """The oracle at the KITTI-padded 384x1248 shape (BASELINE config 4) against the UNMODIFIED reference's outputs
    (oracle/gen_golden_configs.py).  Thread count differs from the generating run, so the gate is the reference's
    own thread-count reproducibility (2.1e-3 px), not bit equality."""
    from cmf.models.cmfsm import cmfsm  # parameter container (seeded init == the reference's, test_boundary.py)

    g = _npz(golden_dir, "cmfsm_configs.npz")
    meta = json.load(open(os.path.join(golden_dir, "cmfsm_configs_meta.json")))
    assert all(c["oracle_equals_reference"] for c in meta["cases"].values())
    torch.manual_seed(gc.WEIGHT_SEED)
    sd = {k: v.detach() for k, v in cmfsm().state_dict().items()}
    left, right = gc.kitti_padded_pair(1)
    assert tuple(left.shape) == (1, 3, 384, 1248)
    outs = orc.forward(sd, left, right, 192)
    for i, o in enumerate(outs, 1):
        d = (o[0, 0, ::meta["sub"], ::meta["sub"]] - g["c4_s1_pred%d_ref32" % i]).abs()
        assert float(d.max()) < 5e-3, (i, float(d.max()))
