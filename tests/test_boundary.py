"""CPU: the drop-in boundary -- registry, constructor, state_dict contract, seeded init, C-ABI exports,
and the no-fallback rule (CPU tensors / missing kernels must raise)."""
import ctypes
import json
import os
import re

import pytest
import torch

import golden_common as gc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_names():
    from cmf.models import get_model

    m = get_model("cmfsm")
    assert type(m).__name__ == "cmfsm" and m.maxdisp == 192
    m8 = get_model("cmfsm_sub_8")
    assert type(m8).__name__ == "cmfsm_sub_8" and m8.maxdisp == 192
    for name in ("cmfsm_sub_16", "cm_sub_4", "cm_sub_8", "cm_sub_16", "bilinear_cmf", "bilinear_cmf_sub_8",
                 "bilinear_cmf_sub_16"):
        assert type(get_model(name)).__name__ == name
    assert type(get_model("cmf")).__name__ == "cmf"
    with pytest.raises(KeyError):
        get_model("no_such_model")


def test_state_dict_contract_and_seeded_init(golden_dir):
    """Keys, order, shapes AND values of the seed-0 initialisation equal the reference's."""
    from cmf.models.cmfsm import cmfsm

    contract = json.load(open(os.path.join(golden_dir, "cmfsm_state_dict.json")))
    torch.manual_seed(contract["weight_seed"])
    sd = cmfsm(maxdisp=192).state_dict()
    assert len(sd) == contract["n_tensors"] == 272
    assert sum(v.numel() for v in sd.values()) == contract["n_params"] == 5255368
    for (k, v), t in zip(sd.items(), contract["tensors"]):
        assert k == t["key"] and list(v.shape) == t["shape"]
        assert abs(float(v.double().sum()) - t["sum"]) < 1e-9, k
        assert abs(float(v.double().abs().sum()) - t["abssum"]) < 1e-9, k


@pytest.mark.parametrize("name,fixture", [("cmfsm_sub_8", "cmfsm_sub8_state_dict.json"),
                                          ("cmfsm_sub_16", "cmfsm_sub16_state_dict.json"),
                                          ("cm_sub_4", "cm_sub4_state_dict.json"), ("cm_sub_8", "cm_sub8_state_dict.json"),
                                          ("cm_sub_16", "cm_sub16_state_dict.json"),
                                          ("bilinear_cmf", "bilinear4_state_dict.json"),
                                          ("bilinear_cmf_sub_8", "bilinear8_state_dict.json"),
                                          ("bilinear_cmf_sub_16", "bilinear16_state_dict.json"), ("cmf", "cmf_state_dict.json")])
def test_variant_state_dict_contract_and_seeded_init(golden_dir, name, fixture):
    """The 1/8- and 1/16-resolution variants: keys, order, shapes and seed-0 values equal the reference's."""
    from cmf.models import get_model

    contract = json.load(open(os.path.join(golden_dir, fixture)))
    torch.manual_seed(contract["weight_seed"])
    sd = get_model(name).state_dict()
    assert len(sd) == contract["n_tensors"]
    assert sum(v.numel() for v in sd.values()) == contract["n_params"]
    for (k, v), t in zip(sd.items(), contract["tensors"]):
        assert k == t["key"] and list(v.shape) == t["shape"]
        assert abs(float(v.double().sum()) - t["sum"]) < 1e-9, k
        assert abs(float(v.double().abs().sum()) - t["abssum"]) < 1e-9, k


def test_checkpoint_round_trip_with_module_prefix(tmp_path):
    """train.py:228-234 saves {'epoch','model_state','optimizer_state'} from a DataParallel wrapper."""
    from cmf.models import get_model

    model = get_model("cmfsm")
    state = {"module." + k: v for k, v in model.state_dict().items()}  # what a DataParallel wrapper would save
    path = tmp_path / "ckpt.pkl"
    torch.save({"epoch": 3, "model_state": state, "optimizer_state": {}}, path)
    ckpt = torch.load(path)
    fresh = get_model("cmfsm")
    wrapped = torch.nn.Module()
    wrapped.module = fresh
    missing, unexpected = wrapped.load_state_dict(ckpt["model_state"], strict=True)
    assert not missing and not unexpected
    # dict-filtered partial load (train_kitti.py:134-144)
    partial = {k: v for k, v in ckpt["model_state"].items() if "dres" in k}
    model_dict = wrapped.state_dict()
    model_dict.update(partial)
    wrapped.load_state_dict(model_dict)


def test_c_abi_exports_every_declared_symbol():
    from cmf_b200 import lib

    header = open(os.path.join(ROOT, "include", "cmfb200.h")).read()
    declared = set(re.findall(r"\b(cmfb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    handle = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    L = lib.load()
    assert L.cmfb200_abi_version() == lib.ABI_VERSION
    assert isinstance(lib.launch_count(), int)


def test_c_abi_argument_validation_without_gpu():
    """Shape/pointer validation happens before any CUDA call, so it is testable on a CPU box."""
    from cmf_b200 import lib

    L = lib.load()
    rc = L.cmfb200_cost_volume_concat_fwd(None, None, None, 1, 32, 8, 8, 4, None)
    assert rc == -1 and b"null pointer" in L.cmfb200_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = L.cmfb200_cost_volume_concat_fwd(p, p, p, 1, 32, 8, 10, 4, None)
    assert rc == -1 and b"multiple of 4" in L.cmfb200_last_error()
    rc = L.cmfb200_conv3d_k3_fwd(p, p, p, None, 1, 32, 48, 4, 4, 4, 1, None)
    assert rc == -1 and b"unsupported" in L.cmfb200_last_error()
    rc = L.cmfb200_ctxmap_weights_fwd(p, p, p, p, p, p, p, 1, 4, 4, 3, 0, 4, None)
    assert rc == -1 and b"odd scale" in L.cmfb200_last_error()


def test_no_cpu_fallback():
    from cmf.models import get_model
    from cmf_b200 import lib, ops

    model = get_model("cmfsm").eval()
    left, right = gc.seeded_pair(1, 256, 512)
    with torch.no_grad(), pytest.raises(lib.CmfB200Error):
        model(left, right)
    with pytest.raises(lib.CmfB200Error):
        ops.cost_volume_concat(torch.zeros(1, 32, 4, 8), torch.zeros(1, 32, 4, 8), 4)
    with pytest.raises(ValueError):
        model(left[:, :, :250], right[:, :, :250])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
