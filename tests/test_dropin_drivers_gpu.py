"""GPU: the UNMODIFIED reference drivers run against this package as a drop-in for `cmf.models` (north star: "drops
into train.py/test.py/eval_kitti.py unchanged"; SURVEY.md section 4 smoke test).

A scratch tree is assembled the way INTEGRATION.md describes the swap: the reference's own `cmf/loader`, `cmf/loss.py`,
`cmf/augmentations.py`, `train.py` (copies staged under the git-ignored oracle/_ref/reference by
`__graft_entry__.build()`), OUR `cmf/models` + `cmf_b200` in place of the reference's `cmf/models`, a `visdom` stub
(the package is not installed), synthetic `[H,W,7]` float `.npy` samples in the reference's on-disk format and a
`--resume` checkpoint with DataParallel `module.` keys.  Then

  * `python train.py --arch cmfsm --dataset flying3d --batch_size 1 --n_epoch 1 --resume ckpt` runs train.py:148-234
    unmodified: two iterations (mask, forward through nn.DataParallel, the three smooth-L1 terms, backward, Adam), the
    visdom calls of :186-223, and the checkpoint of :228-234;
  * the evaluation loop of test.py:63-98 (same loader, masks, DataParallel wrapper, `load_state_dict` of the checkpoint
    train.py just wrote) runs in-process -- test.py itself hard-codes `device_ids=[0,1]` and `/home/lidong/...` paths.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "reference")

VISDOM_STUB = '''"""Stand-in for the `visdom` package (not installed): accepts every call train.py makes."""


class Visdom:
    def __init__(self, *a, **k):
        self.calls = 0

    def _win(self, *a, **k):
        self.calls += 1
        return "win%d" % self.calls

    line = image = images = text = _win
'''


def _sample(h, w, seed):
    """One sample in the reference's on-disk format: float [H,W,7] = left RGB, right RGB (0..255), disparity."""
    g = np.random.default_rng(seed)
    base = g.random((h, w + 64, 3)) * 255.0
    disp = np.full((h, w, 1), 20.0)
    disp[: h // 8] = 0.0  # some invalid (masked) pixels
    return np.concatenate([base[:, 64:], base[:, 44:w + 44], disp], 2)


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    if not os.path.isdir(os.path.join(REF, "cmf", "loader")):
        pytest.skip("reference drivers are not staged (run __graft_entry__.build() where /root/reference exists)")
    work = str(tmp_path_factory.mktemp("dropin"))
    os.makedirs(os.path.join(work, "cmf"))
    shutil.copytree(os.path.join(REF, "cmf", "loader"), os.path.join(work, "cmf", "loader"))
    for name in ("__init__.py", "loss.py", "augmentations.py"):
        shutil.copy(os.path.join(REF, "cmf", name), os.path.join(work, "cmf", name))
    shutil.copy(os.path.join(REF, "train.py"), os.path.join(work, "train.py"))
    os.symlink(os.path.join(PKG, "cmf", "models"), os.path.join(work, "cmf", "models"))  # the drop-in
    os.symlink(os.path.join(PKG, "cmf_b200"), os.path.join(work, "cmf_b200"))
    with open(os.path.join(work, "visdom.py"), "w") as f:
        f.write(VISDOM_STUB)
    data = os.path.join(work, "data")
    for split, (h, w) in (("train", (300, 600)), ("test", (540, 960))):
        os.makedirs(os.path.join(data, split))
        for i in range(2):
            np.save(os.path.join(data, split, "%04d.npy" % i), _sample(h, w, 10 * i + len(split)))
    with open(os.path.join(work, "config.json"), "w") as f:
        json.dump({"flying3d": {"data_path": data + "/"}}, f)
    # --resume checkpoint in the reference's format (train.py:228-234): DataParallel 'module.' keys
    sys.path.insert(0, PKG)
    from cmf.models import get_model

    torch.manual_seed(0)
    sd = get_model("cmfsm").state_dict()
    torch.save({"epoch": 0, "model_state": {"module." + k: v for k, v in sd.items()}, "optimizer_state": {}},
               os.path.join(work, "ckpt.pkl"))
    return work


def test_reference_train_py_runs_unmodified(tree):
    with open(os.path.join(REF, "train.py"), "rb") as a, open(os.path.join(tree, "train.py"), "rb") as b:
        assert a.read() == b.read()  # the driver is byte-identical to the staged reference copy
    env = dict(os.environ, PYTHONPATH=tree, CUDA_VISIBLE_DEVICES="0", PYTHONDONTWRITEBYTECODE="1")
    run = subprocess.run([sys.executable, "train.py", "--arch", "cmfsm", "--dataset", "flying3d", "--batch_size", "1",
                          "--n_epoch", "1", "--resume", "ckpt.pkl"], cwd=tree, env=env, capture_output=True, text=True,
                         timeout=900)
    print(run.stdout[-1500:])
    assert run.returncode == 0, run.stderr[-3000:]
    assert "Loaded checkpoint 'ckpt.pkl' (epoch 0)" in run.stdout
    lines = [ln for ln in run.stdout.splitlines() if ln.startswith("data [")]
    assert len(lines) == 2, run.stdout[-1500:]  # two iterations of train.py:154-225
    losses = [float(ln.split("Loss:")[1]) for ln in lines]
    assert all(np.isfinite(losses)) and all(v > 0 for v in losses)
    saved = torch.load(os.path.join(tree, "0_cmfsm_flying3d_best_model.pkl"), map_location="cpu")
    assert saved["epoch"] == 1 and len(saved["model_state"]) == 272
    assert all(k.startswith("module.") for k in saved["model_state"])
    start = torch.load(os.path.join(tree, "ckpt.pkl"), map_location="cpu")["model_state"]
    moved = sum(float((saved["model_state"][k] - start[k]).abs().sum()) > 0 for k in start)
    assert moved >= 250, moved  # Adam updated (nearly) every tensor: gradients reached the whole network
    assert len(np.load(os.path.join(tree, "loss.npy"))) == 3


def test_reference_eval_loop_with_dataparallel_and_checkpoint(tree):
    """test.py:28-98 with its own loader, masks and metric; only the hard-coded paths / device list differ."""
    sys.path.insert(0, tree)
    for m in [k for k in sys.modules if k == "cmf" or k.startswith("cmf.")]:
        del sys.modules[m]
    cwd = os.getcwd()
    os.chdir(tree)
    try:
        from cmf.loader import get_data_path, get_loader
        from cmf.models import get_model
        from torch.utils import data

        v_loader = get_loader("flying3d")(get_data_path("flying3d"), is_transform=True, split="test", img_size=(540, 960))
        valloader = data.DataLoader(v_loader, batch_size=1, num_workers=0, shuffle=False)
        model = torch.nn.DataParallel(get_model("cmfsm"), device_ids=[0])
        model.cuda(0)
        path = os.path.join(tree, "0_cmfsm_flying3d_best_model.pkl")
        checkpoint = torch.load(path if os.path.exists(path) else os.path.join(tree, "ckpt.pkl"))
        model.load_state_dict(checkpoint["model_state"])
        model.eval()
        errors = []
        for left, right, disparity, image in valloader:
            with torch.no_grad():
                left, right = left.cuda(0), right.cuda(0)
                disparity = disparity.cuda(0)[:, :540, :960]
                mask = (disparity < 192) & (disparity >= 0)
                output1, output2, output3 = model(left, right)
                assert tuple(output3.shape) == (1, 1, 576, 960)
                output1 = torch.squeeze(output3, 1)[:, :540, :960]
                errors.append(torch.mean(torch.abs(output1[mask] - disparity[mask])).item())
        assert len(errors) == 2 and all(np.isfinite(errors))
    finally:
        os.chdir(cwd)
        sys.path.remove(tree)
        for m in [k for k in sys.modules if k == "cmf" or k.startswith("cmf.")]:
            del sys.modules[m]
