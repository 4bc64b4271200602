"""GPU, needs >= 2 devices (skipped otherwise): the two N>1 modes on real hardware through torchrun + NCCL.
The host-side logic of both is covered on CPU by tests/test_parallel_cpu.py (gloo, world_size 2)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "tools", script)] + list(args)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_band_sharding_matches_single_gpu():
    rep = _torchrun("check_row_bands.py", "--height", "512", "--width", "512", "--steps", "1")
    # the same arithmetic per voxel, only the GroupNorm summation order can differ: measured 0.0 on the B200
    assert rep["world"] == 2 and rep["pred3_max_abs"] < 1e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_band_forward_as_one_cuda_graph_per_rank():
    """Peer-memory halo transport + CUDA-graph replay of the whole band forward: still bit-identical to one GPU."""
    rep = _torchrun("check_row_bands.py", "--height", "512", "--width", "768", "--steps", "2", "--graph")
    assert rep["world"] == 2 and rep["pred1_max_abs"] == 0.0 and rep["pred3_max_abs"] == 0.0, rep


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_gradients_match_single_process():
    rep = _torchrun("check_dp_train.py", "--per-rank-batch", "1", "--steps", "1")
    assert rep["world"] == 2 and rep["grad_rel_l2"] < 2e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_reference_style_dataparallel_wrapper():
    """The reference drivers wrap the model in nn.DataParallel (test.py:40-44, one sample per replica thread):
    the drop-in module and libcmfb200 must work re-entrantly on two devices from two Python threads."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import golden_common as gc
    from cmf.models import get_model

    torch.manual_seed(gc.WEIGHT_SEED)
    model = get_model("cmfsm").cuda(0).eval()
    left, right = gc.seeded_pair(2, 256, 512, seed=21)
    left, right = left.cuda(0), right.cuda(0)
    with torch.no_grad():
        single = [model(left[i:i + 1].contiguous(), right[i:i + 1].contiguous()) for i in range(2)]
        dp = torch.nn.DataParallel(model, device_ids=[0, 1])
        out = dp(left, right)
    assert out[2].shape == (2, 1, 256, 512) and out[2].device.index == 0
    for i in range(2):
        for a, b in zip(out, single[i]):
            torch.testing.assert_close(a[i:i + 1], b, rtol=0, atol=5e-3)
