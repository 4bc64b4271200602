"""Import the UNMODIFIED reference package `cmf` (CPU, or the GPU under cuDNN with TF32 off).

TEST INFRASTRUCTURE ONLY -- used in the build container to pin the oracle restatement
(`oracle/cmfsm_oracle.py`) and to generate `tests/golden/*` (see `oracle/gen_golden.py`), and by the
`--impl reference` arm of bench.py / the driver drop-in test.  The reference tree is looked up at
`$CMF_REFERENCE_ROOT`, then `oracle/_ref/reference` (an unmodified copy staged by
`__graft_entry__.build()`, git-ignored, shipped to the GPU box with the snapshot), then `/root/reference`.

Two shims are needed:
  1. `cmf.caffe_pb2` is stale protobuf codegen imported (and never used) at cmf/models/cmfsm.py:19.
  2. the model hard-codes `.cuda()` (cmf/models/cmfsm.py:98,104,117,427-428,671); on a CPU-only box
     `.cuda()` becomes a no-op.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")
REFERENCE_ROOT = os.environ.get("CMF_REFERENCE_ROOT") or (_STAGED if os.path.isdir(os.path.join(_STAGED, "cmf", "models"))
                                                           else "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cmf", "models"))


def import_reference(force_cpu=False):
    """Returns (get_model, ref_module) where ref_module is the python module cmf.models.cmfsm.
    `force_cpu`: neutralise the hard-coded `.cuda()` calls even when a GPU is present (CPU baseline on the GPU box)."""
    import torch

    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if "cmf" in sys.modules and not getattr(sys.modules["cmf"], "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("a different `cmf` package is already imported; run the harness in its own process")
    sys.path.insert(0, REFERENCE_ROOT)
    import cmf  # noqa: E402  (empty cmf/__init__.py)

    stub = types.ModuleType("cmf.caffe_pb2")
    sys.modules["cmf.caffe_pb2"] = stub
    cmf.caffe_pb2 = stub
    if force_cpu or not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    else:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    from cmf.models import get_model  # cmf/models/__init__.py:19

    return get_model, sys.modules["cmf.models.cmfsm"]
