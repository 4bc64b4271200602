"""Model registry with the reference's entry point (reference cmf/models/__init__.py:19-41).

`get_model(name)` instantiates the class with NO arguments, exactly like the reference.  `cmfsm` is the B200-native
hot path (inference, training, row bands, bf16 aggregation).  The other nine registered names run
inference on the same kernels: `cmfsm_sub_8` / `cmfsm_sub_16` (the 1/8- and 1/16-resolution "downsample configs"), the
single-hourglass ablations `cm_sub_4` / `cm_sub_8` / `cm_sub_16`, the no-mapping baselines `bilinear_cmf` /
`bilinear_cmf_sub_8` / `bilinear_cmf_sub_16`, and `cmf` (PSMNet-style extractor + super-resolution refinement head).
An unknown name raises `KeyError` (the reference only prints).
"""
from cmf.models.cmfsm import cmfsm
from cmf.models.cmf import cmf
from cmf.models.cmfsm_sub_8 import cmfsm_sub_8
from cmf.models.cmfsm_sub_16 import cmfsm_sub_16
from cmf.models.cm_sub_4 import cm_sub_4
from cmf.models.cm_sub_8 import cm_sub_8
from cmf.models.cm_sub_16 import cm_sub_16
from cmf.models.bilinear_cmf import bilinear_cmf
from cmf.models.bilinear_cmf_sub_8 import bilinear_cmf_sub_8
from cmf.models.bilinear_cmf_sub_16 import bilinear_cmf_sub_16

_IMPLEMENTED = {"cmf": cmf, "cmfsm": cmfsm, "cmfsm_sub_8": cmfsm_sub_8, "cmfsm_sub_16": cmfsm_sub_16, "cm_sub_4": cm_sub_4,
                "cm_sub_8": cm_sub_8, "cm_sub_16": cm_sub_16, "bilinear_cmf": bilinear_cmf,
                "bilinear_cmf_sub_8": bilinear_cmf_sub_8, "bilinear_cmf_sub_16": bilinear_cmf_sub_16}
_REFERENCE_NAMES = ("cmf", "cmfsm", "bilinear_cmf", "cmfsm_sub_8", "cmfsm_sub_16", "bilinear_cmf_sub_8",
                    "bilinear_cmf_sub_16", "cm_sub_16", "cm_sub_8", "cm_sub_4")


def _get_model_instance(name):
    if name in _IMPLEMENTED:
        return _IMPLEMENTED[name]
    if name in _REFERENCE_NAMES:
        raise NotImplementedError(
            "model '%s' is registered by the reference but not yet built in the B200-native package "
            "(SURVEY.md section 8f); available: %s" % (name, sorted(_IMPLEMENTED)))
    raise KeyError("Model {} not available".format(name))


def get_model(name):
    return _get_model_instance(name)()
