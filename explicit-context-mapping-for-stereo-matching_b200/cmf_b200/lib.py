"""ctypes binding of libcmfb200.so (the C ABI declared in include/cmfb200.h).

There is NO fallback: if the shared library is missing or was built from a different header, importing
the kernels raises.  Build it with `python __graft_entry__.py` (or `make -C <pkg>/csrc`).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcmfb200.so")
ABI_VERSION = 3

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_LL = _c.c_longlong
_F = _c.c_float
_D = _c.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/cmfb200.h one to one
SIGNATURES = {
    "cmfb200_abi_version": [],
    "cmfb200_last_error": [],
    "cmfb200_launch_count": [],
    "cmfb200_cost_volume_concat_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_cost_volume_concat_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_pack_conv3d_weight": [_P, _P, _I, _I, _I, _P],
    "cmfb200_conv3d_k3_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_deconv3d_k3s2_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_pack_igemm_weight_bf16": [_P, _P, _I, _I, _I, _P],
    "cmfb200_conv3d_igemm_bf16_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_conv3d_igemm_cout1_bf16_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_deconv3d_igemm_bf16_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_conv3d_s2_igemm_bf16_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_c8_parity_split": [_P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_cost_volume_concat_c8_bf16": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_gn_apply_c8_bf16": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "cmfb200_c8_bf16_to_f32": [_P, _P, _I, _I, _LL, _P],
    "cmfb200_f32_to_c8_bf16": [_P, _P, _I, _I, _LL, _P],
    "cmfb200_pack_conv2d_weight": [_P, _P, _I, _I, _I, _P],
    "cmfb200_conv2d_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_spp_pool_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "cmfb200_spp_upsample_concat_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_gn_stats": [_P, _P, _I, _I, _LL, _P],
    "cmfb200_gn_apply": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _LL, _F, _I, _P],
    "cmfb200_conv2d_rows_fwd": [_P, _P, _P, _P] + [_I] * 10 + [_P],
    "cmfb200_conv3d_k3_rows_fwd": [_P, _P, _P, _P] + [_I] * 9 + [_P],
    "cmfb200_deconv3d_k3s2_rows_fwd": [_P, _P, _P, _P] + [_I] * 7 + [_P],
    "cmfb200_conv3d_cout1_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_gn_bwd": [_P] * 8 + [_I, _I, _I, _LL, _F, _P],
    "cmfb200_ctxmap_weights_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_ctxmap_weights5_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "cmfb200_softargmin_ctxmap_bwd": [_P] * 11 + [_I] * 5 + [_P],
    "cmfb200_softargmin_ctxmap5_fwd": [_P] * 8 + [_I] * 5 + [_P],
    "cmfb200_spp_upsample_concat_sized_fwd": [_P] * 7 + [_I] * 12 + [_P],
    "cmfb200_ctxmap_weights3_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "cmfb200_trilinear_softargmin_fwd": [_P] * 6 + [_I] * 7 + [_P],
    "cmfb200_volume_mapping_fwd": [_P] * 8 + [_I] * 5 + [_P],
    "cmfb200_ctxmap_weights_bwd": [_P] * 11 + [_I, _I, _I, _I, _P],
    "cmfb200_softargmin_ctxmap_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_cost_volume_concat_c8s3": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "cmfb200_cost_volume_corr_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_copy_2d": [_P, _LL, _P, _LL, _LL, _LL, _P],
    "cmfb200_sum_peers_f64": [_P, _P, _I, _I, _D, _P],
    "cmfb200_masked_smooth_l1_fwd": [_P, _P, _P, _P, _P, _LL, _F, _P],
    "cmfb200_masked_smooth_l1_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _LL, _F, _P],
    "cmfb200_conv_wgrad": [_P, _P, _P] + [_I] * 10 + [_P],
    "cmfb200_gn_apply_tc3_padded": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _LL, _F, _I, _I, _I, _I, _P, _P, _P, _I, _I,
                                    _P],
    "cmfb200_conv_tc3_s2_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_deconv_tc3_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_conv_tc3_s2_rows_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_deconv_tc3_rows_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_cost_volume_concat_c8s3_padded": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cmfb200_pack_tc3_weight": [_P, _P, _I, _I, _I, _I, _P],
    "cmfb200_conv_tc3_fwd": [_P, _P, _P, _P] + [_I] * 10 + [_P],
    "cmfb200_conv_tc3_rows_fwd": [_P, _P, _P, _P] + [_I] * 12 + [_P],
    "cmfb200_gn_apply_tc3": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _LL, _F, _I, _P],
}
_RESTYPES = {"cmfb200_last_error": _c.c_char_p, "cmfb200_launch_count": _c.c_ulonglong}

_lib = None
_lock = threading.Lock()


class CmfB200Error(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises CmfB200Error if the library is unusable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CmfB200Error(
                "libcmfb200.so not found at %s -- build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a); there is no CPU or PyTorch fallback for the cmfsm hot path" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise CmfB200Error("libcmfb200.so does not export %s (stale build?)" % name) from e
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _I)
        got = lib.cmfb200_abi_version()
        if got != ABI_VERSION:
            raise CmfB200Error("libcmfb200.so ABI version %d != binding version %d" % (got, ABI_VERSION))
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().cmfb200_last_error()
        raise CmfB200Error("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(load().cmfb200_launch_count())
