"""ctypes binding of libnerfpp_b200.so -- the reference-side stub for include/nerfpp_b200.h.

Every symbol declared in the header is bound here with its exact C signature; tensors cross the
boundary as raw device pointers (``tensor.data_ptr()``) plus sizes and the current CUDA stream.
There is no fallback: if the library is missing or a call fails this raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NERFPP_B200_LIB") or os.path.join(HERE, "libnerfpp_b200.so")   # override: A/B diagnostics only
NLAYERS = 12
ABI_VERSION = 1

FIELD_TC, FIELD_SIMT, FIELD_TC_SPLIT = 0, 1, 2
DEPTH_NONE, DEPTH_MSE, DEPTH_L1, DEPTH_KL = 0, 1, 2, 3


class NetParams(Structure):
    _fields_ = [("w", c_void_p * NLAYERS), ("b", c_void_p * NLAYERS)]


class NetGrads(Structure):
    _fields_ = [("w", c_void_p * NLAYERS), ("b", c_void_p * NLAYERS)]


class RenderOut(Structure):
    _fields_ = [(k, c_void_p) for k in ("rgb", "fg_weights", "bg_weights", "fg_dists", "fg_rgb", "fg_depth",
                                        "bg_rgb", "bg_depth", "bg_lambda", "depth")]


class Mip360MlpParams(Structure):
    _fields_ = [("kernel", c_void_p * 12), ("bias", c_void_p * 12)]


class RenderGrads(Structure):
    _fields_ = [(k, c_void_p) for k in ("rgb", "fg_weights", "bg_weights", "fg_dists", "fg_rgb", "fg_depth",
                                        "bg_rgb", "bg_depth", "bg_lambda", "depth")]


P = c_void_p
SIGNATURES = {
    "nerfpp_abi_version": (c_int, []),
    "nerfpp_last_error": (c_char_p, []),
    "nerfpp_intersect_sphere": (c_int, [P, P, c_int, P, P, P]),
    "nerfpp_coarse_depths": (c_int, [P, P, P, c_int, c_int, P, P, P, P, P]),
    "nerfpp_perturb_samples": (c_int, [P, P, c_int, c_int, P, P]),
    "nerfpp_sample_pdf": (c_int, [P, c_int, P, c_int, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "nerfpp_sample_cdf": (c_int, [P, c_int, P, c_int, P, c_int, c_int, c_int, c_int, P, P, P]),
    "nerfpp_resample_merge": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P]),
    "nerfpp_resample_merge_pair": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "nerfpp_packed_bytes": (c_int64, [c_int, c_int]),
    "nerfpp_pack_weights": (c_int, [POINTER(NetParams), c_int, c_int, P, P]),
    "nerfpp_field_forward": (c_int, [P, c_int, c_int, P, P, P, c_int, c_int, P, P, P, P]),
    "nerfpp_field_train_workspace_bytes": (c_int64, [c_int, c_int]),
    "nerfpp_field_forward_train": (c_int, [P, c_int, P, P, P, c_int, c_int, P, P, P, P, P]),
    "nerfpp_depth2pts_outside": (c_int, [P, P, P, c_int64, P, P, P]),
    "nerfpp_composite": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, POINTER(RenderOut), P]),
    "nerfpp_composite_backward": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, POINTER(RenderOut), POINTER(RenderOut),
                                          P, P, P, P, P]),
    "nerfpp_forward_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "nerfpp_forward": (c_int, [P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, POINTER(RenderOut), P, P]),
    "nerfpp_forward_train_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "nerfpp_forward_train": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, POINTER(RenderOut), P, P, P]),
    "nerfpp_backward_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "nerfpp_backward": (c_int, [POINTER(NetParams), POINTER(NetParams), P, P, P, P, c_int, c_int, c_int, POINTER(RenderOut),
                                POINTER(RenderOut), P, P, POINTER(NetGrads), POINTER(NetGrads), P, P]),
    "nerfpp_loss_workspace_bytes": (c_int64, []),
    "nerfpp_loss": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_float, P, P, P]),
    "nerfpp_depth_loss": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_float, P, P, P]),
    "nerfpp_depth_loss_backward": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_float, P, P, P, P]),
    "nerfpp_gen_rays": (c_int, [P, P, c_float, c_int, P, c_int64, P, P, P, P, P, P, P, P, P, P]),
    "nerfpp_decode_pixels": (c_int, [P, c_int, c_int64, c_float, c_float, c_float, P, P]),
    "nerfpp_image_metrics": (c_int, [P, P, P, P, c_int64, c_float, c_float, P, P, P]),
    "mip360_sample_intervals": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_float, c_float, P, P]),
    "mip360_lossfun_outer": (c_int, [P, P, P, P, c_int, c_int, c_int, c_float, P, P, P, P]),
    "mip360_max_dilate": (c_int, [P, P, c_int, c_int, c_float, c_float, c_float, c_int, c_int, c_float, P, P, P]),
    "mip360_lossfun_distortion": (c_int, [P, P, c_int, c_int, P, P, P, P, P]),
    "mip360_compute_alpha_weights": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, P]),
    "mip360_volumetric_rendering": (c_int, [P, P, P, P, c_int, P, c_int, c_int, P, P, P]),
    "mip360_depth_loss_workspace_bytes": (c_int64, [c_int]),
    "mip360_depth_loss": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_float, P, P, P]),
    "mip360_mlp_packed_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "mip360_mlp_pack": (c_int, [POINTER(Mip360MlpParams), c_int, c_int, c_int, c_int, P, P]),
    "mip360_field_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int]),
    "mip360_field_forward": (c_int, [P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, c_int, P, P, P, P, P]),
    "mip360_cast_encode": (c_int, [P, P, P, P, P, P, P, c_int, c_int, P, P, P, P, P, P, P, P]),
    "mip360_dense_f16": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "mip360_resample_logits": (c_int, [P, P, c_int, c_int, c_float, c_float, P, P]),
    "mip360_resample_level": (c_int, [P, c_int, P, c_int, c_int, c_int, c_float, c_float, P, c_int, P, c_float, c_int, c_float, c_float, P, P]),
}
OPTIONAL = {}
_lib = None


class NerfppError(RuntimeError):
    pass


def lib():
    """Loads the shared library once; raises NerfppError when it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfppError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    for name, (res, args) in OPTIONAL.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    if L.nerfpp_abi_version() != ABI_VERSION:
        raise NerfppError("ABI mismatch: library %d, binding %d" % (L.nerfpp_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().nerfpp_last_error()
        raise NerfppError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))
