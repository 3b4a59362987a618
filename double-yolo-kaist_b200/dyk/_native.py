"""ctypes binding of libdyk_b200.so (include/dyk_b200.h).

The shared object is built in-tree by ``double-yolo-kaist_b200/build.py`` (plain nvcc, sm_100a only)
and lives next to the package.  There is no fallback: if the library is missing or the device is not
a B200-class GPU every entry point raises, so a silent CPU / eager path can never stand in for it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent
import os as _os
LIB_PATH = Path(_os.environ.get("DYK_B200_LIB") or PKG_ROOT / "libdyk_b200.so")   # override: the -DDYK_CONV_PROFILE build

DYK_F16, DYK_BF16 = 0, 1
TRAIN_MAX_SLABS = 1024       # DYK_TRAIN_MAX_SLABS
STEM_WGRAD_STRIPS = 592      # DYK_STEM_WGRAD_STRIPS
DW_WGRAD_SLABS = 256         # DYK_DW_WGRAD_SLABS
ACT_IDS = {
    "linear": 0, "leaky": 1, "mish": 2, "relu": 3, "relu6": 4, "hard-swish": 5, "hard-sigmoid": 6,
}

_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ConvParams(C.Structure):
    """struct dyk_conv_params (include/dyk_b200.h)."""
    _fields_ = [
        ("x", _vp), ("x_pix_stride", _i64),
        ("w", _vp), ("scale", _vp), ("bias", _vp),
        ("y", _vp), ("y_pix_stride", _i64),
        ("res", _vp), ("res_pix_stride", _i64),
        ("N", _i32), ("H", _i32), ("W", _i32), ("Cin", _i32),
        ("Cout", _i32), ("Cout_store", _i32),
        ("kh", _i32), ("kw", _i32), ("stride", _i32), ("pad", _i32),
        ("act", _i32), ("dtype", _i32), ("upsample2x", _i32), ("out_f32", _i32),
        ("out_h", _i32), ("out_w", _i32), ("y_plane", _i32),
        ("x2", _vp), ("x2_pix_stride", _i64), ("x_wts_raw", _vp),
        ("w_image_stride", _i64), ("sm_limit", _i32),
    ]


# name -> (restype, argtypes); must list every symbol the header declares (tests check this).
SIGNATURES = {
    "dyk_abi_version": (_i32, []),
    "dyk_conv_params_size": (_i32, []),
    "dyk_last_error": (C.c_char_p, []),
    "dyk_check_device": (_i32, []),
    "dyk_conv2d_fwd": (_i32, [C.POINTER(ConvParams), _vp]),
    "dyk_conv2d_dual_source_supported": (_i32, [C.POINTER(ConvParams)]),
    "dyk_scale_weights_per_image": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp]),
    "dyk_fused_add_gated": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _i32, _vp]),
    "dyk_conv_set_profile": (_i32, [_vp]),
    "dyk_conv2d_stem_nchw_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64] + [_i32] * 11 + [_vp]),
    "dyk_conv2d_stem_nchw_resize_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64] + [_i32] * 13 + [_vp]),
    "dyk_frames_to_im2col32_resize": (_i32, [_vp, _vp] + [_i32] * 7 + [_vp]),
    "dyk_dwconv2d_fwd": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32,
                                _i32, _i32, _i32, _vp]),
    "dyk_fused_add": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _i32, _vp]),
    "dyk_fusion_weights": (_i32, [_vp, _vp, _i32, _vp]),
    "dyk_copy_slice": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp]),
    "dyk_maxpool2d": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_upsample_nearest": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_se_gate": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp]),
    "dyk_scale_channels": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "dyk_se_mlp": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    # ---- fp32-accurate mode
    "dyk_f32_split6": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp]),
    "dyk_f32_pack_split6": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "dyk_f32_stem_nchw_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64] + [_i32] * 10 + [_vp]),
    "dyk_f32_dwconv2d_fwd": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i64] + [_i32] * 8 + [_vp]),
    "dyk_f32_fused_add": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp]),
    "dyk_f32_maxpool2d": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_f32_upsample_nearest": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_f32_se": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "dyk_yolo_decode": (_i32, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _f32, _i32, _i64,
                               _i64, _i32, _vp]),
    "dyk_nms_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "dyk_nms_batched": (_i32, [_vp, _i32, _i32, _i32, _f32, C.c_double, _i32, C.c_uint64, _i32, _i32, _vp, _vp,
                               _vp, _i64, _vp]),
    # ---- training path
    "dyk_bn_train_stats": (_i32, [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dyk_bn_act_apply": (_i32, [_vp, _i64, _vp, _vp, _i32, _vp, _i64, _i64, _i32, _i32, _vp]),
    "dyk_bn_act_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _i64, _vp, _vp,
                              _vp, _vp, _vp]),
    "dyk_chan_sum": (_i32, [_vp, _i64, _i64, _i32, _i32, _vp, _i32, _vp, _vp]),
    "dyk_axpby": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp]),
    "dyk_fusion_weights_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dyk_maxpool2d_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "dyk_upsample_nearest_bwd": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_se_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp,
                          _vp, _vp, _i32, _i32, _vp, _vp]),
    "dyk_yolo_train_bwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp]),
    "dyk_pack_weights_dgrad": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_conv2d_wgrad_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "dyk_conv2d_wgrad": (_i32, [_vp, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                _i32, _vp, _i64, _vp]),
    "dyk_frames_to_nhwc8": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_conv2d_stem_wgrad": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                     _vp, _vp]),
    "dyk_dwconv2d_dgrad": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_dwconv2d_wgrad": (_i32, [_vp, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "dyk_scale_coords": (_i32, [_vp, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "dyk_yolo_build_targets": (_i32, [_vp, _i32, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "dyk_yolo_loss_workspace_floats": (_i64, [_i32, _i32, _i64]),
    "dyk_yolo_loss_head": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32,
                                  _f32, _f32, _f32, _f32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "dyk_yolo_loss_scale_grad": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "dyk_frames_to_im2col32": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_pack_weights_multi": (_i32, [_vp, _i32, _i32, _i32, _vp]),
    "dyk_fold_bn_multi": (_i32, [_vp, _i32, _vp]),
    "dyk_pack_weights_ohwi": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_optim_block_elems": (_i32, []),
    "dyk_optim_sgd_multi": (_i32, [_vp, _i32, _i64, _f32, _f32, _f32, _f32, _i32, _i32, _vp, _vp, _vp]),
    "dyk_optim_adam_multi": (_i32, [_vp, _i32, _i64, _f32, _f32, _f32, _f32, _f32, _i64, _vp, _vp, _vp]),
    "dyk_nchw_f32_to_nhwc": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dyk_nhwc_to_nchw_f32": (_i32, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
}

_lib = None
_device_ok = False


class NativeError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads libdyk_b200.so and declares the prototypes.  Raises when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python {PKG_ROOT / 'build.py'}` "
            "(there is no CPU or eager fallback for the dual-stream YOLO hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library went out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.dyk_conv_params_size() != C.sizeof(ConvParams):
        raise NativeError(f"struct dyk_conv_params is {lib.dyk_conv_params_size()} bytes in {LIB_PATH.name} but {C.sizeof(ConvParams)} "
                          "in dyk/_native.py: header and binding are out of sync")
    _lib = lib
    return lib


def check_device() -> None:
    """Raises unless the current CUDA device can run the sm_100a kernels."""
    global _device_ok
    if _device_ok:
        return
    lib = load()
    rc = lib.dyk_check_device()
    if rc != 0:
        raise NativeError(f"dyk_check_device failed ({rc}): {lib.dyk_last_error().decode()}")
    _device_ok = True


def call(name: str, *args) -> None:
    """Calls an int-returning entry point and raises NativeError with dyk_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise NativeError(f"{name} failed ({rc}): {lib.dyk_last_error().decode()}")


def launch_count() -> int:
    return _launches[0]


_launches = [0]


def count_launches(n: int = 1) -> None:
    _launches[0] += n
