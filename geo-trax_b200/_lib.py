"""ctypes binding of include/geotrax_b200.h.  The product path: no CPU fallback; a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgeotrax_b200.so")

GT_ABI_VERSION = 1
GT_INPUT_BGR24, GT_INPUT_NV12 = 0, 1
GT_TASK_DETECT, GT_TASK_OBB = 0, 1
GT_CODEC_H264, GT_CODEC_HEVC = 0, 1
GT_ACT_BF16, GT_ACT_FP16 = 0, 1
GT_MAX_KP = 8192
GT_ORB_LEVELS = 8


class GtError(RuntimeError):
    pass


class gt_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("frame_h", C.c_int32), ("frame_w", C.c_int32), ("max_batch", C.c_int32),
        ("imgsz", C.c_int32), ("nc", C.c_int32), ("task", C.c_int32), ("max_det", C.c_int32), ("max_nms", C.c_int32),
        ("downsample_ratio", C.c_float), ("max_features", C.c_int32), ("ref_multiplier", C.c_float),
        ("mask_use", C.c_int32), ("mask_margin_ratio", C.c_float), ("filter_ratio", C.c_float),
        ("ransac_threshold", C.c_float), ("ransac_max_iter", C.c_int32), ("query_is_current", C.c_int32),
        ("ransac_full_res", C.c_int32), ("seed", C.c_uint32), ("act_dtype", C.c_int32), ("clahe", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class gt_conv_desc(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32),
                ("stride", C.c_int32), ("act", C.c_int32)]


_H = C.c_void_p
_P = C.c_void_p
_i, _f, _u = C.c_int, C.c_float, C.c_uint32
_ip = C.POINTER(C.c_int32)

# every symbol include/geotrax_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "gt_default_config": (None, [C.POINTER(gt_config)]),
    "gt_create": (_i, [C.POINTER(gt_config), _i, C.POINTER(_H)]),
    "gt_destroy": (_i, [_H]),
    "gt_last_error": (C.c_char_p, [_H]),
    "gt_abi_version": (_i, []),
    "gt_conv_count": (_i, [_H]),
    "gt_conv_info": (_i, [_H, _i, C.POINTER(gt_conv_desc)]),
    "gt_load_weights": (_i, [_H, C.POINTER(_P), C.POINTER(_P), _i]),
    "gt_preprocess": (_i, [_H, _P, _i, _P]),
    "gt_prefetch_frames": (_i, [_H, _P, _i]),
    "gt_prefetch_frames_deferred": (_i, [_H, _P, _i]),
    "gt_set_input_format": (_i, [_H, _i]),
    "gt_nvdec_available": (_i, []),
    "gt_decoder_create": (_i, [_H, _i, _i, C.POINTER(_H)]),
    "gt_decoder_destroy": (_i, [_H]),
    "gt_decoder_feed": (_i, [_H, _P, C.c_size_t]),
    "gt_decoder_pending": (_i, [_H]),
    "gt_decoder_take": (_i, [_H, _i, C.POINTER(_P), _ip]),
    "gt_decoder_last_error": (C.c_char_p, [_H]),
    "gt_get_net_input": (_i, [_H, _i, _P, _ip, _ip]),
    "gt_get_gray": (_i, [_H, _i, _P, _ip, _ip]),
    "gt_detect": (_i, [_H, _i, _f, _f, _i, _u, _P, _P, _P, _P]),
    "gt_set_class_filter": (_i, [_H, _ip, _i]),
    "gt_get_health": (_i, [_H, C.POINTER(C.c_int64)]),
    "gt_get_candidate_counts": (_i, [_H, _i, _P]),
    "gt_get_raw_head": (_i, [_H, _i, _P, _ip, _ip]),
    "gt_get_feature": (_i, [_H, _i, _i, _P, _ip, _ip, _ip]),
    "gt_nms": (_i, [_H, _P, _i, _i, _i, _i, _f, _f, _i, _u, _i, _P, _P, _P, _P]),
    "gt_conv2d": (_i, [_H, _P, _i, _i, _i, _i, _P, _P, _i, _i, _i, _i, _P, _P, _i, _P]),
    "gt_set_reference": (_i, [_H, _i, _P, _i, _P]),
    "gt_stabilize": (_i, [_H, _i, _P, _P, _i, _P, _P, _P, _P]),
    "gt_warp_boxes": (_i, [_H, _P, _P, _i, _P]),
    "gt_warp_frames": (_i, [_H, _P, _P, _i, _P, _P]),
    "gt_orb_level_info": (_i, [_H, _i, _ip, _ip, _ip, _ip]),
    "gt_get_pyramid_level": (_i, [_H, _i, _i, _i, _P, _P]),
    "gt_get_keypoints": (_i, [_H, _i, _i, _i, _P, _P, _ip]),
    "gt_orb_get_candidates": (_i, [_H, _i, _i, _i, _i, _P, _P, _ip]),
    "gt_orb_detect": (_i, [_H, _P, _P, _i, _i, _P]),
    "gt_match": (_i, [_H, _P, _i, _P, _i, _P, _P, _P]),
    "gt_match_l2": (_i, [_H, _P, _i, _P, _i, _i, _P, _P, _P]),
    "gt_find_homography": (_i, [_H, _P, _P, _i, _f, _i, _P, _ip, _P]),
    "gt_extract_batch": (_i, [_H, _P, _i, _i, _f, _f, _i, _u, _P, _P, _i, _P, _P, _P, _P, _P, _P, _P]),
    "gt_extract_batch_async": (_i, [_H, _P, _i, _i, _f, _f, _i, _u, _P, _P, _i, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gt_wait": (_i, [_H, _i]),
    "gt_stage_times": (_i, [_H, _P]),
    "gt_launch_count": (C.c_int64, [_H]),
    "gt_conv_stack_stats": (_i, [_H, _P, _P]),
    "gt_conv_kernel_info": (_i, [_H, _ip, _ip]),
    "gt_conv_pair_count": (_i, [_H]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen the CUDA library and bind every symbol.  Raises GtError when it is missing (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise GtError(f"{p} not found: build it with `python __graft_entry__.py` / geo-trax_b200/build.py "
                      "(the B200 path has no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted from the header
        fn.restype = res
        fn.argtypes = args
    if lib.gt_abi_version() != GT_ABI_VERSION:
        raise GtError(f"ABI version mismatch: library {lib.gt_abi_version()} vs binding {GT_ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib
