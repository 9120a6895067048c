"""Engine sharing between the detector shim and the stabilizer shim.

In the reference loop one ``YOLO`` and one ``Stabilizer`` object see the same frame back to back
(/root/reference/geotrax/extract.py:153 then :177/:181).  Both shims resolve to ONE ``Engine`` (one gt_handle per GPU and
frame geometry), so the frame is uploaded and pre-processed once: ``gt_preprocess`` leaves the letterboxed tensor for the
detector and the half-resolution gray for ORB in the handle's workspaces.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from ._lib import GtError

# keys of the stabilo: block (/root/reference/geotrax/cfg/default.yaml:103-145) that map onto gt_config fields
STAB_DEFAULTS = dict(downsample_ratio=0.5, max_features=2000, ref_multiplier=2.0, mask_use=True, mask_margin_ratio=0.15,
                     filter_ratio=0.9, ransac_epipolar_threshold=2.0, ransac_max_iter=5000, match_query_frame="current", clahe=False,
                     ransac_space="working")

_engines: Dict[tuple, "object"] = {}
_latest_stab_cfg: Optional[dict] = None


def stab_key(cfg: dict) -> tuple:
    c = {**STAB_DEFAULTS, **{k: v for k, v in cfg.items() if k in STAB_DEFAULTS}}
    return tuple(sorted(c.items()))


def register_stab_cfg(cfg: dict) -> None:
    """Called by ``Stabilizer.__init__`` so that a detector engine created afterwards carries the right ORB/RANSAC setup."""
    global _latest_stab_cfg
    _latest_stab_cfg = dict(cfg)


def _engine_kwargs(cfg: dict) -> dict:
    c = {**STAB_DEFAULTS, **{k: v for k, v in cfg.items() if k in STAB_DEFAULTS}}
    return dict(downsample_ratio=float(c["downsample_ratio"]), max_features=int(c["max_features"]), ref_multiplier=float(c["ref_multiplier"]),
                mask_use=int(bool(c["mask_use"])), mask_margin_ratio=float(c["mask_margin_ratio"]), filter_ratio=float(c["filter_ratio"]),
                ransac_threshold=float(c["ransac_epipolar_threshold"]), ransac_max_iter=int(c["ransac_max_iter"]),
                query_is_current=int(c["match_query_frame"] != "reference"), clahe=int(bool(c["clahe"])),
                ransac_full_res=int(c["ransac_space"] == "full"))


def device_index(device) -> int:
    """ultralytics ``device`` values -> CUDA ordinal.  'cpu' is refused: this path has no CPU fallback."""
    if device is None or device == "":
        return 0
    if isinstance(device, (list, tuple)):
        device = device[0]
    if isinstance(device, str):
        d = device.lower().replace("cuda:", "").strip()
        if d in ("cpu", "mps"):
            raise GtError(f"device={device!r}: the B200 extraction path has no CPU fallback")
        if d == "cuda":
            return 0
        return int(d.split(",")[0])
    return max(int(device), 0)


def acquire(frame_hw: Tuple[int, int], imgsz: Optional[int], nc: int, task: str, device: int, max_batch: int,
            stab_cfg: Optional[dict], act_dtype: str = "fp16", max_det: int = 1000):
    """Engine for this geometry; created on first use.  ``stab_cfg`` None = the most recently constructed Stabilizer's."""
    from .engine import Engine

    cfg = stab_cfg if stab_cfg is not None else (_latest_stab_cfg or {})
    if imgsz is None:
        imgsz = max(frame_hw) // 2          # the exact-1/2 letterbox is the implemented geometry (default preset: 3840 -> 1920)
    key = (device, tuple(frame_hw), int(imgsz), int(nc), task, stab_key(cfg), act_dtype, int(max_det))
    eng = _engines.get(key)
    if eng is not None and eng.max_batch >= max_batch:
        return eng
    if eng is not None:
        eng.close()      # a larger batch was requested: the handle is re-created; owners (YOLO weights, Stabilizer reference) notice through
                         # `eng.h` / `_weights_owner` / `_ref_owner` and restore their state on the new handle
    eng = Engine(frame_hw=tuple(frame_hw), imgsz=int(imgsz), nc=int(nc), task=task, max_batch=int(max_batch), device=device, max_det=int(max_det),
                 act_dtype=act_dtype, **_engine_kwargs(cfg))
    eng._frame_token = None
    _engines[key] = eng
    return eng


def find_for_stabilizer(frame_hw: Tuple[int, int], device: int, stab_cfg: dict):
    """An existing engine (normally the detector's) with this frame geometry and stabilizer setup, else None."""
    want = stab_key(stab_cfg)
    for key, eng in _engines.items():
        if key[0] == device and key[1] == tuple(frame_hw) and key[5] == want and eng.h:
            return eng
    return None


def frame_token(frame: np.ndarray) -> tuple:
    """Cheap identity of a host frame: same object + same sparse checksum => the engine's workspaces already hold it."""
    return (id(frame), frame.__array_interface__["data"][0], frame.shape, int(frame[::97, ::89].sum(dtype=np.int64)))


def close_all() -> None:
    for eng in list(_engines.values()):
        eng.close()
    _engines.clear()
