"""Import-name injection: makes ``from ultralytics import YOLO, RTDETR`` / ``from stabilo import Stabilizer`` resolve to the
B200 implementations, so the reference's ``geotrax extract`` / ``batch`` CLI runs unchanged on top of them.

    import geotrax_b200; geotrax_b200.install_shims()
    from geotrax import cli; cli.main()          # /root/reference/geotrax/cli.py, unmodified

Names the reference imports (/root/reference/geotrax/extract.py:90-94, utils/config_utils.py:18-21, utils/registration.py:18):
``ultralytics.{YOLO, RTDETR}``, ``ultralytics.utils.checks.check_yolo``, ``ultralytics.utils.files.increment_path``,
``stabilo.Stabilizer``.
"""
from __future__ import annotations

import logging
import sys
import types
from pathlib import Path

log = logging.getLogger("geotrax_b200")


def check_yolo(verbose: bool = True, device=""):
    """ultralytics.utils.checks.check_yolo: environment summary.  Here: the GPU the library will run on (raises without one)."""
    import torch

    from . import session
    idx = session.device_index(device)
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the B200 extraction path has no CPU fallback")
    p = torch.cuda.get_device_properties(idx)
    log.info("geotrax_b200: device %d %s, %d SMs, %.0f GB", idx, p.name, p.multi_processor_count, p.total_memory / 2 ** 30)


def increment_path(path, exist_ok: bool = False, sep: str = "", mkdir: bool = False) -> Path:
    """ultralytics.utils.files.increment_path: runs/exp -> runs/exp2, runs/exp3, ... (files keep their suffix)."""
    path = Path(path)
    if path.exists() and not exist_ok:
        path, suffix = (path.with_suffix(""), path.suffix) if path.is_file() else (path, "")
        for n in range(2, 9999):
            p = f"{path}{sep}{n}{suffix}"
            if not Path(p).exists():
                break
        path = Path(p)
    if mkdir:
        path.mkdir(parents=True, exist_ok=True)
    return path


def install_shims(force: bool = False) -> None:
    """Registers the fake ``ultralytics`` and ``stabilo`` packages in sys.modules (idempotent).

    With ``force=False`` an already-imported real package is left alone and a warning is logged."""
    from .stabilizer import Stabilizer
    from .yolo import RTDETR, YOLO

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__geotrax_b200_shim__ = True
        m.__path__ = []  # mark as package so that sub-module imports resolve through sys.modules
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    if not force:
        for name in ("ultralytics", "stabilo"):
            have = sys.modules.get(name)
            if have is not None and not getattr(have, "__geotrax_b200_shim__", False):
                log.warning("%s is already imported; not replacing it (install_shims(force=True) overrides)", name)
                return
    checks = mod("ultralytics.utils.checks", check_yolo=check_yolo)
    files = mod("ultralytics.utils.files", increment_path=increment_path)
    utils = mod("ultralytics.utils", checks=checks, files=files)
    trackers = mod("ultralytics.trackers")
    track = mod("ultralytics.trackers.track")
    trackers.track = track
    ul = mod("ultralytics", YOLO=YOLO, RTDETR=RTDETR, utils=utils, trackers=trackers, __version__="8.4.80+geotrax_b200")
    sys.modules.update({"ultralytics": ul, "ultralytics.utils": utils, "ultralytics.utils.checks": checks, "ultralytics.utils.files": files,
                        "ultralytics.trackers": trackers, "ultralytics.trackers.track": track})
    sys.modules["stabilo"] = mod("stabilo", Stabilizer=Stabilizer, __version__="1.2.3+geotrax_b200")
