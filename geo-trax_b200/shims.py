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


def _real_package_available(name: str) -> bool:
    """True when an importable, real (non-shim) distribution of `name` exists on sys.path."""
    import importlib.util

    have = sys.modules.get(name)
    if have is not None:
        return not getattr(have, "__geotrax_b200_shim__", False)
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install_shims(force: bool = False) -> None:
    """Makes ``from ultralytics import YOLO, RTDETR`` and ``from stabilo import Stabilizer`` resolve to the B200 classes (idempotent).

    * real ``ultralytics`` importable: ONLY ``YOLO`` / ``RTDETR`` are substituted, as attributes of the real package --
      ``ultralytics.trackers`` (BoT-SORT / ByteTrack, which the reference's `model.track(..., tracker=<yaml>)` relies on,
      /root/reference/geotrax/extract.py:153, cfg/default.yaml:361-379), ``ultralytics.utils`` and ``ultralytics.cfg`` stay the real ones.
    * no real ``ultralytics`` (this image): a minimal package with the names the reference imports is registered; it has NO
      ``trackers`` sub-package, so ``tracker.make_tracker`` cannot mistake it for the real one.
    * ``stabilo`` is always the B200 ``Stabilizer`` (the stage is entirely replaced).

    With ``force=False`` an already-imported real ``stabilo`` is left alone and a warning is logged."""
    from .stabilizer import Stabilizer
    from .yolo import RTDETR, YOLO

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__geotrax_b200_shim__ = True
        m.__path__ = []  # mark as package so that sub-module imports resolve through sys.modules
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    for name in [k for k, m in sys.modules.items() if k.split(".")[0] == "ultralytics" and getattr(m, "__geotrax_b200_shim__", False)]:
        del sys.modules[name]            # an earlier shim registration: rebuilt below
    if _real_package_available("ultralytics"):
        import importlib

        ul = importlib.import_module("ultralytics")
        ul.YOLO, ul.RTDETR = YOLO, RTDETR      # module attributes win over ultralytics' lazy __getattr__
        ul.__geotrax_b200_patched__ = True
        log.info("geotrax_b200: real ultralytics %s found -- YOLO / RTDETR substituted, trackers / utils / cfg left untouched",
                 getattr(ul, "__version__", "?"))
    else:
        checks = mod("ultralytics.utils.checks", check_yolo=check_yolo)
        files = mod("ultralytics.utils.files", increment_path=increment_path)
        utils = mod("ultralytics.utils", checks=checks, files=files)
        ul = mod("ultralytics", YOLO=YOLO, RTDETR=RTDETR, utils=utils, __version__="8.4.80+geotrax_b200")
        sys.modules.update({"ultralytics": ul, "ultralytics.utils": utils, "ultralytics.utils.checks": checks, "ultralytics.utils.files": files})
    have = sys.modules.get("stabilo")
    if have is not None and not getattr(have, "__geotrax_b200_shim__", False) and not force:
        log.warning("stabilo is already imported; not replacing it (install_shims(force=True) overrides)")
        return
    sys.modules["stabilo"] = mod("stabilo", Stabilizer=Stabilizer, __version__="1.2.3+geotrax_b200")
