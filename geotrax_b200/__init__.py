"""Importable alias of the hyphenated source directory ``geo-trax_b200/`` (a hyphen cannot appear in a Python name).

``import geotrax_b200`` resolves sub-modules from ``<repo>/geo-trax_b200/``; nothing else lives here.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "geo-trax_b200")
__path__.insert(0, _SRC)

from ._lib import GtError, LIB_PATH, load_library  # noqa: E402,F401

_LAZY = {"Engine": "engine", "Decoder": "engine", "DeviceFrames": "engine", "YOLO": "yolo", "RTDETR": "yolo", "Results": "results", "Boxes": "results", "OBB": "results",
         "Stabilizer": "stabilizer", "install_shims": "shims"}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)
