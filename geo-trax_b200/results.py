"""ultralytics-shaped result containers (``engine/results.py``: Results / Boxes / OBB) for the detector shim.

Only the attribute set the reference reads is provided:
``results[0].boxes`` with ``__len__``, ``.id`` (Tensor or None), ``.xywh``, ``.cls``, ``.conf``
(/root/reference/geotrax/extract.py:154-168, all ``torch.Tensor`` so that ``.detach().numpy(force=True)`` works),
``results[0].speed`` (extract.py:155-156, 266-268), plus what the host tracker callback needs
(``.cpu().numpy()``, ``.xyxy``, ``.xywhr``, ``__getitem__``, ``Results.update``, ``Results.orig_img``).
Values are produced by the CUDA library; these classes are views, they compute nothing but xyxy<->xywh.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch


def _xp(a):
    return torch if isinstance(a, torch.Tensor) else np


def _stack(parts, axis):
    return torch.stack(parts, axis) if isinstance(parts[0], torch.Tensor) else np.stack(parts, axis)


class _Base:
    """Tensor wrapper with the ``cpu()/numpy()/cuda()/to()`` chain and row indexing of ultralytics' BaseTensor: ``data`` is a
    ``torch.Tensor`` (what the reference reads, extract.py:154-168) or, after ``.numpy()``, a ``np.ndarray`` (what the host tracker
    indexes with boolean masks, ultralytics ``trackers/byte_tracker.py``)."""

    def __init__(self, data, orig_shape: Tuple[int, int]):
        if data.ndim == 1:
            data = data[None, :]
        self.data = data
        self.orig_shape = tuple(orig_shape)

    @property
    def shape(self):
        return self.data.shape

    def __len__(self) -> int:
        return int(self.data.shape[0])

    def __getitem__(self, idx):
        return self.__class__(self.data[idx], self.orig_shape)

    def cpu(self):
        return self if not (isinstance(self.data, torch.Tensor) and self.data.is_cuda) else self.__class__(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return self.__class__(self.data.detach().cpu().numpy(), self.orig_shape) if isinstance(self.data, torch.Tensor) else self

    def cuda(self):
        return self.__class__(torch.as_tensor(self.data).cuda(), self.orig_shape)

    def to(self, *a, **k):
        return self.__class__(torch.as_tensor(self.data).to(*a, **k), self.orig_shape)


class Boxes(_Base):
    """Rows ``[x1, y1, x2, y2, conf, cls]`` (6 columns) or ``[x1, y1, x2, y2, id, conf, cls]`` (7, tracked)."""

    def __init__(self, data, orig_shape):
        super().__init__(data, orig_shape)
        assert self.data.shape[-1] in (6, 7), f"Boxes expects 6 or 7 columns, got {self.data.shape[-1]}"
        self.is_track = self.data.shape[-1] == 7

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return self.data[:, -3] if self.is_track else None

    @property
    def xywh(self):
        b = self.xyxy
        return _stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1)

    def _wh4(self):
        h, w = self.orig_shape
        return torch.tensor([w, h, w, h], dtype=self.data.dtype) if isinstance(self.data, torch.Tensor) else np.array([w, h, w, h], self.data.dtype)

    @property
    def xyxyn(self):
        return self.xyxy / self._wh4()

    @property
    def xywhn(self):
        return self.xywh / self._wh4()


class OBB(_Base):
    """Rows ``[x, y, w, h, r, conf, cls]`` (7 columns) or ``[x, y, w, h, r, id, conf, cls]`` (8, tracked)."""

    def __init__(self, data, orig_shape):
        super().__init__(data, orig_shape)
        assert self.data.shape[-1] in (7, 8), f"OBB expects 7 or 8 columns, got {self.data.shape[-1]}"
        self.is_track = self.data.shape[-1] == 8

    @property
    def xywhr(self):
        return self.data[:, :5]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return self.data[:, -3] if self.is_track else None

    @property
    def xyxyxyxy(self):
        """(n, 4, 2) corner points."""
        xp = _xp(self.data)
        x, y, w, h, r = (self.data[:, i] for i in range(5))
        c, s = xp.cos(r), xp.sin(r)
        vx, vy = _stack([w / 2 * c, w / 2 * s], -1), _stack([-h / 2 * s, h / 2 * c], -1)
        ctr = _stack([x, y], -1)
        return _stack([ctr + vx + vy, ctr + vx - vy, ctr - vx - vy, ctr - vx + vy], 1)

    @property
    def xyxy(self):
        """Axis-aligned envelope of each rotated box (what the stabilizer mask uses)."""
        p = self.xyxyxyxy
        if isinstance(p, torch.Tensor):
            return torch.cat([p.min(1).values, p.max(1).values], -1)
        return np.concatenate([p.min(1), p.max(1)], -1)


class Results:
    """One image's detections.  ``boxes`` for task 'detect', ``obb`` for task 'obb' (the other is None)."""

    def __init__(self, orig_img: np.ndarray, path: str = "", names: Optional[Dict[int, str]] = None, boxes=None, obb=None,
                 speed: Optional[Dict[str, float]] = None):
        self.orig_img = orig_img
        self.orig_shape = tuple(orig_img.shape[:2])
        self.path = path
        self.names = names or {}
        self.boxes = Boxes(self._t(boxes), self.orig_shape) if boxes is not None else None
        self.obb = OBB(self._t(obb), self.orig_shape) if obb is not None else None
        self.speed = speed or {"preprocess": None, "inference": None, "postprocess": None}
        self._keys = ("boxes", "obb")

    @staticmethod
    def _t(a):
        """Detections live as torch tensors on a Results (the reference calls ``.detach().numpy(force=True)`` on them)."""
        return torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a

    def update(self, boxes=None, obb=None, **_unused):
        """ultralytics ``Results.update``: replace the detections (tracker output rows)."""
        if boxes is not None:
            self.boxes = Boxes(self._t(boxes), self.orig_shape)
        if obb is not None:
            self.obb = OBB(self._t(obb), self.orig_shape)

    def __len__(self) -> int:
        for k in self._keys:
            v = getattr(self, k)
            if v is not None:
                return len(v)
        return 0

    def __getitem__(self, idx):
        r = Results(self.orig_img, self.path, self.names, speed=self.speed)
        for k in self._keys:
            v = getattr(self, k)
            if v is not None:
                setattr(r, k, v[idx])
        return r

    def cpu(self):
        return self

    def numpy(self):
        return self
