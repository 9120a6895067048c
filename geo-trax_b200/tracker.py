"""Host-side tracker hand-off: a BOUNDARY of the hot path, not part of it.

The reference tracks with ultralytics' BoT-SORT on the CPU (``trackers/track.py:on_predict_postprocess_end``), fed by the
detector's ``Boxes`` and the frame (/root/reference/geotrax/extract.py:153 ``model.track(..., tracker=<yaml>)``; the yaml is written
by /root/reference/geotrax/utils/config_utils.py:197-226 from cfg/default.yaml:361-379).  That tracker stays host code and is not
rewritten here.  ``make_tracker`` builds the REAL ultralytics tracker (``TRACKER_MAP[tracker_type]``) from that yaml whenever the
package is importable.  When it is not (this image), it raises -- track ids are the product's main output and must not silently come
from something else -- unless the caller opts in to ``GreedyIoUTracker``, a deliberately small stand-in with the same
``update(det, img, feats) -> rows [x1,y1,x2,y2,id,score,cls,det_idx]`` contract: ``tracker='greedy-iou'`` or the environment variable
``GEOTRAX_B200_TRACKER=greedy-iou`` (for the unmodified CLI).
"""
from __future__ import annotations

import logging

import numpy as np

log = logging.getLogger("geotrax_b200")


def _iou_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)), np.float32)
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return (inter / np.maximum(aa[:, None] + ab[None, :] - inter, 1e-9)).astype(np.float32)


class GreedyIoUTracker:
    """Greedy IoU association in descending-score order; ids start at 1; tracks die after ``max_age`` misses."""

    def __init__(self, iou_thr: float = 0.3, max_age: int = 30):
        self.iou_thr, self.max_age = iou_thr, max_age
        self.boxes = np.zeros((0, 4), np.float32)
        self.ids = np.zeros((0,), np.int64)
        self.age = np.zeros((0,), np.int64)
        self.next_id = 1

    def reset(self):
        self.__init__(self.iou_thr, self.max_age)

    def update(self, det, img=None, feats=None) -> np.ndarray:
        """det: object with ``.xyxy``, ``.conf``, ``.cls`` (numpy) -> (m, 8) rows [x1,y1,x2,y2,id,score,cls,det_idx]."""
        xyxy = np.asarray(det.xyxy, np.float32).reshape(-1, 4)
        conf = np.asarray(det.conf, np.float32).reshape(-1)
        cls = np.asarray(det.cls, np.float32).reshape(-1)
        n = len(xyxy)
        assigned = np.full(n, -1, np.int64)
        iou = _iou_matrix(xyxy, self.boxes)
        taken = np.zeros(len(self.boxes), bool)
        for i in np.argsort(-conf, kind="stable"):
            if iou.shape[1] == 0:
                break
            row = np.where(taken, -1.0, iou[i])
            j = int(row.argmax())
            if row[j] >= self.iou_thr:
                assigned[i] = j
                taken[j] = True
        new = assigned < 0
        ids = np.empty(n, np.int64)
        ids[~new] = self.ids[assigned[~new]]
        ids[new] = np.arange(self.next_id, self.next_id + int(new.sum()))
        self.next_id += int(new.sum())
        keep = ~taken & (self.age + 1 <= self.max_age)       # unmatched old tracks survive a while
        self.boxes = np.concatenate([xyxy, self.boxes[keep]])
        self.ids = np.concatenate([ids, self.ids[keep]])
        self.age = np.concatenate([np.zeros(n, np.int64), self.age[keep] + 1])
        return np.concatenate([xyxy, ids[:, None].astype(np.float32), conf[:, None], cls[:, None], np.arange(n, dtype=np.float32)[:, None]], 1)


STAND_IN = "greedy-iou"


def _load_tracker_cfg(tracker_cfg):
    """The tracker yaml the reference passes -> attribute namespace (ultralytics' own loader when present, else PyYAML)."""
    import types

    if isinstance(tracker_cfg, dict):
        data = dict(tracker_cfg)
    else:
        try:
            from ultralytics.utils import YAML  # type: ignore

            data = YAML.load(str(tracker_cfg))
        except ImportError:
            import yaml

            with open(str(tracker_cfg)) as f:
                data = yaml.safe_load(f)
    try:
        from ultralytics.utils import IterableSimpleNamespace  # type: ignore

        return IterableSimpleNamespace(**data)
    except ImportError:
        return types.SimpleNamespace(**data)


def make_tracker(tracker_cfg=None, frame_rate: int = 30):
    """ultralytics BOTSORT / BYTETracker built from the yaml the reference passes (``tracker=<path>``).

    ``tracker_cfg``: path / dict of a tracker yaml (``tracker_type: botsort | bytetrack`` ...), ``None`` (ultralytics' default
    ``botsort.yaml``), or ``'greedy-iou'`` for the stand-in.  Raises ``GtError`` when the real tracker is unavailable and the
    stand-in was not asked for."""
    import importlib
    import os

    from ._lib import GtError

    if tracker_cfg == STAND_IN or (tracker_cfg is None and os.environ.get("GEOTRAX_B200_TRACKER", "") == STAND_IN):
        return GreedyIoUTracker()
    try:
        real = importlib.import_module("ultralytics.trackers.track")
    except ImportError as err:
        if os.environ.get("GEOTRAX_B200_TRACKER", "") == STAND_IN:
            log.warning("ultralytics trackers not importable: GEOTRAX_B200_TRACKER=greedy-iou selects the stand-in tracker (ids only; not BoT-SORT)")
            return GreedyIoUTracker()
        raise GtError("ultralytics.trackers is not importable, so BoT-SORT / ByteTrack cannot run.  Install ultralytics (the shims leave its "
                      "trackers untouched), or opt in to the IoU stand-in with tracker='greedy-iou' / GEOTRAX_B200_TRACKER=greedy-iou") from err
    if getattr(real, "__geotrax_b200_shim__", False):
        raise GtError("ultralytics.trackers resolves to a geotrax_b200 shim module; re-run install_shims()")
    if tracker_cfg is None:
        root = os.path.dirname(importlib.import_module("ultralytics").__file__)
        tracker_cfg = os.path.join(root, "cfg", "trackers", "botsort.yaml")
    cfg = _load_tracker_cfg(tracker_cfg)
    kind = getattr(cfg, "tracker_type", None)
    if kind not in real.TRACKER_MAP:
        raise GtError(f"tracker_type {kind!r} is not one of ultralytics' TRACKER_MAP {sorted(real.TRACKER_MAP)}")
    return real.TRACKER_MAP[kind](args=cfg, frame_rate=frame_rate)
