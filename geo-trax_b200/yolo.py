"""``ultralytics.YOLO``-shaped front end of the CUDA detector (drop-in for /root/reference/geotrax/extract.py:217-236, :153).

    model = YOLO(model=path, task='detect')           # extract.py:222
    model.model.yaml_file / model.model.yaml          # extract.py:223 ('rtdetr' substring switches class)
    model.names                                       # utils/config_utils.py:259
    results = model.track(frame, **cfg['ultralytics'], persist=True)   # extract.py:153

Every key of the YAML ``ultralytics:`` block arrives as a keyword (default.yaml:229-354); the ones that drive the kernels are
``imgsz, conf, iou, max_det, classes, agnostic_nms, device`` -- the rest are display/IO switches and are accepted and ignored.
All arithmetic happens in libgeotrax_b200.so (gt_preprocess + gt_detect); a missing library or GPU raises, nothing falls
back to the CPU.  ``half: false`` in the preset selects ultralytics' fp32 path; this implementation stores activations and
weights in 16 bit with f32 accumulation (north_star), checked against the fp32 oracle to 1e-2.  The storage format is **fp16**
(11-bit significand: 4-6e-3 on the raw head; bf16's 8 bits give 2.4e-2 and miss the gate -- DESIGN.md section 2).  fp16's range
ends at 65,504: every decode counts head rows that arrive as inf / NaN (``Engine.health()``), and on the first such row this front
end re-runs the frame on a bf16 engine (8-bit exponent, fp32's range) and stays there -- loudly, never silently wrong.
"""
from __future__ import annotations

import os
import types
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import session, weights
from ._lib import GtError
from .results import Results

import logging

log = logging.getLogger("geotrax_b200")

DEFAULT_NAMES = {0: "car", 1: "bus", 2: "truck", 3: "motorcycle"}


def _parse_synthetic(spec: str) -> dict:
    """'synthetic:nc=4,seed=0,cls_bias=-4.4,task=detect,hw=2160x3840,imgsz=1920' -> dict (random-init weights; no network)."""
    out = dict(nc=4, seed=0, cls_bias=-4.4, task=None, hw=(2160, 3840), imgsz=1920)
    body = spec.split(":", 1)[1] if ":" in spec else ""
    for kv in filter(None, body.split(",")):
        k, v = kv.split("=")
        if k == "hw":
            out["hw"] = tuple(int(x) for x in v.lower().split("x"))
        elif k == "task":
            out["task"] = v
        elif k == "cls_bias":
            out[k] = float(v)
        else:
            out[k] = int(v)
    return out


class YOLO:
    def __init__(self, model: Union[str, os.PathLike] = "yolov8s.pt", task: Optional[str] = None, verbose: bool = False):
        self.ckpt_path = str(model)
        if self.ckpt_path.startswith("synthetic"):
            s = _parse_synthetic(self.ckpt_path)
            self.task = s["task"] or task or "detect"
            self.nc = s["nc"]
            self._sd = weights.random_state_dict(self.nc, self.task, seed=s["seed"], cls_bias=s["cls_bias"], frame_hw=s["hw"], imgsz=s["imgsz"])
            self.names: Dict[int, str] = dict(DEFAULT_NAMES) if self.nc == 4 else {i: str(i) for i in range(self.nc)}
            yaml_cfg = {"yaml_file": "yolov8s-obb.yaml" if self.task == "obb" else "yolov8s.yaml", "nc": self.nc}
        else:
            if not os.path.exists(self.ckpt_path):
                raise FileNotFoundError(f"detector weights '{self.ckpt_path}' not found (no network: hf:// and auto-download are unavailable)")
            sd, names, ck_task, nc = weights.load_pt(self.ckpt_path)
            self._sd, self.nc = sd, nc
            self.task = task or ck_task
            if ck_task != self.task:
                raise GtError(f"checkpoint is a '{ck_task}' model but task='{self.task}' was requested")
            self.names = {int(k): str(v) for k, v in names.items()} or {i: str(i) for i in range(nc)}
            yaml_cfg = {"yaml_file": "yolov8s-obb.yaml" if ck_task == "obb" else "yolov8s.yaml", "nc": nc}
        if self.task not in ("detect", "obb"):
            raise GtError(f"task '{self.task}' is not implemented (detect and obb are)")
        self._folded = weights.fold(self._sd, self.nc, self.task)   # raises if the checkpoint is not a YOLOv8s graph
        self.model = types.SimpleNamespace(yaml_file=yaml_cfg["yaml_file"], yaml=yaml_cfg, names=self.names, stride=[8, 16, 32])
        self.overrides = {"task": self.task, "model": self.ckpt_path}
        self.max_batch = 1            # raise before the first call to batch frames (list / 4-D array sources)
        self.act_dtype = "fp16"
        self._engine = None
        self._tracker = None
        self.trackers: List = []      # ultralytics keeps predictor.trackers; same idea

    # -- engine ------------------------------------------------------------------------------------------------------------
    def _get_engine(self, frame_hw, imgsz, device, max_det, nb):
        if isinstance(imgsz, (list, tuple)):
            imgsz = max(imgsz)
        imgsz = max(32, -(-int(imgsz) // 32) * 32)      # ultralytics check_imgsz: rounded up to a multiple of the stride
        eng = session.acquire(tuple(frame_hw), int(imgsz), self.nc, self.task, session.device_index(device), max(self.max_batch, nb), None,
                              act_dtype=self.act_dtype, max_det=int(max_det))
        if self._engine is not eng or not getattr(eng, "_weights_owner", None) is self:
            eng.load_weights(self._folded)
            eng._weights_owner = self
            self._engine = eng
        return eng

    @staticmethod
    def _as_batch(source) -> np.ndarray:
        if isinstance(source, np.ndarray) and source.ndim == 3:
            return source[None]
        if isinstance(source, np.ndarray) and source.ndim == 4:
            return source
        if isinstance(source, (list, tuple)) and len(source) and isinstance(source[0], np.ndarray):
            return np.stack(source)
        raise GtError("source must be a BGR uint8 frame (H,W,3), a list of frames or a (B,H,W,3) array; file/stream sources are read by the caller "
                      "(extract.py uses cv2.VideoCapture)")

    # -- inference ---------------------------------------------------------------------------------------------------------
    def predict(self, source=None, stream: bool = False, conf: float = 0.25, iou: float = 0.7, imgsz=1920, max_det: int = 1000,
                classes: Optional[Sequence[int]] = None, agnostic_nms: bool = False, device=None, **_ignored) -> List[Results]:
        import torch

        frames = self._as_batch(source)
        if frames.dtype != np.uint8 or frames.shape[-1] != 3:
            raise GtError(f"frames must be uint8 BGR, got {frames.dtype} {frames.shape}")
        frames = np.ascontiguousarray(frames)
        nb = frames.shape[0]
        eng = self._get_engine(frames.shape[1:3], imgsz, device, max_det, nb)
        out: List[Results] = []
        b0 = 0
        while b0 < nb:
            chunk = frames[b0:b0 + eng.max_batch]
            h0 = eng.health()
            eng.preprocess(chunk)
            eng._frame_token = session.frame_token(source) if (nb == 1 and isinstance(source, np.ndarray) and source.ndim == 3) else None
            boxes, counts = eng.detect(len(chunk), conf=conf, iou=iou, agnostic=agnostic_nms, classes=classes)
            if eng.health() > h0:       # inf / NaN reached the head: fp16 range exceeded by this checkpoint's activations
                if self.act_dtype != "fp16":
                    raise GtError("detector produced non-finite head values in bf16 storage: the checkpoint or the input is broken")
                log.warning("fp16 activation overflow detected (%d head rows non-finite): switching this model to bf16 storage", eng.health() - h0)
                self.act_dtype = "bf16"
                eng = self._get_engine(frames.shape[1:3], imgsz, device, max_det, nb)
                continue                # same chunk again on the bf16 engine
            st = eng.stage_times()
            speed = {k: st[k] / len(chunk) for k in ("preprocess", "inference", "postprocess")}
            for i in range(len(chunk)):
                rows = torch.from_numpy(boxes[i, : counts[i]].copy())
                kw = {"obb": rows} if self.task == "obb" else {"boxes": rows}
                out.append(Results(chunk[i], path="", names=self.names, speed=dict(speed), **kw))
            b0 += len(chunk)
        return out

    __call__ = predict

    def track(self, source=None, stream: bool = False, persist: bool = False, tracker=None, **kwargs) -> List[Results]:
        """Detect, then hand each image's detections to the host tracker (sequential, CPU) exactly where ultralytics does."""
        from .tracker import make_tracker

        results = self.predict(source, **kwargs)
        if self._tracker is None or not persist:
            self._tracker = make_tracker(tracker)
            self.trackers = [self._tracker]
        for r in results:
            det = (r.obb if self.task == "obb" else r.boxes).cpu().numpy()      # numpy-backed Boxes / OBB: what ultralytics hands its trackers
            if len(det) == 0:
                continue
            tracks = self._tracker.update(det, r.orig_img, None)
            if len(tracks) == 0:
                continue                                         # untracked: boxes.id stays None (extract.py:161-164 writes -1)
            idx = tracks[:, -1].astype(int)
            if self.task == "obb" and tracks.shape[1] == 9:      # ultralytics trackers: [x, y, w, h, angle, id, score, cls, det_idx]
                r.update(obb=tracks[:, :-1].astype(np.float32))
            elif self.task == "obb":                             # stand-in tracker rows are axis-aligned: keep the detector's xywhr
                d = np.asarray(det.data)[idx]
                r.update(obb=np.concatenate([d[:, :5], tracks[:, 4:5], d[:, 5:7]], 1).astype(np.float32))
            else:
                r.update(boxes=tracks[:, :-1].astype(np.float32))
        return results

    # -- misc ultralytics surface --------------------------------------------------------------------------------------------
    def fuse(self):
        return self

    def to(self, *a, **k):
        return self

    def info(self, *a, **k):
        n = sum(int(np.prod(w.shape)) + int(np.prod(b.shape)) for w, b in self._folded.values())
        return len(self._folded), n


class RTDETR(YOLO):
    def __init__(self, *a, **k):
        raise GtError("RT-DETR checkpoints are outside the B200 hot path (YOLOv8s detect / obb only)")
