"""OpenCV restatement of ``stabilo.Stabilizer`` for the ORB / BF / ratio / projective preset (oracle; test infra only).

``stabilo>=1.2.3`` (/root/reference/pyproject.toml:59) is not vendored; this follows its use at
/root/reference/geotrax/extract.py:139,177-187 and /root/reference/geotrax/utils/registration.py:57-93, the parameter
block /root/reference/geotrax/cfg/default.yaml:103-145, the same-author mask-margin arithmetic at
/root/reference/tools/annotate_frames.py:255-264, and SURVEY.md Appendix A-3.  It *calls* OpenCV (cv2.ORB, BFMatcher,
findHomography USAC_MAGSAC) because that is the engine stabilo itself calls.

Pinned by the reference's golden files (tests/test_oracle_golden.py): ``transform_cur_boxes`` semantics (AABB of the 4
warped corners), H direction (current -> reference), units (full-resolution pixels), h33 = 1, row-major flatten.
Unpinned ([U] in SURVEY.md): mask rounding, query/train order default, where the downsample rescale is applied.  Those
are constructor switches here (``match_query_frame``, ``ransac_space``) so the CUDA path can be tested under either.
"""
from __future__ import annotations

from typing import Optional, Tuple

import cv2
import numpy as np


def mask_rects(boxes_xywh: np.ndarray, margin: float, ratio: float, w: int, h: int) -> np.ndarray:
    """Integer rectangles [x0,y0,x1,y1) at working resolution that are zeroed in the mask.

    Each box is grown by ``margin`` (w*(1+m), h*(1+m)), scaled by ``ratio`` and truncated with int() like
    tools/annotate_frames.py:258-261; then clipped to the image.
    """
    out = np.zeros((len(boxes_xywh), 4), dtype=np.int32)
    for i, (xc, yc, bw, bh) in enumerate(np.asarray(boxes_xywh, dtype=np.float64)):
        gw, gh = bw * (1.0 + margin), bh * (1.0 + margin)
        x0 = int(np.floor((xc - gw / 2) * ratio))
        y0 = int(np.floor((yc - gh / 2) * ratio))
        x1 = int(np.ceil((xc + gw / 2) * ratio))
        y1 = int(np.ceil((yc + gh / 2) * ratio))
        out[i] = (min(max(x0, 0), w), min(max(y0, 0), h), min(max(x1, 0), w), min(max(y1, 0), h))
    return out


def build_mask(boxes_xywh: Optional[np.ndarray], margin: float, ratio: float, w: int, h: int) -> np.ndarray:
    m = np.full((h, w), 255, dtype=np.uint8)
    if boxes_xywh is not None and len(boxes_xywh):
        for x0, y0, x1, y1 in mask_rects(boxes_xywh, margin, ratio, w, h):
            m[y0:y1, x0:x1] = 0
    return m


def to_gray_half(frame_bgr: np.ndarray, ratio: float) -> np.ndarray:
    g = cv2.cvtColor(frame_bgr, cv2.COLOR_BGR2GRAY)
    if ratio != 1.0:
        g = cv2.resize(g, (int(g.shape[1] * ratio), int(g.shape[0] * ratio)), interpolation=cv2.INTER_LINEAR)
    return g


def warp_boxes_xywh(boxes_xywh: np.ndarray, H: np.ndarray) -> np.ndarray:
    """4 corners -> perspectiveTransform -> axis-aligned envelope -> xywh (pinned by the golden files)."""
    b = np.asarray(boxes_xywh, dtype=np.float64).reshape(-1, 4)
    if len(b) == 0:
        return np.zeros((0, 4), dtype=np.float32)
    x0, y0, x1, y1 = b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2
    cx = np.stack([x0, x1, x1, x0], 1)
    cy = np.stack([y0, y0, y1, y1], 1)
    w = H[2, 0] * cx + H[2, 1] * cy + H[2, 2]
    u = (H[0, 0] * cx + H[0, 1] * cy + H[0, 2]) / w
    v = (H[1, 0] * cx + H[1, 1] * cy + H[1, 2]) / w
    ux0, ux1, vy0, vy1 = u.min(1), u.max(1), v.min(1), v.max(1)
    return np.stack([(ux0 + ux1) / 2, (vy0 + vy1) / 2, ux1 - ux0, vy1 - vy0], 1).astype(np.float32)


class Stabilizer:
    """ORB + BF(kNN 2, Hamming) + Lowe ratio + cv2.findHomography, frame-to-reference."""

    def __init__(self, detector_name="orb", matcher_name="bf", filter_type="ratio", transformation_type="projective",
                 clahe=False, downsample_ratio=0.5, max_features=2000, ref_multiplier=2.0, mask_use=True,
                 mask_margin_ratio=0.15, filter_ratio=0.9, ransac_method=38, ransac_epipolar_threshold=2.0,
                 ransac_max_iter=5000, ransac_confidence=0.999999, match_query_frame="current",
                 ransac_space="working", **_ignored):
        assert detector_name == "orb" and matcher_name == "bf" and transformation_type == "projective"
        self.ratio = float(downsample_ratio)
        self.max_features, self.ref_multiplier = int(max_features), float(ref_multiplier)
        self.mask_use, self.margin = bool(mask_use), float(mask_margin_ratio)
        self.filter_type, self.filter_ratio = filter_type, float(filter_ratio)
        self.method, self.thr = int(ransac_method), float(ransac_epipolar_threshold)
        self.max_iter, self.confidence = int(ransac_max_iter), float(ransac_confidence)
        self.clahe = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)) if clahe else None
        self.query = match_query_frame
        self.ransac_space = ransac_space
        self.ref = None
        self.cur = None
        self.H = None
        self.cur_boxes = None
        self.n_matches = 0
        self.n_inliers = 0

    # -- per-frame front end ------------------------------------------------------------------------------------
    def _process(self, frame, boxes, is_ref):
        g = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 else frame
        if self.clahe is not None:
            g = self.clahe.apply(g)
        if self.ratio != 1.0:
            g = cv2.resize(g, (int(g.shape[1] * self.ratio), int(g.shape[0] * self.ratio)), interpolation=cv2.INTER_LINEAR)
        mask = None
        if self.mask_use and boxes is not None and len(boxes):
            mask = build_mask(boxes, self.margin, self.ratio, g.shape[1], g.shape[0])
        n = int(self.max_features * (self.ref_multiplier if is_ref else 1.0))
        kps, desc = cv2.ORB_create(nfeatures=n).detectAndCompute(g, mask)
        pts = np.array([k.pt for k in kps], dtype=np.float32).reshape(-1, 2)
        return dict(gray=g, mask=mask, kps=kps, pts=pts, desc=desc)

    def set_ref_frame(self, frame, boxes=None):
        self.ref = self._process(frame, boxes, True)
        self.cur, self.H = None, None
        self.cur_boxes = None if boxes is None else np.asarray(boxes, dtype=np.float32)

    def stabilize(self, frame, boxes=None):
        self.cur = self._process(frame, boxes, False)
        self.cur_boxes = None if boxes is None else np.asarray(boxes, dtype=np.float32)
        self.H, self.n_matches, self.n_inliers = self._estimate()

    # -- matching + robust fit ------------------------------------------------------------------------------------
    def match(self) -> Tuple[np.ndarray, np.ndarray]:
        """-> (cur_idx, ref_idx) of ratio-filtered matches."""
        dc, dr = self.cur["desc"], self.ref["desc"]
        if dc is None or dr is None or len(dc) < 2 or len(dr) < 2:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        q, t = (dc, dr) if self.query == "current" else (dr, dc)
        qi, ti = [], []
        for pair in bf.knnMatch(q, t, k=2):
            if len(pair) < 2:
                continue
            m, n = pair
            if self.filter_type != "ratio" or m.distance < self.filter_ratio * n.distance:
                qi.append(m.queryIdx)
                ti.append(m.trainIdx)
        qi, ti = np.asarray(qi, np.int64), np.asarray(ti, np.int64)
        return (qi, ti) if self.query == "current" else (ti, qi)

    def _estimate(self):
        ci, ri = self.match()
        if len(ci) < 4:
            return None, len(ci), 0
        pc, pr = self.cur["pts"][ci], self.ref["pts"][ri]
        if self.ransac_space == "full":
            pc, pr = pc / self.ratio, pr / self.ratio
        H, inl = cv2.findHomography(pc, pr, self.method, self.thr, maxIters=self.max_iter, confidence=self.confidence)
        if H is None:
            return None, len(ci), 0
        if self.ransac_space != "full" and self.ratio != 1.0:
            S = np.diag([self.ratio, self.ratio, 1.0])
            H = np.linalg.inv(S) @ H @ S
            H = H / H[2, 2]
        return H, len(ci), int(inl.sum())

    # -- getters (stabilo surface) ------------------------------------------------------------------------------
    def get_cur_trans_matrix(self):
        return self.H

    def transform_cur_boxes(self):
        if self.cur_boxes is None:
            return None
        if self.H is None:
            return self.cur_boxes.copy()
        return warp_boxes_xywh(self.cur_boxes, self.H)

    def get_cur_num_keypoints(self):
        return (0 if self.ref is None else len(self.ref["kps"]), 0 if self.cur is None else len(self.cur["kps"]))

    def get_cur_inliers_count(self):
        return self.n_inliers

    def get_cur_num_matches(self):
        return self.n_matches
