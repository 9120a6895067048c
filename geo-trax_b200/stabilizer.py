"""``stabilo.Stabilizer``-shaped front end of the CUDA stabilizer (drop-in for /root/reference/geotrax/extract.py:139,
177-187 and /root/reference/geotrax/utils/registration.py:59-85).

    stabilizer = Stabilizer(**config['stabilo'])                 # default.yaml:103-145
    stabilizer.set_ref_frame(frame, boxes_xywh or None)          # first processed frame
    stabilizer.stabilize(frame, boxes_xywh or None)              # every later frame
    stabilizer.transform_cur_boxes() -> (n,4) xywh f32           # 4 corners -> H -> envelope (pinned by the golden files)
    stabilizer.get_cur_trans_matrix() -> (3,3) f64 | None        # current -> reference, full-res px, h33 = 1
    stabilizer.get_cur_num_keypoints() / get_cur_inliers_count() / get_cur_num_matches()

ORB, matching, RANSAC and the box warp run in libgeotrax_b200.so (gt_set_reference / gt_stabilize / gt_warp_boxes).  Poor
matches never raise: the matrix is ``None`` (extract.py:185 then skips the transform row).  Unsupported presets raise
``NotImplementedError`` at construction instead of silently computing something else.

Presets of the reference that run here: ``default.yaml`` and ``stable.yaml`` (CLAHE, full-resolution working image, 4000 / 8000 key
points, ratio 0.8; /root/reference/geotrax/cfg/stable.yaml:115-128), and any ``downsample_ratio`` / ``max_features`` / ``filter_ratio``
/ mask setting; ``detector_name`` sift / rsift (the registration preset: OpenCV's SIFT on the host, brute-force L2 matching and the
robust fit on the GPU).  Not implemented (constructor raises): brisk, kaze, akaze detectors, the FLANN
matcher, filter types other than ``ratio``, affine models, ``ransac_method`` 4 (LMEDS) and 16 (RHO).  ``ransac_method`` 8 and the
USAC family (32..38, default 38 = MAGSAC++) all run the library's own estimator (fixed ``ransac_max_iter`` hypotheses, MSAC score,
Gauss-Newton polish -- DESIGN.md section 4.6); ``ransac_confidence`` only ever shortens OpenCV's iteration count and has no
counterpart here (all ``ransac_max_iter`` hypotheses are always scored).
"""
from __future__ import annotations

import logging
from typing import Optional

import numpy as np

from . import session
from ._lib import GtError

log = logging.getLogger("geotrax_b200")


class Stabilizer:
    def __init__(self, detector_name: str = "orb", matcher_name: str = "bf", filter_type: str = "ratio", transformation_type: str = "projective",
                 clahe: bool = False, downsample_ratio: float = 0.5, max_features: int = 2000, ref_multiplier: float = 2.0,
                 mask_use: bool = True, mask_margin_ratio: float = 0.15, filter_ratio: float = 0.9, ransac_method: int = 38,
                 ransac_epipolar_threshold: float = 2.0, ransac_max_iter: int = 5000, ransac_confidence: float = 0.999999,
                 match_query_frame: str = "current", gpu: bool = False, viz: bool = False, benchmark: bool = False,
                 min_good_match_count_warning: int = 20, min_inliers_match_count_warning: int = 10, device=None, **other):
        # `other` swallows the detector-specific keys of the YAML block that do not apply to ORB
        # (sift_enable_precise_upscale, rsift_eps, brisk_threshold, kaze_threshold, akaze_threshold)
        unsupported = []
        if detector_name not in ("orb", "sift", "rsift"): unsupported.append(f"detector_name={detector_name!r} (orb, sift, rsift)")
        if matcher_name != "bf": unsupported.append(f"matcher_name={matcher_name!r} (bf)")
        if filter_type != "ratio": unsupported.append(f"filter_type={filter_type!r} (ratio)")
        if transformation_type != "projective": unsupported.append(f"transformation_type={transformation_type!r} (projective)")
        if int(ransac_method) not in (8, 32, 33, 34, 35, 36, 37, 38):
            unsupported.append(f"ransac_method={ransac_method} (8 RANSAC or a USAC method 32..38; LMEDS / RHO have no counterpart)")
        if not (0.0 < float(downsample_ratio) <= 1.0): unsupported.append(f"downsample_ratio={downsample_ratio} (0 < r <= 1)")
        if match_query_frame not in ("current", "reference"): unsupported.append(f"match_query_frame={match_query_frame!r}")
        if unsupported:
            raise NotImplementedError("B200 stabilizer implements the ORB (and SIFT / RootSIFT) + BF + ratio + projective pipeline; unsupported: " + ", ".join(unsupported))
        self.cfg = dict(downsample_ratio=float(downsample_ratio), max_features=int(max_features), ref_multiplier=float(ref_multiplier),
                        mask_use=bool(mask_use), mask_margin_ratio=float(mask_margin_ratio), filter_ratio=float(filter_ratio),
                        ransac_epipolar_threshold=float(ransac_epipolar_threshold), ransac_max_iter=int(ransac_max_iter),
                        match_query_frame=match_query_frame, clahe=bool(clahe), ransac_space=str(other.get("ransac_space", "working")))
        if int(ransac_method) != 38:
            log.warning("stabilizer: ransac_method=%d requested; the B200 estimator (MSAC + local optimisation) is used for every supported method", int(ransac_method))
        self.ransac_method, self.ransac_confidence = int(ransac_method), float(ransac_confidence)  # the estimator is the library's own
        self.min_good, self.min_inl = int(min_good_match_count_warning), int(min_inliers_match_count_warning)
        self.device = session.device_index(device)
        if detector_name == "orb":
            session.register_stab_cfg(self.cfg)      # (a detector engine created afterwards carries this ORB / RANSAC setup)
        self._eng = None
        self._H: Optional[np.ndarray] = None
        self._boxes: Optional[np.ndarray] = None
        self._stats = np.zeros(4, np.int32)
        self._have_ref = False
        self._ref_frame: Optional[np.ndarray] = None      # host copy of the reference frame (restores the handle's state, see _engine_for)
        self._ref_boxes: Optional[np.ndarray] = None
        # SIFT / RootSIFT (the registration preset, /root/reference/geotrax/utils/registration.py:59-77): key points and descriptors come from
        # OpenCV's SIFT on the host, the brute-force L2 2-NN and the robust fit run on the GPU (registration.match_and_fit -> gt_match_l2,
        # gt_find_homography).  The ORB path below is untouched by it.
        self.detector_name = detector_name
        self._sift = dict(precise_upscale=bool(other.get("sift_enable_precise_upscale", False)), eps=float(other.get("rsift_eps", 1e-8)))
        self._sift_ref = None                              # (points in working-image pixels, descriptors)

    # -- engine --------------------------------------------------------------------------------------------------------------
    def _engine_for(self, frame: np.ndarray):
        if frame.ndim != 3 or frame.shape[2] != 3 or frame.dtype != np.uint8:
            raise GtError(f"Stabilizer expects a BGR uint8 frame (H,W,3); got {frame.dtype} {frame.shape}")
        hw = tuple(frame.shape[:2])
        eng = self._eng
        if eng is None or (eng.cfg.frame_h, eng.cfg.frame_w) != hw or not eng.h:
            eng = session.find_for_stabilizer(hw, self.device, self.cfg)
            if eng is None:
                eng = session.acquire(hw, None, 4, "detect", self.device, 1, self.cfg)
            self._eng = eng
        # The reference lives in the handle.  If the handle was re-created (a larger batch was requested from the session) or another
        # Stabilizer object set its own reference on the shared handle, this object's reference is restored from its host copy.
        if self._have_ref and getattr(eng, "_ref_owner", None) is not self:
            if self._ref_frame is None or tuple(self._ref_frame.shape[:2]) != hw:
                raise GtError("Stabilizer: frame size changed after set_ref_frame()")
            eng.preprocess(self._ref_frame[None])
            eng._frame_token = None
            eng.set_reference(0, self._ref_boxes)
            eng._ref_owner = self
        return eng

    def _upload(self, eng, frame: np.ndarray):
        tok = session.frame_token(frame)
        if getattr(eng, "_frame_token", None) != tok:      # the detector shim has not just pre-processed this very frame
            eng.preprocess(np.ascontiguousarray(frame)[None])
            eng._frame_token = tok

    @staticmethod
    def _clean_boxes(boxes) -> Optional[np.ndarray]:
        if boxes is None:
            return None
        b = np.asarray(boxes, np.float32).reshape(-1, 4)
        return b if len(b) else None

    # -- SIFT / RootSIFT path (host detector, GPU matcher + robust fit) -------------------------------------------------------------
    def _sift_features(self, frame: np.ndarray, boxes, n_features: int):
        import cv2
        from . import registration
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 else frame
        r = self.cfg["downsample_ratio"]
        if r != 1.0:
            gray = cv2.resize(gray, (int(round(gray.shape[1] * r)), int(round(gray.shape[0] * r))), interpolation=cv2.INTER_LINEAR)
        if self.cfg["clahe"]:
            gray = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(gray)
        mask = None
        b = self._clean_boxes(boxes)
        if self.cfg["mask_use"] and b is not None:           # the vehicle rectangles (grown by the margin) are excluded, as in the ORB path
            mask = np.full(gray.shape, 255, np.uint8)
            m = 1.0 + self.cfg["mask_margin_ratio"]
            for x, y, w, h in b.astype(np.float64):
                x0, y0 = int(np.floor((x - w * m / 2) * r)), int(np.floor((y - h * m / 2) * r))
                x1, y1 = int(np.ceil((x + w * m / 2) * r)), int(np.ceil((y + h * m / 2) * r))
                mask[max(y0, 0):max(y1, 0), max(x0, 0):max(x1, 0)] = 0
        return registration.detect_and_describe(gray, self.detector_name, n_features, self._sift["precise_upscale"], self._sift["eps"], mask), b

    def _sift_set_ref(self, frame: np.ndarray, boxes) -> None:
        n_ref = int(round(self.cfg["max_features"] * self.cfg["ref_multiplier"]))
        self._sift_ref, b = self._sift_features(frame, boxes, n_ref)
        self._have_ref = True
        self._H, self._boxes = None, b
        self._stats[:] = 0

    def _sift_stabilize(self, frame: np.ndarray, boxes) -> None:
        from . import registration
        if self._sift_ref is None:
            raise GtError("Stabilizer.stabilize() called before set_ref_frame()")
        (pts_c, desc_c), b = self._sift_features(frame, boxes, self.cfg["max_features"])
        pts_r, desc_r = self._sift_ref
        eng = registration._engine(self.device)
        self._eng = eng
        H, inl, nm = registration.match_and_fit(pts_c, desc_c, pts_r, desc_r, self.cfg["filter_ratio"], self.cfg["ransac_epipolar_threshold"],
                                                self.cfg["ransac_max_iter"], engine=eng, query_is_src=self.cfg["match_query_frame"] == "current")
        r = self.cfg["downsample_ratio"]
        if H is not None and r != 1.0:                        # working-image pixels -> source-frame pixels: S^-1 H S, S = diag(r, r, 1)
            H = H.copy()
            H[0, 2] /= r; H[1, 2] /= r; H[2, 0] *= r; H[2, 1] *= r
        self._boxes = b
        self._stats = np.array([len(pts_r), len(pts_c), nm, inl], np.int32)
        self._H = H
        if nm < self.min_good:
            log.warning("stabilizer: only %d good matches", int(nm))
        if H is not None and inl < self.min_inl:
            log.warning("stabilizer: only %d RANSAC inliers", int(inl))

    # -- stabilo surface -----------------------------------------------------------------------------------------------------
    def set_ref_frame(self, frame: np.ndarray, boxes=None) -> None:
        if self.detector_name != "orb":
            return self._sift_set_ref(frame, boxes)
        self._have_ref = False
        eng = self._engine_for(frame)
        self._upload(eng, frame)
        b = self._clean_boxes(boxes)
        eng.set_reference(0, b)
        eng._ref_owner = self
        self._have_ref = True
        self._ref_frame, self._ref_boxes = np.ascontiguousarray(frame).copy(), (None if b is None else b.copy())
        self._H, self._boxes = None, b
        self._stats[:] = 0

    def stabilize(self, frame: np.ndarray, boxes=None) -> None:
        if self.detector_name != "orb":
            return self._sift_stabilize(frame, boxes)
        eng = self._engine_for(frame)
        if not self._have_ref:
            raise GtError("Stabilizer.stabilize() called before set_ref_frame()")
        self._upload(eng, frame)
        b = self._clean_boxes(boxes)
        H, status, stats = eng.stabilize(1, [b])
        self._boxes = b
        self._stats = stats[0].copy()
        self._H = H[0].copy() if int(status[0]) == 0 else None
        if self._stats[2] < self.min_good:
            log.warning("stabilizer: only %d good matches", int(self._stats[2]))
        if self._H is not None and self._stats[3] < self.min_inl:
            log.warning("stabilizer: only %d RANSAC inliers", int(self._stats[3]))

    def get_cur_trans_matrix(self) -> Optional[np.ndarray]:
        return self._H

    def transform_cur_boxes(self) -> Optional[np.ndarray]:
        if self._boxes is None:
            return None
        if self._H is None:
            return self._boxes.copy()
        out = np.empty_like(self._boxes)
        md = self._eng.max_det
        for i in range(0, len(self._boxes), md):
            out[i:i + md] = self._eng.warp_boxes(self._H, self._boxes[i:i + md])
        return out

    def get_cur_num_keypoints(self):
        return int(self._stats[0]), int(self._stats[1])     # (reference, current)

    def get_cur_inliers_count(self) -> int:
        return int(self._stats[3])

    def get_cur_num_matches(self) -> int:
        return int(self._stats[2])
