"""`registration.estimate_homography` with the matching and the robust fit on the B200.

Mirror of /root/reference/geotrax/utils/registration.py:21-95 -- same signature, same return tuple, same halve-and-retry loop.  The
reference delegates to a stabilo ``Stabilizer(detector_name='rsift', matcher_name='bf', filter_type='ratio', mask_use=False,
downsample_ratio=1.0, ref_multiplier=1.0, match_query_frame='current', ...)``: SIFT key points, RootSIFT descriptors, brute-force L2
2-NN + Lowe ratio, ``cv2.findHomography(cur, ref, USAC_MAGSAC, ...)``.  Here

* detection + description stay OpenCV's SIFT on the host (``cv2.SIFT_create(nfeatures, enable_precise_upscale)``; the RootSIFT map
  ``sqrt(d / (sum(d) + eps))`` is applied in numpy) -- a scale-space detector is not part of this library;
* the 2-NN over up to 250,000 x 250,000 descriptors (the reference's default ``max_features``: 8 TFLOP of multiply-adds, minutes on
  the CPU) is ``gt_match_l2``: a tcgen05 GEMM on fp16-rounded operands that keeps four candidates per query, re-ranked exactly in
  fp32 (csrc/match_tc.cu);
* the homography is the library's estimator (``gt_find_homography``: MSAC over 4-point hypotheses + Gauss-Newton polish, the one
  the extract path uses) with ``ransac_epipolar_threshold`` in destination pixels and ``ransac_max_iter`` hypotheses.

Only the reference's own configuration is implemented (rsift / sift + bf + ratio + projective); anything else raises.
"""
from __future__ import annotations

import logging
from typing import Optional

import cv2
import numpy as np

from ._lib import GT_MAX_KP, GtError

_ENGINE = None
_MAX_ITER_CAP = 10000
_MATCH_CAP = 262144   # GT_MATCH_L2_MAX (include/geotrax_b200.h)


def _engine(device: int = 0):
    """A small handle that only serves gt_match_l2 / gt_find_homography (its detector workspaces are for a 256 x 384 frame)."""
    global _ENGINE
    if _ENGINE is None:
        from .engine import Engine
        _ENGINE = Engine(frame_hw=(256, 384), imgsz=192, nc=1, max_batch=16, max_det=16, max_features=500, device=device,
                         ransac_max_iter=_MAX_ITER_CAP)
    return _ENGINE


def _gray(img: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if img.ndim == 3 else img


def root_sift(desc: np.ndarray, eps: float) -> np.ndarray:
    """RootSIFT (Arandjelovic & Zisserman): L1-normalise, element-wise square root."""
    d = desc.astype(np.float32, copy=True)
    d /= (d.sum(axis=1, keepdims=True) + np.float32(eps))
    return np.sqrt(d, out=d)


def detect_and_describe(img: np.ndarray, detector_name: str, max_features: int, precise_upscale: bool, rsift_eps: float, mask: Optional[np.ndarray] = None):
    sift = cv2.SIFT_create(nfeatures=int(max_features), enable_precise_upscale=bool(precise_upscale))
    kps, desc = sift.detectAndCompute(_gray(img), mask)
    if desc is None or len(kps) == 0:
        return np.zeros((0, 2), np.float32), np.zeros((0, 128), np.float32)
    pts = np.array([k.pt for k in kps], np.float32)
    if detector_name == "rsift":
        desc = root_sift(desc, rsift_eps)
    return pts, np.ascontiguousarray(desc, np.float32)


def match_and_fit(pts_src, desc_src, pts_dst, desc_dst, filter_ratio, threshold, max_iter, engine=None, query_is_src: bool = True):
    """GPU half: L2 2-NN, ratio test, robust homography src -> dst.  The query set is the source / current frame and the train set the
    destination / reference (stabilo's ``match_query_frame='current'``, what registration.py:72 passes) unless ``query_is_src`` is False.
    Returns (H or None, inliers, good matches)."""
    eng = engine or _engine()
    if len(desc_src) < 2 or len(desc_dst) < 2:
        return None, 0, 0
    if max(len(desc_src), len(desc_dst)) > _MATCH_CAP:
        raise GtError(f"registration: {max(len(desc_src), len(desc_dst))} descriptors exceed the matcher's capacity of {_MATCH_CAP}")
    idx, dist = eng.match_l2(desc_src, desc_dst) if query_is_src else eng.match_l2(desc_dst, desc_src)
    good = (idx[:, 1] >= 0) & (dist[:, 0].astype(np.float64) < float(filter_ratio) * dist[:, 1].astype(np.float64))
    q = np.nonzero(good)[0]
    cap = eng.max_batch * GT_MAX_KP
    if len(q) > cap:          # more good matches than the pair buffers hold: keep the most distinctive ones
        order = np.argsort(dist[q, 0] / np.maximum(dist[q, 1], 1e-12), kind="stable")[:cap]
        q = np.sort(q[order])
    if len(q) < 4:
        return None, 0, int(len(q))
    i_src, i_dst = (q, idx[q, 0]) if query_is_src else (idx[q, 0], q)
    src = np.ascontiguousarray(pts_src[i_src], np.float32)
    dst = np.ascontiguousarray(pts_dst[i_dst], np.float32)
    H, inl = eng.find_homography(src, dst, float(threshold), int(min(max_iter, _MAX_ITER_CAP)))
    return H, int(inl), int(len(q))


def estimate_homography(img_src: np.ndarray, img_dst: np.ndarray, logger: Optional[logging.Logger] = None, *, detector_name: str = "rsift",
                        matcher_name: str = "bf", filter_type: str = "ratio", sift_enable_precise_upscale: bool = True,
                        max_features: int = 250000, filter_ratio: float = 0.55, ransac_method: int = cv2.USAC_MAGSAC,
                        ransac_epipolar_threshold: float = 3.0, ransac_max_iter: int = 10000, ransac_confidence: float = 0.999999,
                        rsift_eps: float = 1e-8, engine=None) -> tuple:
    """(H src -> dst, inliers, good matches, (n_src_kpts, n_dst_kpts)) or (None, None, None, None) -- registration.py:36-56."""
    logger = logger or logging.getLogger(__name__)
    if detector_name not in ("rsift", "sift") or matcher_name != "bf" or filter_type != "ratio":
        raise NotImplementedError(f"registration on the B200 path supports rsift / sift + bf + ratio, got {detector_name} / {matcher_name} / {filter_type}")
    if ransac_method in (cv2.LMEDS, cv2.RHO):
        raise NotImplementedError("registration: LMEDS / RHO have no counterpart in this library's estimator")
    del ransac_confidence   # every one of the ransac_max_iter hypotheses is scored
    max_features_to_try = int(max_features)
    while max_features_to_try > 10000:
        pts_d, desc_d = detect_and_describe(img_dst, detector_name, max_features_to_try, sift_enable_precise_upscale, rsift_eps)
        pts_s, desc_s = detect_and_describe(img_src, detector_name, max_features_to_try, sift_enable_precise_upscale, rsift_eps)
        H, inl, nm = match_and_fit(pts_s, desc_s, pts_d, desc_d, filter_ratio, ransac_epipolar_threshold, ransac_max_iter, engine)
        if H is not None:
            return H, inl, nm, (len(pts_s), len(pts_d))
        max_features_to_try //= 2
        logger.warning(f"Feature detection or matching failed with {max_features_to_try * 2} max_features. "
                       f"Trying with {max_features_to_try} max_features.")
    logger.error("Feature detection failed with all attempted feature counts.")
    return None, None, None, None
