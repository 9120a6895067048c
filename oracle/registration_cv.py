"""TEST INFRASTRUCTURE -- CPU restatement of `registration.estimate_homography`.  Never imported by the product.

Follows /root/reference/geotrax/utils/registration.py:57-93: a stabilo Stabilizer with detector 'rsift', matcher 'bf', filter
'ratio', projective model, no mask, no downsampling, reference multiplier 1.0, query = current frame; destination image = reference
frame, source image = current frame, so H maps src -> dst.  stabilo itself is not installable here (parity unpinned for its glue):
the RootSIFT map, the SIFT constructor arguments and the match direction are the standard ones and are marked [U].
"""
import cv2
import numpy as np


def _gray(img):
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if img.ndim == 3 else img


def root_sift(desc, eps=1e-8):
    d = desc.astype(np.float32, copy=True)      # [U] stabilo: desc /= (desc.sum(axis=1, keepdims=True) + eps); desc = np.sqrt(desc)
    d /= (d.sum(axis=1, keepdims=True) + np.float32(eps))
    return np.sqrt(d)


def detect(img, max_features, detector_name="rsift", precise_upscale=True, eps=1e-8):
    sift = cv2.SIFT_create(nfeatures=int(max_features), enable_precise_upscale=bool(precise_upscale))   # [U]
    kps, desc = sift.detectAndCompute(_gray(img), None)
    pts = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    if detector_name == "rsift" and desc is not None:
        desc = root_sift(desc, eps)
    return pts, desc


def knn_l2(query, train):
    """cv2.BFMatcher(NORM_L2).knnMatch(query, train, k=2) as arrays (idx [nq, 2], dist [nq, 2]; -1 where missing)."""
    ms = cv2.BFMatcher(cv2.NORM_L2).knnMatch(np.ascontiguousarray(query, np.float32), np.ascontiguousarray(train, np.float32), k=2)
    idx = -np.ones((len(query), 2), np.int32)
    dist = -np.ones((len(query), 2), np.float32)
    for q, pair in enumerate(ms):
        for k, m in enumerate(pair):
            idx[q, k], dist[q, k] = m.trainIdx, m.distance
    return idx, dist


def estimate_homography(img_src, img_dst, *, detector_name="rsift", max_features=250000, filter_ratio=0.55, ransac_method=cv2.USAC_MAGSAC,
                        ransac_epipolar_threshold=3.0, ransac_max_iter=10000, ransac_confidence=0.999999, rsift_eps=1e-8,
                        sift_enable_precise_upscale=True):
    """registration.py:57-93 on OpenCV alone.  Returns (H, inliers, good matches, (n_src, n_dst)) or four Nones."""
    n = int(max_features)
    while n > 10000:
        pd, dd = detect(img_dst, n, detector_name, sift_enable_precise_upscale, rsift_eps)
        ps, ds = detect(img_src, n, detector_name, sift_enable_precise_upscale, rsift_eps)
        H = None
        if ds is not None and dd is not None and len(ds) >= 2 and len(dd) >= 2:
            idx, dist = knn_l2(ds, dd)
            good = (idx[:, 1] >= 0) & (dist[:, 0].astype(np.float64) < filter_ratio * dist[:, 1].astype(np.float64))
            q = np.nonzero(good)[0]
            if len(q) >= 4:
                H, mask = cv2.findHomography(ps[q], pd[idx[q, 0]], ransac_method, ransac_epipolar_threshold, maxIters=int(ransac_max_iter),
                                             confidence=ransac_confidence)
                if H is not None:
                    return H, int(mask.sum()), int(len(q)), (len(ps), len(pd))
        n //= 2
    return None, None, None, None
