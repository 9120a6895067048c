"""CPU checks of the registration oracle (oracle/registration_cv.py) and of the host half of the product's mirror."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(__file__))


def _pair(seed=2):
    from test_gpu_registration import _corner_err, _image_pair
    return _image_pair(seed), _corner_err


def test_oracle_registration_recovers_ground_truth():
    from oracle import registration_cv
    (src, dst, Hgt), corner_err = _pair()
    H, inl, nm, (ns, nd) = registration_cv.estimate_homography(src, dst, max_features=20000)
    assert H is not None and nm >= 200 and inl >= 0.8 * nm and ns > 1000 and nd > 1000
    assert corner_err(H, Hgt, *src.shape[:2]) < 0.3
    flat = np.full((300, 400, 3), 127, np.uint8)
    assert registration_cv.estimate_homography(flat, flat, max_features=20000) == (None, None, None, None)


def test_root_sift_is_the_same_map_on_both_sides():
    """The product's host-side RootSIFT and the oracle's are the same arithmetic (unit L2 norm, bitwise equal)."""
    from geotrax_b200 import registration
    from oracle import registration_cv
    rng = np.random.default_rng(0)
    d = np.floor(rng.gamma(0.6, 18.0, (64, 128))).astype(np.float32)
    a, b = registration.root_sift(d, 1e-8), registration_cv.root_sift(d, 1e-8)
    assert np.array_equal(a, b)
    nz = d.sum(1) > 0
    assert np.allclose(np.linalg.norm(a[nz], axis=1), 1.0, atol=1e-6)


def test_registration_signature_matches_the_reference():
    """Keyword names and defaults of /root/reference/geotrax/utils/registration.py:21-37 (checked against the source when it is present)."""
    import inspect
    from geotrax_b200 import registration
    sig = inspect.signature(registration.estimate_homography)
    want = dict(detector_name="rsift", matcher_name="bf", filter_type="ratio", sift_enable_precise_upscale=True, max_features=250000,
                filter_ratio=0.55, ransac_method=cv2.USAC_MAGSAC, ransac_epipolar_threshold=3.0, ransac_max_iter=10000,
                ransac_confidence=0.999999, rsift_eps=1e-8)
    for k, v in want.items():
        assert sig.parameters[k].default == v and sig.parameters[k].kind is inspect.Parameter.KEYWORD_ONLY, k
    assert list(sig.parameters)[:3] == ["img_src", "img_dst", "logger"]
    ref = "/root/reference/geotrax/utils/registration.py"
    if os.path.exists(ref):
        text = open(ref).read()
        for k, v in want.items():
            assert f"{k}: " in text, k
            if k not in ("ransac_method", "rsift_eps"):
                assert f"= {v!r}" in text, k
