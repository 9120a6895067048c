"""GPU parity of the registration path: brute-force L2 2-NN on the tensor cores and `estimate_homography`
against the OpenCV restatement (oracle/registration_cv.py) of /root/reference/geotrax/utils/registration.py:57-93."""
import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reg_engine():
    import geotrax_b200
    eng = geotrax_b200.Engine(frame_hw=(256, 384), imgsz=192, nc=1, max_batch=16, max_det=16, max_features=500, ransac_max_iter=10000)
    yield eng
    eng.close()


def _rootsift_like(rng, n):
    """Descriptors with SIFT statistics: sparse non-negative integer histograms, RootSIFT-mapped (unit L2 norm)."""
    d = rng.gamma(0.6, 18.0, (n, 128)).astype(np.float32)
    d[rng.random((n, 128)) < 0.35] = 0
    d = np.minimum(np.floor(d), 255).astype(np.float32)
    d[:, 0] += 1
    d /= d.sum(1, keepdims=True) + 1e-8
    return np.sqrt(d)


def _check_against_bf(eng, query, train, tol=2e-6):
    from oracle import registration_cv
    idx, dist = eng.match_l2(query, train)
    ri, rd = registration_cv.knn_l2(query, train)
    k = min(2, len(train))
    # distances: fp32 either way, different summation order
    assert np.allclose(dist[:, :k], rd[:, :k], rtol=1e-5, atol=1e-6), np.abs(dist[:, :k] - rd[:, :k]).max()
    if k < 2:
        assert (idx[:, 1] == -1).all() and (ri[:, 1] == -1).all()
    # indices: identical wherever OpenCV's own decision has a margin (nearest vs second, second vs third nearest)
    d2 = ((query[:, None, :].astype(np.float64) - train[None, :, :].astype(np.float64)) ** 2).sum(-1) if len(query) * len(train) <= 4_000_000 else None
    same = (idx[:, :k] == ri[:, :k]).all(1)
    if d2 is not None and len(train) >= 3:
        s = np.sqrt(np.sort(d2, axis=1)[:, :3])
        margin = np.minimum(s[:, 1] - s[:, 0], s[:, 2] - s[:, 1])
        assert same[margin > tol].all(), f"{(~same[margin > tol]).sum()} queries with a clear margin differ"
        assert (margin > tol).mean() > 0.95
    else:
        assert same.mean() > 0.999, f"only {same.mean():.4f} of the index pairs agree"
    return idx, dist


@pytest.mark.parametrize("nq,nt", [(1, 1), (3, 2), (130, 257), (257, 129), (1000, 1777), (5000, 4100)])
def test_match_l2_equals_bfmatcher(reg_engine, nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt)
    train = _rootsift_like(rng, nt)
    query = _rootsift_like(rng, nq)
    m = min(nq, nt) // 2                       # half of the queries are noisy copies of train rows (true matches), as in registration
    if m:
        src = rng.choice(nt, m, replace=False)
        noisy = np.maximum(train[src] + rng.normal(0, 0.01, (m, 128)).astype(np.float32), 0)
        query[:m] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
    _check_against_bf(reg_engine, query, train)


def test_match_l2_ties_and_plain_sift_range(reg_engine):
    """Exact duplicates in the train set (equal distances -> the lower index first, as BFMatcher) and un-normalised 0..255 SIFT values."""
    rng = np.random.default_rng(3)
    train = np.minimum(np.floor(rng.gamma(0.6, 18.0, (600, 128))), 255).astype(np.float32)
    train[400] = train[17]; train[590] = train[17]
    query = train[rng.integers(0, 600, 300)].copy()
    query[0] = train[17]
    query[1:] = np.maximum(query[1:] + np.round(rng.normal(0, 2.0, (299, 128))).astype(np.float32), 0)
    idx, dist = reg_engine.match_l2(query, train)
    assert idx[0].tolist() == [17, 400] and dist[0].tolist() == [0.0, 0.0]
    from oracle import registration_cv
    ri, rd = registration_cv.knn_l2(query, train)
    assert np.array_equal(idx[:, 0], ri[:, 0])
    assert np.allclose(dist, rd, rtol=1e-5, atol=1e-4)


def test_match_l2_large(reg_engine):
    """40,000 x 50,000 descriptors (the buffers grow on demand); checked against BFMatcher on a slice of the queries."""
    from oracle import registration_cv
    rng = np.random.default_rng(9)
    train = _rootsift_like(rng, 50_000)
    query = _rootsift_like(rng, 40_000)
    src = rng.choice(50_000, 20_000, replace=False)
    noisy = np.maximum(train[src] + rng.normal(0, 0.01, (20_000, 128)).astype(np.float32), 0)
    query[:20_000] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
    idx, dist = reg_engine.match_l2(query, train)
    assert (idx[:20_000, 0] == src).mean() > 0.999
    sl = np.r_[0:400, 39_600:40_000]
    ri, rd = registration_cv.knn_l2(query[sl], train)
    assert np.allclose(dist[sl], rd, rtol=1e-5, atol=1e-6)
    assert (idx[sl] == ri).all(1).mean() > 0.99


def _image_pair(seed=2, h=720, w=1080):
    from geotrax_b200 import synth
    rng = np.random.default_rng(seed)
    frames, _, _ = synth.make_flight(1, h, w, seed=seed, n_vehicles=0)
    dst = frames[0]
    H = synth.small_homography(rng, h, w, max_t=25.0, max_rot_deg=2.0, max_persp=2e-5)    # maps dst -> src pixel coordinates ...
    src = cv2.warpPerspective(dst, H, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    return src, dst, np.linalg.inv(H)                                                       # ... so src -> dst is its inverse


def _corner_err(Ha, Hb, h, w):
    p = np.array([[0, 0, 1], [w, 0, 1], [w, h, 1], [0, h, 1], [w / 2, h / 2, 1.0]])
    a, b = p @ Ha.T, p @ Hb.T
    return np.linalg.norm(a[:, :2] / a[:, 2:] - b[:, :2] / b[:, 2:], axis=1).max()


def test_estimate_homography_matches_oracle_and_truth(reg_engine):
    from geotrax_b200 import registration
    from oracle import registration_cv
    src, dst, Hgt = _image_pair()
    kw = dict(max_features=20000, filter_ratio=0.55, ransac_epipolar_threshold=3.0, ransac_max_iter=10000)
    H, inl, nm, (ns, nd) = registration.estimate_homography(src, dst, None, engine=reg_engine, **kw)
    Ho, inl_o, nm_o, (ns_o, nd_o) = registration_cv.estimate_homography(src, dst, **kw)
    assert H is not None and Ho is not None
    assert (ns, nd) == (ns_o, nd_o)                       # same detector on the host
    assert abs(nm - nm_o) <= max(2, 0.002 * nm_o)         # ratio test on (almost) bit-identical distances
    assert abs(inl - inl_o) <= 0.03 * nm_o
    e_gt, e_or, e_o_gt = _corner_err(H, Hgt, *src.shape[:2]), _corner_err(H, Ho, *src.shape[:2]), _corner_err(Ho, Hgt, *src.shape[:2])
    print(f"registration: {ns}/{nd} key points, {nm} matches, {inl} inliers; corner error vs truth {e_gt:.3f} px, vs OpenCV {e_or:.3f} px (OpenCV vs truth {e_o_gt:.3f})")
    assert e_gt < 0.5 and e_or < 0.5


def test_estimate_homography_failure_returns_nones(reg_engine):
    from geotrax_b200 import registration
    flat = np.full((300, 400, 3), 127, np.uint8)
    assert registration.estimate_homography(flat, flat, None, engine=reg_engine, max_features=20000) == (None, None, None, None)
    with pytest.raises(NotImplementedError):
        registration.estimate_homography(flat, flat, None, engine=reg_engine, detector_name="orb")


def test_find_homography_on_more_pairs_than_one_frame_holds(reg_engine):
    """A single pair set may use the pair buffers of the whole batch (max_batch x 8192): 40,000 matches with 30 % outliers, 10,000 hypotheses."""
    from geotrax_b200 import synth
    rng = np.random.default_rng(5)
    Hgt = synth.small_homography(rng, 2160, 3840, max_t=40.0, max_rot_deg=2.0, max_persp=1e-5)
    n = 40_000
    src = np.stack([rng.uniform(0, 3840, n), rng.uniform(0, 2160, n)], 1)
    p = np.c_[src, np.ones(n)] @ Hgt.T
    dst = p[:, :2] / p[:, 2:] + rng.normal(0, 0.4, (n, 2))
    out = rng.random(n) < 0.30
    dst[out] = np.stack([rng.uniform(0, 3840, out.sum()), rng.uniform(0, 2160, out.sum())], 1)
    H, inl = reg_engine.find_homography(src.astype(np.float32), dst.astype(np.float32), 3.0, 10000)
    assert H is not None
    assert _corner_err(H, Hgt, 2160, 3840) < 0.1
    assert abs(inl - int((~out).sum())) < 0.02 * n
    with pytest.raises(Exception):
        reg_engine.find_homography(np.zeros((16 * 8192 + 1, 2), np.float32), np.zeros((16 * 8192 + 1, 2), np.float32))
