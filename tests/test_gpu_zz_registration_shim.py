"""Drop-in proof for the registration path, REAL engine: the reference's own, unmodified
`geotrax/utils/registration.py` (from baseline/_ref) runs `from stabilo import Stabilizer` -> the shim's rsift preset -> host SIFT +
`gt_match_l2` + `gt_find_homography`, and returns the same tuple as the product's mirror and the OpenCV restatement.

(Named to run last: the host glue it exercises is covered on the CPU with a stand-in engine in tests/test_host_logic.py, the kernels in
tests/test_gpu_registration.py; this file joins the two on the B200.)"""
import importlib
import logging
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, os.path.dirname(__file__))


def test_unmodified_reference_registration_on_the_b200():
    if not os.path.exists(os.path.join(REF, "geotrax", "utils", "registration.py")):
        pytest.skip("baseline/_ref (pip --target install of the reference) not present")
    import geotrax_b200
    from geotrax_b200 import registration
    from oracle import registration_cv
    from test_gpu_registration import _corner_err, _image_pair
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("ultralytics", "stabilo", "geotrax")}
    geotrax_b200.install_shims(force=True)
    sys.path.insert(0, REF)
    try:
        for k in [k for k in sys.modules if k.split(".")[0] == "geotrax"]:
            del sys.modules[k]
        ref_reg = importlib.import_module("geotrax.utils.registration")          # the reference's own module, byte-for-byte
        assert os.path.abspath(ref_reg.__file__).startswith(os.path.abspath(REF))
        assert ref_reg.Stabilizer is geotrax_b200.Stabilizer
        src, dst, Hgt = _image_pair(seed=4, h=540, w=720)
        kw = dict(max_features=20000)
        H, inl, nm, (n_src, n_dst) = ref_reg.estimate_homography(src, dst, logging.getLogger("ref"), **kw)
        H2, inl2, nm2, (n_src2, n_dst2) = registration.estimate_homography(src, dst, None, **kw)
        Ho, inl_o, nm_o, (ns_o, nd_o) = registration_cv.estimate_homography(src, dst, **kw)
        assert H is not None and H2 is not None and Ho is not None
        assert (nm, n_src, n_dst) == (nm2, n_src2, n_dst2) and abs(inl - inl2) <= 2 and _corner_err(H, H2, *src.shape[:2]) < 0.05   # same kernels, same inputs
        assert (n_src, n_dst) == (ns_o, nd_o) and abs(nm - nm_o) <= max(2, 0.002 * nm_o)
        e_gt, e_or = _corner_err(H, Hgt, *src.shape[:2]), _corner_err(H, Ho, *src.shape[:2])
        print(f"reference registration.py on the shims: {n_src}/{n_dst} key points, {nm} matches, {inl} inliers; {e_gt:.3f} px from truth, {e_or:.3f} px from OpenCV")
        assert e_gt < 0.5 and e_or < 0.5
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split(".")[0] in ("ultralytics", "stabilo", "geotrax")]:
            del sys.modules[k]
        sys.modules.update(saved)
