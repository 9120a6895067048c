"""Pins the CPU oracle against the reference's own golden output files.

The reference's tests hold no vectors for the hot path; what it ships is the README-documented output of
``geotrax batch data/U_video_cut.mp4 --no-geo`` (/root/reference/data/results-pixel/U_video_cut{,_vid_transf}.txt).
tests/golden/make_golden.py copies the needed rows into u_video_cut_golden.npz.  These pin: box-warp semantics, H direction /
units / normalisation / layout, frame numbering, and the output column layout.  CPU only.
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "u_video_cut_golden.npz")


@pytest.fixture(scope="module")
def golden():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


def test_golden_layout(golden):
    t, tr = golden["tracks"], golden["transforms"]
    assert t.shape[1] == 14 and tr.shape == (149, 10)           # frame,id,x,y,w,h,xs,ys,ws,hs,cls,conf,len,wid | frame,H
    assert int(golden["n_rows_total"]) == 19817
    assert tr[0, 0] == 1 and tr[-1, 0] == 149                   # first transform row is cut_frame_left + 1 (README.md:328)
    assert np.array_equal(tr[:, 0], np.arange(1, 150))
    assert t[:, 1].min() >= 1                                   # ids start at 1, untracked (-1) rows dropped (extract.py:287)
    assert set(np.unique(t[:, 10]).astype(int)) <= {0, 1, 2, 3}
    assert t[:, 11].min() > 0.25                                # conf threshold of the default preset
    d = golden["dets_per_frame"]
    assert d.min() >= 120 and d.max() <= 140                    # workload statistic the synthetic generator mimics


def test_golden_homography_conventions(golden):
    tr = golden["transforms"]
    H = tr[:, 1:].reshape(-1, 3, 3)
    assert np.allclose(H[:, 2, 2], 1.0, atol=1e-12)             # h33 = 1
    assert (np.linalg.det(H) > 0).all()
    assert np.abs(H[:, 2, :2]).max() < 1e-6                     # perspective terms tiny (drone hover)
    assert np.abs(H[:, 0, 2]).max() < 10 and np.abs(H[:, 1, 2]).max() < 10   # a few px of drift at 4K => full-resolution units


def test_reference_frame_rows_unstabilised(golden):
    t = golden["tracks"]
    r0 = t[t[:, 0] == 0]
    assert len(r0) > 100
    assert np.array_equal(r0[:, 2:6], r0[:, 6:10])              # extract.py:178-179: bbox_stab = bbox on the reference frame


def test_oracle_box_warp_reproduces_golden(golden):
    """transform_cur_boxes == axis-aligned envelope of the 4 warped corners, H maps current -> reference."""
    from oracle.stabilo_cv import warp_boxes_xywh
    t, tr = golden["tracks"], golden["transforms"]
    Hby = {int(r[0]): r[1:].reshape(3, 3) for r in tr}
    n = 0
    for f in np.unique(t[:, 0]).astype(int):
        if f == 0:
            continue
        rows = t[t[:, 0] == f]
        out = warp_boxes_xywh(rows[:, 2:6], Hby[f])
        assert np.abs(out[:, :2] - rows[:, 6:8]).max() < 2.5e-2   # '%g' prints 6 significant digits (xxxx.xx)
        assert np.abs(out[:, 2:] - rows[:, 8:10]).max() < 5e-3
        # the inverse direction must NOT fit (guards the direction convention) once the drift is visible
        if np.abs(Hby[f][:2, 2]).max() > 1.0:
            inv = warp_boxes_xywh(rows[:, 2:6], np.linalg.inv(Hby[f]))
            assert np.abs(inv[:, :2] - rows[:, 6:8]).max() > 0.5
        n += len(rows)
    assert n > 1000


def test_centre_only_warp_does_not_fit(golden):
    """The alternative semantics (warp the centre, keep w/h) is rejected by the golden w/h columns."""
    t, tr = golden["tracks"], golden["transforms"]
    rows = t[t[:, 0] == 149]
    assert np.abs(rows[:, 4:6] - rows[:, 8:10]).max() > 5e-3
