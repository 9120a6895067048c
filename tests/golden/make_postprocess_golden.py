"""Generates tests/golden/postprocess_golden.npz from the REFERENCE's own functions (run in the build container only; /root/reference does
not exist on the GPU box and no test reads it):

    python tests/golden/make_postprocess_golden.py

`geotrax.extract` imports `ultralytics` / `stabilo`, which are not installable here, so it is imported on top of geotrax_b200.install_shims()
(class definitions only -- no GPU is touched).  `get_video_dimensions` opens the source video; there is none, so it is replaced by a
constant 3840 x 2160 while the golden outputs are produced.  Inputs are seeded synthetic track tables in the layout
`aggregate_results` builds (/root/reference/geotrax/extract.py:273-293): 12 columns with stabilisation, 8 without."""
import logging
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def synth_tracks(seed: int, n_tracks: int, n_frames: int, stabilised: bool, w=3840, h=2160) -> np.ndarray:
    """A flight-like table: vehicles entering / leaving, cardinal / diagonal / stationary motion, frame gaps, short tracks, class flicker,
    boxes near the frame border, ids that are neither contiguous nor sorted."""
    rng = np.random.default_rng(seed)
    rows = []
    ids = rng.permutation(np.arange(1, 4 * n_tracks))[:n_tracks]
    for k, tid in enumerate(ids):
        f0 = int(rng.integers(0, n_frames - 2))
        length = int(rng.choice([1, 2, 3, 5, 40, 200, n_frames]))
        frames = np.arange(f0, min(f0 + length, n_frames))
        keep = rng.random(len(frames)) > rng.choice([0.0, 0.05, 0.3])          # detection drop-outs -> gaps (some longer than track_buffer)
        if len(frames) > 80 and rng.random() < 0.3:
            keep[20:20 + int(rng.integers(25, 60))] = False
        keep[0] = True
        frames = frames[keep]
        mode = k % 5
        speed = [0.0, 9.0, 14.0, 6.0, 0.4][mode]
        ang = [0.0, 0.0, np.pi / 2, np.pi / 4, 0.3][mode] + rng.normal(0, 0.05)
        x0, y0 = rng.uniform(-50, w + 50), rng.uniform(-50, h + 50)
        t = frames - frames[0]
        x = x0 + speed * t * np.cos(ang) + rng.normal(0, 0.6, len(t))
        y = y0 - speed * t * np.sin(ang) + rng.normal(0, 0.6, len(t))
        L, W = rng.uniform(60, 400), rng.uniform(40, 120)
        horiz = np.abs(np.cos(ang)) > 0.7
        bw = np.where(horiz, L, W) + rng.normal(0, 1.5, len(t))
        bh = np.where(horiz, W, L) + rng.normal(0, 1.5, len(t))
        cls = np.full(len(t), float(rng.integers(0, 4)))
        flick = rng.random(len(t)) < 0.2
        cls[flick] = rng.integers(0, 4, flick.sum())
        conf = rng.uniform(0.25, 0.99, len(t))
        if stabilised:
            dx, dy = rng.normal(0, 3, len(t)), rng.normal(0, 3, len(t))
            rows.append(np.stack([frames, np.full(len(t), tid), x, y, bw, bh, x + dx, y + dy, bw * 1.01, bh * 0.99, cls, conf], 1))
        else:
            rows.append(np.stack([frames, np.full(len(t), tid), x, y, bw, bh, cls, conf], 1))
    tr = np.concatenate(rows, 0).astype(np.float64)
    return tr[np.lexsort((tr[:, 1], tr[:, 0]))]        # frame-major, as the extraction loop appends them


CASES = [("stab_small", 1, 20, 160, True, True), ("raw_small", 3, 20, 160, False, True), ("tiny", 4, 3, 12, True, True),
         ("stab_medium", 5, 80, 400, True, False), ("stab_large", 2, 400, 1500, True, False), ("raw_large", 6, 300, 1200, False, False)]


def digest(arr: np.ndarray) -> bytes:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(arr, np.float64).tobytes()).digest()


CFG_DIM = dict(gsd=0.02725, eps=4, r0=1.25, theta_bar=15, tau_c={0: 1.83, 1: 2.85, 2: 1.70, 3: 1.80, -1: 1.70})


def main():
    import geotrax_b200
    geotrax_b200.install_shims()
    import geotrax.extract as ex
    ex.get_video_dimensions = lambda src: (3840, 2160)
    logger = logging.getLogger("golden")
    out = {}
    # small cases are stored whole (inputs and every stage's output); the larger ones as SHA-256 of the output bytes -- the test regenerates
    # their inputs from the seed with synth_tracks() (this file is importable without the reference)
    for name, seed, n_tracks, n_frames, stab, whole in CASES:
        tr = synth_tracks(seed, n_tracks, n_frames, stab)
        a = ex.remove_short_tracks(tr.copy(), logger, 3)
        b = ex.calculate_unique_classes(a.copy())
        cfg_main = dict(args=types.SimpleNamespace(source="none.mp4", interpolate=True), extraction=dict(min_track_length=3, dimension_estimation=CFG_DIM),
                        tracker=dict(active="botsort", botsort=dict(track_buffer=30)))
        c = ex.estimate_vehicle_dimensions(b.copy(), cfg_main)
        d = ex.interpolate_tracks(c.copy(), logger, 30)
        e = ex.postprocess_tracks(tr.copy(), dict(main=cfg_main), logger)
        assert np.array_equal(e, d, equal_nan=True)
        for stage, arr in (("short", a), ("classes", b), ("dims", c), ("interp", d)):
            out[f"{name}/{stage}/shape"] = np.array(arr.shape)
            out[f"{name}/{stage}/sha256"] = np.frombuffer(digest(arr), np.uint8)
            if whole:
                out[f"{name}/{stage}"] = arr
        print(name, tr.shape, "->", a.shape, c.shape, d.shape, "nan dims:", int(np.isnan(c[:, -1]).sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "postprocess_golden.npz"), **out)


if __name__ == "__main__":
    main()
