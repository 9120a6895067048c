"""Vectorised `postprocess_tracks` (second half): the reference's track post-processing without its Python loops.

Mirror of /root/reference/geotrax/extract.py:296-484 -- `remove_short_tracks`, `calculate_unique_classes`, `estimate_vehicle_dimensions`,
`interpolate_tracks`, `postprocess_tracks` with the same signatures, column conventions, log messages and -- bit for bit -- the same
output arrays (tests/test_postprocess.py against golden vectors produced by the reference's own functions).  The reference walks the
rows in Python: `remove_short_tracks` alone is O(track ids x rows), the class vote, the dimension estimate and the gap interpolation are
one interpreter iteration per row.  That is invisible on the 100-frame sample and minutes on a 27,000-frame flight (3.5 M rows), i.e.
far longer than the extraction the B200 path finishes in seconds.  Here every step is numpy over the whole table:

* grouping = one stable argsort by track id (row order inside a track is preserved, which is what every per-track list of the reference is);
* class vote = `np.bincount` with weights over (track, class) pairs -- it accumulates in row order like the reference's dict, so the
  float sums are identical; ties go to the lowest class id;
* the azimuth filter of the dimension estimate is sequential PER TRACK (an anchor point moves whenever the vehicle has travelled r0);
  it is run as "find the next anchor with one vectorised distance test over a look-ahead window" instead of one Python step per row;
* the 25th percentile is `np.percentile` per track on the filtered values (same interpolation rule), NaN for tracks without any;
* gap interpolation builds all missing rows at once with `np.repeat` and the reference's formula `a * (1 - alpha) + b * alpha`.

Host-side numpy on purpose: the table is a few hundred MB at most and lives on the host, next to the tracker that produced it.
"""
from __future__ import annotations

import logging
from typing import Dict, Optional, Tuple

import numpy as np

_CARDINALS = np.array([0, np.pi / 2, np.pi, -np.pi / 2, -np.pi])


def remove_short_tracks(tracks: np.ndarray, logger: logging.Logger, min_length: int = 3) -> np.ndarray:
    """extract.py:362-378: drop every track with fewer than `min_length` rows (row order kept)."""
    if tracks.size == 0:
        return tracks
    _, inv, cnt = np.unique(tracks[:, 1].astype(int), return_inverse=True, return_counts=True)
    short = cnt < min_length
    if short.any():
        tracks = tracks[~short[inv]]
        logger.info(f'{int(short.sum())} short tracks removed.')
    return tracks


def calculate_unique_classes(tracks: np.ndarray) -> np.ndarray:
    """extract.py:381-404: one class per track = the highest confidence-weighted vote, ties to the lowest class id.  In place, like the reference."""
    if tracks.size == 0:
        return tracks
    ids = tracks[:, 1].astype(np.int64)
    cls = tracks[:, -2].astype(np.int64)
    uid, iid = np.unique(ids, return_inverse=True)
    ucl, icl = np.unique(cls, return_inverse=True)
    pair = iid * len(ucl) + icl
    votes = np.bincount(pair, weights=tracks[:, -1], minlength=len(uid) * len(ucl)).reshape(len(uid), len(ucl))
    seen = np.bincount(pair, minlength=len(uid) * len(ucl)).reshape(len(uid), len(ucl)) > 0
    votes = np.where(seen, votes, -np.inf)                  # a class the track never had cannot win (weights may be <= 0 in principle)
    best = ucl[np.argmax(votes, axis=1)]                    # argmax returns the first maximum = the lowest class id
    tracks[:, -2] = best[iid]
    return tracks


def _azimuth_mask(x: np.ndarray, y: np.ndarray, radius: float, theta: float) -> Tuple[np.ndarray, bool]:
    """extract.py:444-456 for one track: rows between two anchors count when the travel direction between them is within `theta` of a
    cardinal direction.  Returns (mask, any anchor found)."""
    n = len(x)
    mask = np.zeros(n, dtype=bool)
    prev, found = 0, False
    while prev + 1 < n:
        lo, win, idx = prev + 1, 64, -1
        while lo < n:                                        # look-ahead windows: a moving vehicle re-anchors every few rows
            hi = min(lo + win, n)
            d = np.sqrt((x[lo:hi] - x[prev]) ** 2 + (y[lo:hi] - y[prev]) ** 2)
            hit = np.flatnonzero(d >= radius)
            if hit.size:
                idx = lo + int(hit[0])
                break
            lo, win = hi, win * 4
        if idx < 0:
            break
        found = True
        azimuth = np.arctan2(-(y[idx] - y[prev]), x[idx] - x[prev])
        if np.any(np.abs(azimuth - _CARDINALS) <= theta):
            mask[prev:idx] = True
        prev = idx
    return mask, found


def estimate_vehicle_dimensions(tracks: np.ndarray, config: Dict, frame_size: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """extract.py:407-476: appends the per-track length / width estimate (25th percentile over the rows where the vehicle is fully
    visible and travels along an image axis; NaN without such rows).  `frame_size` = (w, h) replaces the reference's look at the source
    video (`get_video_dimensions(config['args'].source)`) when the caller knows it."""
    if frame_size is None:
        import cv2
        reader = cv2.VideoCapture(str(config['args'].source))
        frame_size = (int(reader.get(cv2.CAP_PROP_FRAME_WIDTH)), int(reader.get(cv2.CAP_PROP_FRAME_HEIGHT)))
        reader.release()
    w_I, h_I = frame_size
    de = config['extraction']['dimension_estimation']
    eps = de['eps']
    vis = (tracks[:, 2] - tracks[:, 4] / 2 > eps) & (tracks[:, 3] - tracks[:, 5] / 2 > eps)
    vis &= (tracks[:, 2] + tracks[:, 4] / 2 < w_I - 1 - eps) & (tracks[:, 3] + tracks[:, 5] / 2 < h_I - 1 - eps)
    valid = tracks[vis]
    idx_x, idx_y, idx_c = (6, 7, 10) if valid.shape[1] > 8 else (2, 3, 6)
    radius = de['r0'] / de['gsd']
    theta = np.deg2rad(de['theta_bar'])
    tau_c = de['tau_c']

    uid, inv = np.unique(valid[:, 1].astype(int), return_inverse=True) if len(valid) else (np.zeros(0, int), np.zeros(0, int))
    order = np.argsort(inv, kind="stable")                   # rows of a track stay in table order
    bounds = np.searchsorted(inv[order], np.arange(len(uid) + 1))
    length = np.maximum(valid[:, 4], valid[:, 5])
    width = np.minimum(valid[:, 4], valid[:, 5])
    est_l = np.full(len(uid), np.nan)
    est_w = np.full(len(uid), np.nan)
    for k in range(len(uid)):
        rows = order[bounds[k]:bounds[k + 1]]
        ls, ws = length[rows], width[rows]
        mask, found = _azimuth_mask(valid[rows, idx_x], valid[rows, idx_y], radius, theta)
        if not found:                                        # the vehicle never travelled r0: keep the rows that look elongated enough
            mask = ls >= ws * tau_c.get(int(valid[rows[0], idx_c]), tau_c[-1])
        if mask.any():
            est_l[k] = np.percentile(ls[mask], 25)
            est_w[k] = np.percentile(ws[mask], 25)
    out = np.append(tracks, np.zeros((len(tracks), 2)), axis=1)
    all_ids = tracks[:, 1].astype(int)
    pos = np.searchsorted(uid, all_ids)
    pos_c = np.minimum(pos, max(len(uid) - 1, 0))
    known = (pos < len(uid)) & (uid[pos_c] == all_ids) if len(uid) else np.zeros(len(tracks), bool)
    out[:, -2] = np.where(known, est_l[pos_c] if len(uid) else np.nan, np.nan)
    out[:, -1] = np.where(known, est_w[pos_c] if len(uid) else np.nan, np.nan)
    return out


def interpolate_tracks(tracks: np.ndarray, logger: logging.Logger, max_gap: int) -> np.ndarray:
    """extract.py:309-359: fills per-track frame gaps of at most `max_gap` frames by linear interpolation of EVERY column (as the reference
    does), appends the is_interpolated flag column, returns the table sorted by (track id, frame)."""
    if tracks.size == 0:
        return tracks
    order = np.lexsort((tracks[:, 0], tracks[:, 1]))         # by track id, then frame: consecutive rows of a track are neighbours
    t = tracks[order]
    frames = t[:, 0].astype(int)
    same = t[1:, 1] == t[:-1, 1]
    gap = np.where(same, frames[1:] - frames[:-1], 0)
    skipped = int(np.count_nonzero(gap > max_gap))
    fill = np.flatnonzero((gap > 1) & (gap <= max_gap))      # index i - 1 of the row BEFORE the gap
    flag = np.zeros((len(tracks), 1), dtype=tracks.dtype)
    out = np.concatenate([tracks, flag], axis=1)
    if skipped > 0:
        logger.warning(f"Skipped {skipped} frame gap(s) exceeding the tracker's track_buffer ({max_gap} frames); left unfilled.")
    if fill.size:
        g = gap[fill]
        rep = np.repeat(np.arange(fill.size), g - 1)                               # one entry per missing frame
        step = np.arange(rep.size) - np.repeat(np.cumsum(g - 1) - (g - 1), g - 1) + 1
        alpha = (step / g[rep])[:, None]
        a, b = t[fill[rep]], t[fill[rep] + 1]
        rows = a * (1.0 - alpha) + b * alpha
        rows[:, 0] = (frames[fill[rep]] + step).astype(np.float64)
        rows = np.concatenate([rows.astype(tracks.dtype), np.ones((len(rows), 1), dtype=tracks.dtype)], axis=1)
        out = np.concatenate([out, rows], axis=0)
        out = out[np.lexsort((out[:, 0], out[:, 1]))]
        logger.info(f"Interpolated {len(rows)} missing frame(s) across {len(np.unique(t[fill, 1]))} track(s).")
    return out


def postprocess_tracks(tracks: np.ndarray, config: Dict, logger: logging.Logger, frame_size: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """extract.py:296-306 (same `config` structure: config['main'] holds extraction / tracker / args)."""
    main = config['main']
    tracks = remove_short_tracks(tracks, logger, main['extraction']['min_track_length'])
    tracks = calculate_unique_classes(tracks)
    tracks = estimate_vehicle_dimensions(tracks, main, frame_size)
    if main['args'].interpolate:
        max_gap = main['tracker'][main['tracker']['active']]['track_buffer']
        tracks = interpolate_tracks(tracks, logger, max_gap)
    return tracks
