"""Seed-deterministic synthetic bird's-eye-view frames (SURVEY.md section 8d "synthetic frame generator").

The reference's sample clip ``data/U_video_cut.mp4`` is absent from the mount (/root/reference/.MISSING_LARGE_BLOBS) and
there is no network, so parity tests and bench.py use frames generated here: a multi-scale textured ground plane with
road markings, warped per frame by a known small homography (the golden drift envelope: a few px of translation,
<= 0.1 deg rotation, perspective <= 1e-6 -- BASELINE.md section 1), plus 100-200 filled rectangles ("vehicles", golden box
statistics: w 60-110 px, h 30-50 px) that move linearly and therefore do NOT follow the background homography.
"""
from __future__ import annotations

from typing import List, Tuple

import cv2
import numpy as np


def _texture(h: int, w: int, rng: np.random.Generator) -> np.ndarray:
    acc = np.zeros((h, w), np.float32)
    amp = 1.0
    for cell in (256, 64, 16, 4):
        gh, gw = max(2, h // cell + 2), max(2, w // cell + 2)
        g = rng.standard_normal((gh, gw)).astype(np.float32)
        acc += amp * cv2.resize(g, (w, h), interpolation=cv2.INTER_CUBIC)
        amp *= 0.6
    acc += 0.35 * rng.standard_normal((h, w)).astype(np.float32)
    acc = (acc - acc.mean()) / (acc.std() + 1e-6)
    return acc


def make_ground(h: int, w: int, seed: int = 0) -> np.ndarray:
    """u8 BGR ground plane of size (h, w)."""
    rng = np.random.default_rng(seed)
    base = _texture(h, w, rng)
    img = np.empty((h, w, 3), np.float32)
    for c, (m, s) in enumerate(((105, 34), (110, 36), (112, 38))):
        img[..., c] = m + s * base + 6.0 * cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), 1.5)
    img = np.clip(img, 0, 255).astype(np.uint8)
    # road markings: long light lines + dashed segments
    n_lines = max(4, w // 320)
    for i in range(n_lines):
        y = int(rng.integers(h // 10, h - h // 10))
        cv2.line(img, (0, y), (w - 1, y + int(rng.integers(-h // 20, h // 20 + 1))), (215, 215, 215), max(1, h // 540))
        x = int(rng.integers(w // 10, w - w // 10))
        for y0 in range(0, h, max(8, h // 27)):
            cv2.line(img, (x, y0), (x, y0 + max(4, h // 54)), (225, 225, 225), max(1, h // 540))
    return img


def make_boxes(n: int, h: int, w: int, rng: np.random.Generator, scale: float = 1.0) -> np.ndarray:
    """(n, 4) xywh f32, golden-like vehicle sizes at 4K (scaled for smaller frames)."""
    bw = rng.uniform(60, 110, n) * scale
    bh = rng.uniform(30, 50, n) * scale
    flip = rng.random(n) < 0.35
    bw, bh = np.where(flip, bh, bw), np.where(flip, bw, bh)
    xc = rng.uniform(bw / 2 + 2, w - bw / 2 - 2, n)
    yc = rng.uniform(bh / 2 + 2, h - bh / 2 - 2, n)
    return np.stack([xc, yc, bw, bh], 1).astype(np.float32)


def draw_vehicles(img: np.ndarray, boxes: np.ndarray, rng: np.random.Generator) -> None:
    for xc, yc, bw, bh in boxes:
        col = tuple(int(v) for v in rng.integers(20, 250, 3))
        p0 = (int(round(xc - bw / 2)), int(round(yc - bh / 2)))
        p1 = (int(round(xc + bw / 2)), int(round(yc + bh / 2)))
        cv2.rectangle(img, p0, p1, col, -1)
        cv2.rectangle(img, (p0[0] + 3, p0[1] + 3), (p1[0] - 3, p1[1] - 3), tuple(min(255, c + 35) for c in col), 1)


def small_homography(rng: np.random.Generator, h: int, w: int, max_t: float = 6.0, max_rot_deg: float = 0.1, max_persp: float = 1e-6):
    """current -> reference homography inside the golden drift envelope (full-resolution pixel units)."""
    t = rng.uniform(-max_t, max_t, 2)
    a = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg))
    s = 1.0 + rng.uniform(-5e-4, 5e-4)
    cx, cy = w / 2.0, h / 2.0
    R = np.array([[s * np.cos(a), -s * np.sin(a), 0], [s * np.sin(a), s * np.cos(a), 0], [0, 0, 1.0]])
    C = np.array([[1, 0, cx], [0, 1, cy], [0, 0, 1.0]])
    Ci = np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1.0]])
    H = C @ R @ Ci
    H[0, 2] += t[0]
    H[1, 2] += t[1]
    H[2, 0], H[2, 1] = rng.uniform(-max_persp, max_persp, 2)
    return H / H[2, 2]


def make_flight(n_frames: int, h: int = 2160, w: int = 3840, seed: int = 0, n_vehicles: int = 132, noise: float = 2.0,
                max_t: float = 6.0) -> Tuple[List[np.ndarray], List[np.ndarray], List[np.ndarray]]:
    """-> (frames u8 BGR, boxes xywh per frame, H_gt per frame (current -> reference; identity for frame 0))."""
    rng = np.random.default_rng(seed)
    ground = make_ground(h, w, seed)
    scale = w / 3840.0
    boxes0 = make_boxes(n_vehicles, h, w, rng, scale)
    vel = rng.uniform(-3.0, 3.0, (n_vehicles, 2)).astype(np.float32) * scale
    frames, boxes, Hs = [], [], []
    for t in range(n_frames):
        H = np.eye(3) if t == 0 else small_homography(rng, h, w, max_t * scale)
        # p_ref = H p_cur  =>  cur(p) = ground(H p): warpPerspective with WARP_INVERSE_MAP uses dst(p) = src(M p)
        img = ground.copy() if t == 0 else cv2.warpPerspective(ground, H, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                                                borderMode=cv2.BORDER_REFLECT_101)
        b = boxes0.copy()
        b[:, :2] += vel * t
        b[:, 0] = np.clip(b[:, 0], b[:, 2] / 2 + 1, w - b[:, 2] / 2 - 1)
        b[:, 1] = np.clip(b[:, 1], b[:, 3] / 2 + 1, h - b[:, 3] / 2 - 1)
        draw_vehicles(img, b, np.random.default_rng(seed + 7))  # same colours every frame
        if noise > 0:
            nz = rng.standard_normal((h, w, 1)).astype(np.float32) * noise
            img = np.clip(img.astype(np.float32) + nz, 0, 255).astype(np.uint8)
        frames.append(img)
        boxes.append(b)
        Hs.append(H)
    return frames, boxes, Hs


def make_frames(n: int, h: int = 2160, w: int = 3840, seed: int = 0, n_vehicles: int = 132) -> np.ndarray:
    """(n, h, w, 3) u8 BGR batch (bench input)."""
    frames, _, _ = make_flight(n, h, w, seed, n_vehicles)
    return np.stack(frames)


def bgr_to_nv12(frame_bgr: np.ndarray) -> np.ndarray:
    """BGR24 -> NV12 (u8 [H * 3 / 2][W]: Y plane, then interleaved U, V rows) through cv2: synthetic input for the decoder-format ingest."""
    import cv2
    h, w = frame_bgr.shape[:2]
    i420 = cv2.cvtColor(frame_bgr, cv2.COLOR_BGR2YUV_I420)
    flat, q = i420.reshape(-1), (h // 2) * (w // 2)
    out = np.empty((h * 3 // 2, w), np.uint8)
    out[:h] = i420[:h]
    out[h:] = np.stack([flat[h * w:h * w + q].reshape(h // 2, w // 2), flat[h * w + q:h * w + 2 * q].reshape(h // 2, w // 2)], -1).reshape(h // 2, w)
    return out
