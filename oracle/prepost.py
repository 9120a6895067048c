"""numpy / torch restatement of ultralytics pre- and post-processing (oracle; test infrastructure only).

Restates (un-vendored): ``data/augment.py:LetterBox``, ``engine/predictor.py:preprocess``,
``utils/nms.py:non_max_suppression`` + ``nms_rotated``, ``utils/metrics.py:batch_probiou``,
``utils/ops.py:{scale_boxes,clip_boxes,xywh2xyxy,regularize_rboxes}``.  Reference call site:
/root/reference/geotrax/extract.py:153 (``model.track``) with the keys of
/root/reference/geotrax/cfg/default.yaml:229-250 (conf 0.25, iou 0.7, max_det 1000, classes, agnostic_nms).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np
import torch
import torchvision

MAX_WH = 7680
MAX_NMS = 30000


def letterbox_params(h0: int, w0: int, imgsz: int = 1920, stride: int = 32):
    """-> (new_w, new_h, top, bottom, left, right, gain).  LetterBox(auto=True, center=True, scaleup=True)."""
    r = min(imgsz / h0, imgsz / w0)
    new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = (imgsz - new_w) % stride, (imgsz - new_h) % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_w, new_h, top, bottom, left, right, r


def letterbox_u8(frame_bgr: np.ndarray, imgsz: int = 1920, stride: int = 32) -> np.ndarray:
    """u8 HWC BGR -> u8 HWC BGR letterboxed (pad value 114)."""
    h0, w0 = frame_bgr.shape[:2]
    new_w, new_h, top, bottom, left, right, _ = letterbox_params(h0, w0, imgsz, stride)
    img = frame_bgr
    if (w0, h0) != (new_w, new_h):
        img = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))


def resize_linear_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """Integer restatement of ``cv2.resize(u8, (new_w, new_h), interpolation=cv2.INTER_LINEAR)`` (OpenCV imgproc/src/resize.cpp:
    ``resize`` -> ``HResizeLinear`` / ``VResizeLinear<uchar,int,short>``): 11-bit coefficients ``cvRound(f * 2048)`` from the
    *float* fractions of ``(d + 0.5) * scale - 0.5`` with ``scale = 1 / (dst / src)`` in double, source indices clamped at the borders,
    horizontal pass in int32, vertical pass ``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``; an exact 2x2 decimation
    is rerouted to INTER_AREA = ``(a + b + c + d + 2) >> 2``.  Pinned against cv2 in
    tests/test_oracle_model.py::test_resize_linear_restatement_matches_cv2; the CUDA letterbox follows this function."""
    src = img if img.ndim == 3 else img[..., None]
    h, w, cn = src.shape
    if (w, h) == (new_w, new_h):
        return img.copy()
    if w == 2 * new_w and h == 2 * new_h:
        s = src.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8).reshape((new_h, new_w) + ((cn,) if img.ndim == 3 else ()))

    def axis(n_src, n_dst):
        scale = 1.0 / (float(n_dst) / float(n_src))
        f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        i = np.floor(f).astype(np.int64)
        f = (f - i.astype(np.float32)).astype(np.float32)
        return i, f

    sx, fx = axis(w, new_w)
    lo, hi = sx < 0, sx >= w - 1           # horizontal: the fraction is zeroed where the index is clamped
    fx = np.where(lo | hi, np.float32(0), fx)
    sx = np.clip(sx, 0, w - 1)
    sx1 = np.minimum(sx + 1, w - 1)
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int32)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int32)
    sy, fy = axis(h, new_h)               # vertical: rows are clamped, the fraction is kept
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int32)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int32)
    r0, r1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    s = src.astype(np.int32)
    hrow = s[:, sx, :] * a0[None, :, None] + s[:, sx1, :] * a1[None, :, None]       # (h, new_w, cn) int32
    S0, S1 = hrow[r0], hrow[r1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if img.ndim == 3 else out[..., 0]


def nv12_to_bgr(nv12: np.ndarray) -> np.ndarray:
    """Integer restatement of ``cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12)`` (OpenCV imgproc/src/color_yuv.simd.hpp, ITU-R BT.601 limited
    range, 20-bit fixed point): nv12 is u8 [H * 3 / 2][W] (Y plane, then interleaved U, V rows at half resolution).  The decoder-format
    ingest of the CUDA path (gt_set_input_format(GT_INPUT_NV12)) follows this function; pinned against cv2 in
    tests/test_oracle_model.py::test_nv12_restatement_matches_cv2.  (The reference's reader converts with FFmpeg swscale, whose rounding
    is not reproducible here: this ingest is SURVEY 8f rank 1, next to -- not inside -- the drop-in path.)"""
    h = nv12.shape[0] * 2 // 3
    w = nv12.shape[1]
    y = nv12[:h].astype(np.int64)
    uv = nv12[h:].reshape(h // 2, w // 2, 2).astype(np.int64)
    u = np.repeat(np.repeat(uv[..., 0], 2, 0), 2, 1) - 128
    v = np.repeat(np.repeat(uv[..., 1], 2, 0), 2, 1) - 128
    cy, cub, cug, cvg, cvr, sh = 1220542, 2116026, -409993, -852492, 1673527, 20
    yy = np.maximum(0, y - 16) * cy
    half = 1 << (sh - 1)
    r, g, b = (yy + half + cvr * v) >> sh, (yy + half + cvg * v + cug * u) >> sh, (yy + half + cub * u) >> sh
    return np.stack([b, g, r], -1).clip(0, 255).astype(np.uint8)


def bgr_to_nv12(frame_bgr: np.ndarray) -> np.ndarray:
    """Test helper: BGR24 -> NV12 through cv2 (I420, then U and V interleaved)."""
    h, w = frame_bgr.shape[:2]
    i420 = cv2.cvtColor(frame_bgr, cv2.COLOR_BGR2YUV_I420)
    out = np.empty((h * 3 // 2, w), np.uint8)
    out[:h] = i420[:h]
    flat, q = i420.reshape(-1), (h // 2) * (w // 2)
    u = flat[h * w:h * w + q].reshape(h // 2, w // 2)
    v = flat[h * w + q:h * w + 2 * q].reshape(h // 2, w // 2)
    out[h:] = np.stack([u, v], -1).reshape(h // 2, w)
    return out


def preprocess(frames_bgr: Sequence[np.ndarray], imgsz: int = 1920) -> torch.Tensor:
    """list of u8 HWC BGR -> f32 NCHW RGB in [0,1]."""
    im = np.stack([letterbox_u8(f, imgsz) for f in frames_bgr])
    im = np.ascontiguousarray(im[..., ::-1].transpose((0, 3, 1, 2)))
    return torch.from_numpy(im).float() / 255.0


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:4] / 2
    y[..., :2] = xy - wh
    y[..., 2:4] = xy + wh
    return y


def _cov(boxes: torch.Tensor):
    gbbs = torch.cat((boxes[:, 2:4].pow(2) / 12, boxes[:, 4:]), dim=-1)
    a, b, c = gbbs.split(1, dim=-1)
    cos, sin = c.cos(), c.sin()
    cos2, sin2 = cos.pow(2), sin.pow(2)
    return a * cos2 + b * sin2, a * sin2 + b * cos2, (a - b) * cos * sin


def batch_probiou(obb1: torch.Tensor, obb2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """(N,5) x (M,5) xywhr -> (N,M) ProbIoU."""
    x1, y1 = obb1[..., :2].split(1, dim=-1)
    x2, y2 = (x.squeeze(-1)[None] for x in obb2[..., :2].split(1, dim=-1))
    a1, b1, c1 = _cov(obb1)
    a2, b2, c2 = (x.squeeze(-1)[None] for x in _cov(obb2))
    t1 = (((a1 + a2) * (y1 - y2).pow(2) + (b1 + b2) * (x1 - x2).pow(2)) / ((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2) + eps)) * 0.25
    t2 = (((c1 + c2) * (x2 - x1) * (y1 - y2)) / ((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2) + eps)) * 0.5
    t3 = (((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2))
          / (4 * ((a1 * b1 - c1.pow(2)).clamp_(0) * (a2 * b2 - c2.pow(2)).clamp_(0)).sqrt() + eps) + eps).log() * 0.5
    bd = (t1 + t2 + t3).clamp(eps, 100.0)
    hd = (1.0 - (-bd).exp() + eps).sqrt()
    return 1 - hd


def nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, threshold: float) -> torch.Tensor:
    """Fast-NMS semantics: j survives iff no higher-scored i has probiou(i,j) >= thr (suppressed boxes still suppress)."""
    order = torch.argsort(scores, descending=True, stable=True)
    b = boxes[order]
    ious = batch_probiou(b, b).triu_(diagonal=1)
    pick = torch.nonzero(ious.max(dim=0)[0] < threshold).squeeze(-1)
    return order[pick]


def non_max_suppression(pred: torch.Tensor, conf_thres=0.25, iou_thres=0.7, classes: Optional[Sequence[int]] = None,
                        agnostic=False, max_det=1000, nc=0, rotated=False, return_idxs=False):
    """pred (B, 4+nc+nm, A) -> list of (n, 6+nm) [xyxy|xywh, conf, cls, extra].  Also the kept anchor indices."""
    bs = pred.shape[0]
    nc = nc or (pred.shape[1] - 4)
    nm = pred.shape[1] - nc - 4
    mi = 4 + nc
    xc = pred[:, 4:mi].amax(1) > conf_thres
    pred = pred.transpose(-1, -2).clone()
    if not rotated:
        pred[..., :4] = xywh2xyxy(pred[..., :4])
    out = [torch.zeros((0, 6 + nm))] * bs
    idxs = [torch.zeros((0,), dtype=torch.long)] * bs
    cls_t = None if classes is None else torch.tensor(list(classes), dtype=torch.float32)
    for xi, x in enumerate(pred):
        ai = torch.nonzero(xc[xi]).squeeze(-1)
        x = x[ai]
        if not x.shape[0]:
            continue
        box, cls, extra = x.split((4, nc, nm), 1)
        conf, j = cls.max(1, keepdim=True)
        keep = conf.view(-1) > conf_thres
        x = torch.cat((box, conf, j.float(), extra), 1)[keep]
        ai = ai[keep]
        if cls_t is not None:
            k2 = (x[:, 5:6] == cls_t).any(1)
            x, ai = x[k2], ai[k2]
        n = x.shape[0]
        if not n:
            continue
        if n > MAX_NMS:
            top = x[:, 4].argsort(descending=True, stable=True)[:MAX_NMS]
            x, ai = x[top], ai[top]
        c = x[:, 5:6] * (0 if agnostic else MAX_WH)
        scores = x[:, 4]
        if rotated:
            boxes = torch.cat((x[:, :2] + c, x[:, 2:4], x[:, -1:]), dim=-1)
            i = nms_rotated(boxes, scores, iou_thres)
        else:
            i = torchvision.ops.nms(x[:, :4] + c, scores, iou_thres)
        i = i[:max_det]
        out[xi], idxs[xi] = x[i], ai[i]
    return (out, idxs) if return_idxs else out


def clip_boxes(boxes: torch.Tensor, shape: Tuple[int, int]) -> torch.Tensor:
    boxes[..., 0].clamp_(0, shape[1])
    boxes[..., 1].clamp_(0, shape[0])
    boxes[..., 2].clamp_(0, shape[1])
    boxes[..., 3].clamp_(0, shape[0])
    return boxes


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape, xywh=False) -> torch.Tensor:
    """Letterboxed-pixel boxes -> original-frame pixels (in place on a clone)."""
    boxes = boxes.clone()
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    boxes[..., 0] -= pad_x
    boxes[..., 1] -= pad_y
    if not xywh:
        boxes[..., 2] -= pad_x
        boxes[..., 3] -= pad_y
    boxes[..., :4] /= gain
    return boxes if xywh else clip_boxes(boxes, img0_shape)


def regularize_rboxes(rboxes: torch.Tensor) -> torch.Tensor:
    x, y, w, h, t = rboxes.unbind(dim=-1)
    swap = t % math.pi >= math.pi / 2
    w_ = torch.where(swap, h, w)
    h_ = torch.where(swap, w, h)
    t = t % (math.pi / 2)
    return torch.stack([x, y, w_, h_, t], dim=-1)


def xyxy2xywh(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


def postprocess_detect(decoded: torch.Tensor, in_shape, orig_shape, conf=0.25, iou=0.7, classes=None, agnostic=True,
                       max_det=1000) -> List[torch.Tensor]:
    """decoded (B,4+nc,A) -> per image (n,6) [x1,y1,x2,y2,conf,cls] in original-frame pixels."""
    outs = non_max_suppression(decoded, conf, iou, classes, agnostic, max_det, nc=decoded.shape[1] - 4)
    res = []
    for o in outs:
        o = o.clone()
        if o.shape[0]:
            o[:, :4] = scale_boxes(in_shape, o[:, :4], orig_shape)
        res.append(o)
    return res


def postprocess_obb(decoded: torch.Tensor, in_shape, orig_shape, conf=0.25, iou=0.7, classes=None, agnostic=True,
                    max_det=1000) -> List[torch.Tensor]:
    """decoded (B,4+nc+1,A) -> per image (n,7) [x,y,w,h,r,conf,cls] in original-frame pixels."""
    nc = decoded.shape[1] - 5
    outs = non_max_suppression(decoded, conf, iou, classes, agnostic, max_det, nc=nc, rotated=True)
    res = []
    for o in outs:
        if not o.shape[0]:
            res.append(torch.zeros((0, 7)))
            continue
        rb = regularize_rboxes(torch.cat([o[:, :4], o[:, -1:]], dim=-1))
        rb[:, :4] = scale_boxes(in_shape, rb[:, :4], orig_shape, xywh=True)
        res.append(torch.cat([rb, o[:, 4:6]], dim=-1))
    return res


def clahe_u8(src: np.ndarray, clip_limit: float = 2.0, tiles=(8, 8)) -> np.ndarray:
    """``cv2.createCLAHE(clipLimit, tileGridSize).apply(src)`` for 8-bit images, restated (OpenCV imgproc/src/clahe.cpp).

    stabilo applies it to the gray frame when ``clahe: true`` (/root/reference/geotrax/cfg/stable.yaml:115; clip 2.0, 8x8 tiles).
    Steps: extend the image to a multiple of the grid with BORDER_REFLECT_101; per tile a 256-bin histogram, clipped at
    ``max(1, int(clip * area / 256))``, the excess redistributed (equal batch + strided residual); LUT = rint(cumsum * 255 / area);
    per pixel the bilinear blend of the four neighbouring tile LUTs in float32 (every product and sum rounded separately).
    Pinned bit-exact against cv2 in tests/test_oracle_model.py; the CUDA kernels (csrc/clahe.cu) are checked against this / cv2."""
    import cv2

    tx, ty = tiles
    h, w = src.shape
    ext = src if (w % tx == 0 and h % ty == 0) else cv2.copyMakeBorder(src, 0, ty - (h % ty), 0, tx - (w % tx), cv2.BORDER_REFLECT_101)
    tw, th = ext.shape[1] // tx, ext.shape[0] // ty
    area = tw * th
    lut_scale = np.float32(255.0) / np.float32(area)
    clip = max(int(clip_limit * area / 256), 1) if clip_limit > 0 else 0
    luts = np.zeros((ty, tx, 256), np.uint8)
    for j in range(ty):
        for i in range(tx):
            hist = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(hist - clip, 0).sum())
                hist = np.minimum(hist, clip)
                batch = clipped // 256
                resid = clipped - batch * 256
                hist += batch
                if resid:
                    hist[np.arange(0, 256, max(256 // resid, 1))[:resid]] += 1
            luts[j, i] = np.clip(np.rint(np.cumsum(hist).astype(np.float32) * lut_scale), 0, 255).astype(np.uint8)

    def axis(n, tile, nt):
        f = np.arange(n, dtype=np.float32) * (np.float32(1.0) / np.float32(tile)) - np.float32(0.5)
        t1 = np.floor(f).astype(np.int32)
        a = (f - t1.astype(np.float32)).astype(np.float32)
        return np.maximum(t1, 0), np.minimum(t1 + 1, nt - 1), a, np.float32(1.0) - a

    x1, x2, xa, xa1 = axis(w, tw, tx)
    y1, y2, ya, ya1 = axis(h, th, ty)
    out = np.empty_like(src)
    for y in range(h):
        v = src[y]
        top = luts[y1[y], x1, v].astype(np.float32) * xa1 + luts[y1[y], x2, v].astype(np.float32) * xa
        bot = luts[y2[y], x1, v].astype(np.float32) * xa1 + luts[y2[y], x2, v].astype(np.float32) * xa
        out[y] = np.clip(np.rint(top * ya1[y] + bot * ya[y]), 0, 255).astype(np.uint8)
    return out


def warp_perspective_u8(src: np.ndarray, H: np.ndarray, dsize=None) -> np.ndarray:
    """``cv2.warpPerspective(src, H, (w, h))`` (INTER_LINEAR, BORDER_CONSTANT 0) for 8-bit images, restated -- what the reference's
    visualisation does with every frame and its transform (/root/reference/geotrax/visualize.py:285-289).

    OpenCV inverts H, walks the destination in 64-pixel-wide blocks with ``X0 = M0*bx + M1*y + M2`` per block row and
    ``X = round(((X0 + M0*x1) * (32 / W)))`` in double precision (1/32-pixel fixed point), and blends the 2 x 2 source pixels with the
    15-bit weights ``(32 - ay)(32 - ax) * 32`` ... , rounding with ``(sum + 2^14) >> 15``.  Pinned bit-exact against cv2 in
    tests/test_oracle_model.py; the CUDA kernel (gt_warp_frames) is checked against cv2 on the GPU."""
    import cv2

    h, w = src.shape[:2]
    W, Hh = dsize if dsize is not None else (w, h)
    M = cv2.invert(np.asarray(H, np.float64))[1].ravel()
    cn = src.shape[2] if src.ndim == 3 else 1
    s = src.reshape(h, w, cn).astype(np.int64)
    out = np.zeros((Hh, W, cn), np.uint8)
    xs = np.arange(W)
    bx = (xs // 64) * 64
    x1 = xs - bx

    def fetch(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        return np.where(ok[:, None], s[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], 0)

    for y in range(Hh):
        X0, Y0, W0 = M[0] * bx + M[1] * y + M[2], M[3] * bx + M[4] * y + M[5], M[6] * bx + M[7] * y + M[8]
        Wd = W0 + M[6] * x1
        Wd = np.where(Wd != 0, 32.0 / np.where(Wd != 0, Wd, 1.0), 0.0)
        X = np.rint(np.clip((X0 + M[0] * x1) * Wd, -2147483648.0, 2147483647.0)).astype(np.int64)
        Y = np.rint(np.clip((Y0 + M[3] * x1) * Wd, -2147483648.0, 2147483647.0)).astype(np.int64)
        sx, sy, ax, ay = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767), X & 31, Y & 31
        acc = (fetch(sy, sx) * ((32 - ay) * (32 - ax) * 32)[:, None] + fetch(sy, sx + 1) * ((32 - ay) * ax * 32)[:, None]
               + fetch(sy + 1, sx) * (ay * (32 - ax) * 32)[:, None] + fetch(sy + 1, sx + 1) * (ay * ax * 32)[:, None])
        out[y] = np.clip((acc + (1 << 14)) >> 15, 0, 255)
    return out if src.ndim == 3 else out[..., 0]
