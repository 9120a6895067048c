"""Thin Python owner of one gt_handle (one per GPU).  numpy / torch buffers in, numpy out; all compute is in the .so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import GT_MAX_KP, GT_ORB_LEVELS, GtError, gt_config, gt_conv_desc


def _ptr(a) -> int:
    """Address of a numpy array or a torch tensor (host or CUDA); None -> 0."""
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "buffer must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return a.data_ptr()
    raise TypeError(type(a))


def normalize_classes(classes) -> Optional[list]:
    """``classes: int | list[int] | None`` (/root/reference/geotrax/cfg/default.yaml:243) -> None or a sorted list of ints."""
    if classes is None:
        return None
    if isinstance(classes, (int, np.integer)):
        classes = [int(classes)]
    out = sorted({int(c) for c in classes})
    for c in out:
        if not 0 <= c < 96:
            raise GtError(f"classes: id {c} is outside the supported range 0..95")
    return out


def classes_to_mask(classes) -> int:
    """32-bit per-call mask of the ABI; only valid when every id is below 32 (Engine._classes_arg routes the rest)."""
    cl = normalize_classes(classes)
    if not cl:
        return 0
    if cl[-1] >= 32:
        raise GtError("classes >= 32 do not fit the per-call mask: use Engine (gt_set_class_filter)")
    m = 0
    for c in cl:
        m |= 1 << c
    return m


class Engine:
    """One B200's worth of the extract hot path: preprocess -> detect -> stabilize -> warp."""

    def __init__(self, frame_hw: Tuple[int, int] = (2160, 3840), imgsz: int = 1920, nc: int = 4, task: str = "detect", max_batch: int = 16,
                 device: int = 0, max_det: int = 1000, act_dtype: str = "fp16", **stab):
        self.lib = _lib.load_library()
        cfg = gt_config()
        self.lib.gt_default_config(C.byref(cfg))
        cfg.frame_h, cfg.frame_w, cfg.imgsz, cfg.nc, cfg.max_batch, cfg.max_det = frame_hw[0], frame_hw[1], imgsz, nc, max_batch, max_det
        cfg.task = _lib.GT_TASK_OBB if task == "obb" else _lib.GT_TASK_DETECT
        assert act_dtype in ("fp16", "bf16")
        cfg.act_dtype = _lib.GT_ACT_FP16 if act_dtype == "fp16" else _lib.GT_ACT_BF16
        self.act_dtype = act_dtype
        for k, v in stab.items():
            if not hasattr(cfg, k):
                raise GtError(f"unknown engine option {k}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.task = task
        self.h = C.c_void_p()
        rc = self.lib.gt_create(C.byref(cfg), device, C.byref(self.h))
        if rc != 0:
            raise GtError(f"gt_create failed ({rc}): {self.lib.gt_last_error(None).decode()}")
        self.max_batch, self.max_det, self.nc = max_batch, max_det, nc
        self.row = 7 if task == "obb" else 6
        A, no = C.c_int32(), C.c_int32()
        self._ck(self.lib.gt_get_raw_head(self.h, 0, None, C.byref(A), C.byref(no)))
        self.A, self.no = A.value, no.value
        nh, nw = C.c_int32(), C.c_int32()
        self._ck(self.lib.gt_get_net_input(self.h, 0, None, C.byref(nh), C.byref(nw)))
        self.net_h, self.net_w = nh.value, nw.value
        self._ck(self.lib.gt_get_gray(self.h, 0, None, C.byref(nh), C.byref(nw)))
        self.work_h, self.work_w = nh.value, nw.value
        self._keep = []

    # -- plumbing --------------------------------------------------------------------------------------------------------
    def _classes_arg(self, classes) -> int:
        """Per-call `classes_mask` argument.  Lists with ids >= 32 (or the empty list = "keep nothing") go through the sticky
        gt_set_class_filter allow-list and the per-call mask is 0; everything else uses the mask and clears the sticky filter."""
        cl = normalize_classes(classes)
        sticky = cl is not None and (len(cl) == 0 or cl[-1] >= 32)
        if sticky:
            arr = (C.c_int32 * max(len(cl), 1))(*cl)
            self._ck(self.lib.gt_set_class_filter(self.h, arr, len(cl)))
            self._sticky_classes = True
            return 0
        if getattr(self, "_sticky_classes", False):
            self._ck(self.lib.gt_set_class_filter(self.h, None, 0))
            self._sticky_classes = False
        return classes_to_mask(cl)

    def health(self) -> int:
        """Cumulative count of anchors dropped because their head row was inf / NaN (16-bit overflow guard); 0 = healthy."""
        n = C.c_int64()
        self._ck(self.lib.gt_get_health(self.h, C.byref(n)))
        return int(n.value)

    def _ck(self, rc: int):
        if rc < 0:
            raise GtError(f"geotrax_b200 error {rc}: {self.lib.gt_last_error(self.h).decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.gt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights -----------------------------------------------------------------------------------------------------------
    def conv_descs(self):
        n = self._ck(self.lib.gt_conv_count(self.h))
        out = []
        for i in range(n):
            d = gt_conv_desc()
            self._ck(self.lib.gt_conv_info(self.h, i, C.byref(d)))
            out.append((d.name.decode(), d.cin, d.cout, d.k, d.stride, d.act))
        return out

    def load_weights(self, folded: Dict[str, Tuple[np.ndarray, np.ndarray]]):
        descs = self.conv_descs()
        n = len(descs)
        W, Bv = (C.c_void_p * n)(), (C.c_void_p * n)()
        keep = []
        for i, (name, cin, cout, k, s, act) in enumerate(descs):
            w, b = folded[name]
            w = np.ascontiguousarray(w, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            assert w.shape == (cout, cin, k, k) and b.shape == (cout,), f"{name}: {w.shape} {b.shape}"
            keep += [w, b]
            W[i], Bv[i] = w.ctypes.data, b.ctypes.data
        self._ck(self.lib.gt_load_weights(self.h, W, Bv, n))

    # -- stages ------------------------------------------------------------------------------------------------------------
    def preprocess(self, frames, stream=None):
        B = int(frames.shape[0])
        self._ck(self.lib.gt_preprocess(self.h, _ptr(frames), B, stream))
        return B

    def set_input_format(self, fmt: str = "bgr24"):
        """'bgr24' (default: what the reference's reader delivers) or 'nv12' (decoder format: u8 [B][H * 3 / 2][W], half the PCIe bytes;
        converted on the device with cv2.cvtColor(COLOR_YUV2BGR_NV12) arithmetic)."""
        self._ck(self.lib.gt_set_input_format(self.h, {"bgr24": _lib.GT_INPUT_BGR24, "nv12": _lib.GT_INPUT_NV12}[fmt]))
        self.input_format = fmt

    def prefetch(self, frames, deferred=False):
        """Start the H2D copy of a pinned host batch; the next preprocess / extract_batch on the same buffer consumes it.
        deferred=True: the copy is started inside the next extract_batch, after that call's own small uploads."""
        self._keep_next = frames
        fn = self.lib.gt_prefetch_frames_deferred if deferred else self.lib.gt_prefetch_frames
        self._ck(fn(self.h, _ptr(frames), int(frames.shape[0])))

    def net_input(self, B: int) -> np.ndarray:
        out = np.empty((B, 3, self.net_h, self.net_w), np.uint8)
        self._ck(self.lib.gt_get_net_input(self.h, B, out.ctypes.data, None, None))
        return out

    def gray(self, B: int) -> np.ndarray:
        out = np.empty((B, self.work_h, self.work_w), np.uint8)
        self._ck(self.lib.gt_get_gray(self.h, B, out.ctypes.data, None, None))
        return out

    def detect(self, B: int, conf=0.25, iou=0.7, agnostic=True, classes=None, want_keep=False, stream=None):
        boxes = np.zeros((B, self.max_det, self.row), np.float32)
        counts = np.zeros((B,), np.int32)
        keep = np.zeros((B, self.max_det), np.int32) if want_keep else None
        self._ck(self.lib.gt_detect(self.h, B, conf, iou, int(bool(agnostic)), self._classes_arg(classes), boxes.ctypes.data, counts.ctypes.data,
                                    _ptr(keep), stream))
        return (boxes, counts, keep) if want_keep else (boxes, counts)

    def candidate_counts(self, B: int) -> np.ndarray:
        """Candidates per frame that passed the confidence / class filter in the last decode (before NMS)."""
        out = np.zeros((B,), np.int32)
        self._ck(self.lib.gt_get_candidate_counts(self.h, B, out.ctypes.data))
        return out

    def raw_head(self, B: int) -> np.ndarray:
        out = np.empty((B, self.A, self.no), np.float32)
        self._ck(self.lib.gt_get_raw_head(self.h, B, out.ctypes.data, None, None))
        return out

    def feature(self, layer: int, B: int) -> np.ndarray:
        c, h, w = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.gt_get_feature(self.h, layer, 0, None, C.byref(c), C.byref(h), C.byref(w)))
        out = np.empty((B, h.value, w.value, c.value), np.uint16)
        self._ck(self.lib.gt_get_feature(self.h, layer, B, out.ctypes.data, None, None, None))
        return out

    def nms(self, pred: np.ndarray, nc: int, rotated=False, conf=0.25, iou=0.7, agnostic=True, classes=None, max_det=None):
        pred = np.ascontiguousarray(pred, dtype=np.float32)
        B, A = pred.shape[0], pred.shape[1]
        md = max_det or self.max_det
        row = 7 if rotated else 6
        rows = np.zeros((B, md, row), np.float32)
        counts = np.zeros((B,), np.int32)
        keep = np.zeros((B, md), np.int32)
        self._ck(self.lib.gt_nms(self.h, pred.ctypes.data, B, A, nc, int(rotated), conf, iou, int(bool(agnostic)), self._classes_arg(classes), md,
                                 rows.ctypes.data, counts.ctypes.data, keep.ctypes.data, None))
        return rows, counts, keep

    def conv2d(self, x_bf16: np.ndarray, w: np.ndarray, bias: Optional[np.ndarray], k: int, stride: int, act: bool,
               residual: Optional[np.ndarray] = None, out_f32: bool = False) -> np.ndarray:
        B, H, W, cin = x_bf16.shape
        cout = w.shape[0]
        pad = k // 2
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        out = np.empty((B, Ho, Wo, cout), np.float32 if out_f32 else np.uint16)
        w = np.ascontiguousarray(w, np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, np.float32)
        self._ck(self.lib.gt_conv2d(self.h, _ptr(np.ascontiguousarray(x_bf16)), B, H, W, cin, w.ctypes.data, _ptr(b), cout, k, stride, int(act),
                                    _ptr(residual), out.ctypes.data, int(out_f32), None))
        return out

    # -- stabilizer --------------------------------------------------------------------------------------------------------
    def set_reference(self, slot: int, boxes: Optional[np.ndarray] = None):
        n = 0 if boxes is None else len(boxes)
        b = None if n == 0 else np.ascontiguousarray(boxes, np.float32)
        self._ck(self.lib.gt_set_reference(self.h, slot, _ptr(b), n, None))

    def stabilize(self, B: int, boxes: Optional[Sequence[Optional[np.ndarray]]] = None, stream=None):
        md = self.max_det
        bx = np.zeros((B, md, 4), np.float32)
        nb = np.zeros((B,), np.int32)
        if boxes is not None:
            for i, b in enumerate(boxes):
                if b is not None and len(b):
                    nb[i] = min(len(b), md)
                    bx[i, : nb[i]] = np.asarray(b, np.float32)[: nb[i]]
        H = np.zeros((B, 9), np.float64)
        status = np.zeros((B,), np.int32)
        stats = np.zeros((B, 4), np.int32)
        self._ck(self.lib.gt_stabilize(self.h, B, bx.ctypes.data, nb.ctypes.data, md, H.ctypes.data, status.ctypes.data, stats.ctypes.data, stream))
        return H.reshape(B, 3, 3), status, stats

    def warp_boxes(self, H: np.ndarray, boxes_xywh: np.ndarray) -> np.ndarray:
        b = np.ascontiguousarray(boxes_xywh, np.float32).copy()
        Hc = np.ascontiguousarray(H, np.float64)
        self._ck(self.lib.gt_warp_boxes(self.h, Hc.ctypes.data, b.ctypes.data, len(b), None))
        return b

    def warp_frames(self, frames, H, out=None):
        """``cv2.warpPerspective(frame, H, (w, h))`` for a batch of BGR frames (numpy or CUDA tensor), bit-exact: the stabilised view the
        reference's visualisation renders (/root/reference/geotrax/visualize.py:285-289).  H: (B, 3, 3) / (B, 9) f64 current -> reference."""
        B = int(frames.shape[0])
        Hc = np.ascontiguousarray(np.asarray(H, np.float64).reshape(B, 9))
        if out is None:
            out = np.empty((B, self.cfg.frame_h, self.cfg.frame_w, 3), np.uint8)
        self._ck(self.lib.gt_warp_frames(self.h, _ptr(frames), Hc.ctypes.data, B, _ptr(out), None))
        return out

    def orb_level_info(self):
        out = []
        for l in range(GT_ORB_LEVELS):
            w, h, qc, qr = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            self._ck(self.lib.gt_orb_level_info(self.h, l, C.byref(w), C.byref(h), C.byref(qc), C.byref(qr)))
            out.append((w.value, h.value, qc.value, qr.value))
        return out

    def pyramid_level(self, which: int, b: int, level: int):
        w, h, _, _ = self.orb_level_info()[level]
        img, msk = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
        self._ck(self.lib.gt_get_pyramid_level(self.h, which, b, level, img.ctypes.data, msk.ctypes.data))
        return img, msk

    def keypoints(self, which: int, b: int):
        kp = np.zeros((GT_MAX_KP, 6), np.float32)
        desc = np.zeros((GT_MAX_KP, 32), np.uint8)
        n = C.c_int32()
        self._ck(self.lib.gt_get_keypoints(self.h, which, b, GT_MAX_KP, kp.ctypes.data, desc.ctypes.data, C.byref(n)))
        return kp[: n.value].copy(), desc[: n.value].copy()

    def fast_candidates(self, which: int, b: int, level: int):
        cap = 1 << 18
        xy = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.uint8)
        n = C.c_int32()
        self._ck(self.lib.gt_orb_get_candidates(self.h, which, b, level, cap, xy.ctypes.data, sc.ctypes.data, C.byref(n)))
        m = min(n.value, cap)
        return (xy[:m] & 0xFFFF).astype(np.int32), (xy[:m] >> 16).astype(np.int32), sc[:m].astype(np.int32)

    def orb_detect(self, gray: np.ndarray, mask: Optional[np.ndarray] = None, as_reference=False):
        gray = np.ascontiguousarray(gray, np.uint8)
        if gray.ndim == 2:
            gray = gray[None]
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        self._ck(self.lib.gt_orb_detect(self.h, gray.ctypes.data, _ptr(mask), gray.shape[0], int(as_reference), None))

    def match(self, query: np.ndarray, train: np.ndarray):
        q, t = np.ascontiguousarray(query, np.uint8), np.ascontiguousarray(train, np.uint8)
        idx = np.zeros((len(q), 2), np.int32)
        dist = np.zeros((len(q), 2), np.int32)
        self._ck(self.lib.gt_match(self.h, q.ctypes.data, len(q), t.ctypes.data, len(t), idx.ctypes.data, dist.ctypes.data, None))
        return idx, dist

    def match_l2(self, query: np.ndarray, train: np.ndarray):
        """Brute-force L2 2-NN of float descriptors (SIFT / RootSIFT, 128 elements): (idx [nq, 2] i32, dist [nq, 2] f32)."""
        q, t = np.ascontiguousarray(query, np.float32), np.ascontiguousarray(train, np.float32)
        idx = np.zeros((len(q), 2), np.int32)
        dist = np.zeros((len(q), 2), np.float32)
        self._ck(self.lib.gt_match_l2(self.h, q.ctypes.data, len(q), t.ctypes.data, len(t), q.shape[1], idx.ctypes.data, dist.ctypes.data, None))
        return idx, dist

    def find_homography(self, src: np.ndarray, dst: np.ndarray, thr=2.0, max_iter=5000):
        s, d = np.ascontiguousarray(src, np.float32), np.ascontiguousarray(dst, np.float32)
        H = np.zeros(9, np.float64)
        inl = C.c_int32()
        rc = self._ck(self.lib.gt_find_homography(self.h, s.ctypes.data, d.ctypes.data, len(s), thr, max_iter, H.ctypes.data, C.byref(inl), None))
        return (None if rc != 0 else H.reshape(3, 3)), inl.value

    # -- fused batch -------------------------------------------------------------------------------------------------------
    def alloc_outputs(self, pinned=False):
        """Output buffers for extract_batch (optionally pinned torch tensors for async D2H)."""
        B, md = self.max_batch, self.max_det
        shapes = dict(boxes=((B, md, self.row), np.float32), counts=((B,), np.int32), boxes_stab=((B, md, 4), np.float32),
                      H=((B, 9), np.float64), status=((B,), np.int32), stats=((B, 4), np.int32))
        if pinned:
            import torch
            tmap = {np.float32: torch.float32, np.int32: torch.int32, np.float64: torch.float64}
            ts = [torch.zeros(s, dtype=tmap[d]).pin_memory() for s, d in shapes.values()]
            self._keep.extend(ts)
            return {k: t.numpy() for k, t in zip(shapes, ts)}
        return {k: np.zeros(s, d) for k, (s, d) in shapes.items()}

    def pack_boxes(self, boxes: Sequence[Optional[np.ndarray]]):
        """list of per-frame (n,4) xywh -> ([B][max_det][4] f32, [B] i32) buffers for mask_boxes."""
        B, md = len(boxes), self.max_det
        bx, nb = np.zeros((B, md, 4), np.float32), np.zeros((B,), np.int32)
        for i, b in enumerate(boxes):
            if b is not None and len(b):
                nb[i] = min(len(b), md)
                bx[i, : nb[i]] = np.asarray(b, np.float32)[: nb[i]]
        return bx, nb

    def extract_batch(self, frames, first_is_reference=False, conf=0.25, iou=0.7, agnostic=True, classes=None, out=None, stream=None,
                      mask_boxes=None, sync=True):
        """mask_boxes: None (mask = own detections) or the (bx, nb) pair from pack_boxes() (numpy or device tensors).
        sync=False: pipelined form (gt_extract_batch_async) -- returns (outputs, ticket) at once; the outputs (allocate them with
        alloc_outputs(pinned=True)) are valid after wait(ticket); at most two tickets in flight."""
        B = int(frames.shape[0])
        o = out or self.alloc_outputs()
        mb, mn = (mask_boxes if mask_boxes is not None else (None, None))
        if not sync:
            self._keep_async = (frames, mb, mn)
            t = C.c_int32()
            self._ck(self.lib.gt_extract_batch_async(self.h, _ptr(frames), B, int(first_is_reference), conf, iou, int(bool(agnostic)),
                                                     self._classes_arg(classes), _ptr(mb), _ptr(mn), self.max_det,
                                                     o["boxes"].ctypes.data, o["counts"].ctypes.data, o["boxes_stab"].ctypes.data, o["H"].ctypes.data,
                                                     o["status"].ctypes.data, o["stats"].ctypes.data, stream, C.byref(t)))
            return o, t.value
        self._ck(self.lib.gt_extract_batch(self.h, _ptr(frames), B, int(first_is_reference), conf, iou, int(bool(agnostic)), self._classes_arg(classes),
                                           _ptr(mb), _ptr(mn), self.max_det,
                                           o["boxes"].ctypes.data, o["counts"].ctypes.data, o["boxes_stab"].ctypes.data, o["H"].ctypes.data,
                                           o["status"].ctypes.data, o["stats"].ctypes.data, stream))
        return o

    def wait(self, ticket: int):
        """Block until the batch enqueued under `ticket` (extract_batch(sync=False)) is complete; its outputs are then valid."""
        self._ck(self.lib.gt_wait(self.h, int(ticket)))

    # -- 16-bit activation helpers (bit patterns <-> float32) -----------------------------------------------------------------
    def act_to_f32(self, bits: np.ndarray) -> np.ndarray:
        if self.act_dtype == "fp16":
            return bits.view(np.float16).astype(np.float32)
        return (bits.astype(np.uint32) << 16).view(np.float32)

    def f32_to_act(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        if self.act_dtype == "fp16":
            return x.astype(np.float16).view(np.uint16)
        u = x.view(np.uint32).astype(np.uint64)
        return ((u + (((u >> 16) & 1) + 0x7FFF)) >> 16).astype(np.uint16)

    def stage_times(self):
        ms = (C.c_float * 4)()
        self._ck(self.lib.gt_stage_times(self.h, ms))
        return dict(preprocess=ms[0], inference=ms[1], postprocess=ms[2], stabilize=ms[3])

    def launch_count(self) -> int:
        return int(self.lib.gt_launch_count(self.h))

    def conv_kernel_info(self):
        """-> (fused conv launches per forward, how many run the swapped-operand kernel after autotune)."""
        n, ns = C.c_int32(), C.c_int32()
        self._ck(self.lib.gt_conv_kernel_info(self.h, C.byref(n), C.byref(ns)))
        return n.value, ns.value

    def conv_pair_count(self) -> int:
        """Fused conv launches that run as CTA pairs (conv variant 6)."""
        return int(self.lib.gt_conv_pair_count(self.h))

    def conv_stack_stats(self):
        ms, fl = C.c_float(), C.c_double()
        self._ck(self.lib.gt_conv_stack_stats(self.h, C.byref(ms), C.byref(fl)))
        return ms.value, fl.value


class DeviceFrames:
    """A run of dense frames in device memory owned by someone else (the NVDEC ring): quacks like a contiguous CUDA tensor for
    ``Engine.preprocess`` / ``extract_batch`` (``shape``, ``data_ptr()``), and exposes ``__cuda_array_interface__`` so that
    ``torch.as_tensor(frames, device="cuda")`` aliases it without a copy."""

    is_cuda = True

    def __init__(self, ptr: int, shape: Tuple[int, ...]):
        self.ptr, self.shape = int(ptr), tuple(int(x) for x in shape)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "|u1", "data": (self.ptr, False), "version": 2, "strides": None}

    def data_ptr(self) -> int:
        return self.ptr

    def is_contiguous(self) -> bool:
        return True

    def __len__(self) -> int:
        return self.shape[0]


class Decoder:
    """NVDEC ingest for one engine (gt_decoder_*): H.264 / HEVC Annex-B bytes in, dense NV12 frames in HBM out.

        dec = Decoder(engine, "h264"); engine.set_input_format("nv12")
        dec.feed(chunk)                       # any number of bytes; feed(None) flushes at end of stream
        while dec.pending():
            frames = dec.take(engine.max_batch)        # DeviceFrames (n, H * 3 / 2, W): pass straight to engine.extract_batch
    Replaces the reference's ``reader.read()`` (/root/reference/geotrax/extract.py:146, 248).  Raises ``GtError`` when libnvcuvid or
    an NVDEC engine is not available: there is no software decoder behind this."""

    def __init__(self, engine: "Engine", codec: str = "h264", capacity_frames: int = 0):
        self.eng, self.lib = engine, engine.lib
        if not self.lib.gt_nvdec_available():
            raise GtError("libnvcuvid.so.1 (the NVDEC driver library) is not available on this machine")
        self.h = C.c_void_p()
        code = {"h264": _lib.GT_CODEC_H264, "hevc": _lib.GT_CODEC_HEVC, "h265": _lib.GT_CODEC_HEVC}[codec.lower()]
        rc = self.lib.gt_decoder_create(engine.h, code, int(capacity_frames), C.byref(self.h))
        if rc != 0:
            raise GtError(f"gt_decoder_create failed ({rc}): {self.lib.gt_last_error(engine.h).decode()}")
        self.frame_shape = (engine.cfg.frame_h * 3 // 2, engine.cfg.frame_w)

    def feed(self, data) -> None:
        if data is None or len(data) == 0:
            rc = self.lib.gt_decoder_feed(self.h, None, 0)
        else:
            buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
            rc = self.lib.gt_decoder_feed(self.h, buf.ctypes.data, buf.size)
        if rc != 0:
            raise GtError(f"gt_decoder_feed failed ({rc}): {self.lib.gt_decoder_last_error(self.h).decode()}")

    def pending(self) -> int:
        return int(self.lib.gt_decoder_pending(self.h))

    def take(self, max_frames: int) -> Optional[DeviceFrames]:
        ptr, n = C.c_void_p(), C.c_int32()
        rc = self.lib.gt_decoder_take(self.h, int(max_frames), C.byref(ptr), C.byref(n))
        if rc != 0:
            raise GtError(f"gt_decoder_take failed ({rc})")
        return DeviceFrames(ptr.value, (n.value,) + self.frame_shape) if n.value > 0 else None

    def close(self) -> None:
        if getattr(self, "h", None) and self.h.value:
            self.lib.gt_decoder_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
