"""Batched, frame-range-sharded driver of the hot path and the host-side replay that turns its output
into exactly what /root/reference/geotrax/extract.py:134-214 ``track_with_model`` returns.

One process per GPU.  Rank k owns the contiguous frame range ``frame_ranges(n, world)[k]``; detection is per-frame
independent and stabilisation is frame-to-REFERENCE (extract.py:176-181), so the only shared state is the reference frame,
which every rank recomputes locally (deterministic kernels => identical features).  There is no data-path collective; one
gather of per-frame records (boxes, homography, status) to rank 0 follows, over NCCL on the GPU box (gloo in the CPU tests).
Rank 0 then replays the sequential host tracker over the frames in order and warps the tracker's boxes with each frame's H.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


def frame_ranges(n_frames: int, world: int, start: int = 0) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) ranges, sizes differing by at most one, earlier ranks get the longer ones."""
    base, rem = divmod(max(n_frames, 0), world)
    out, lo = [], start
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def run_range(engine, get_frames: Callable[[int, int], np.ndarray], lo: int, hi: int, ref_frame_index: int, batch: int = 16, conf: float = 0.25,
              iou: float = 0.7, agnostic: bool = True, classes: Optional[Sequence[int]] = None, stream=None,
              get_masks: Optional[Callable[[int, int], tuple]] = None, set_reference: bool = True, pipelined: Optional[bool] = None,
              on_batch: Optional[Callable[[int, int], None]] = None, next_range_frames=None) -> Dict[str, np.ndarray]:
    """Runs frames [lo, hi) of a flight through the engine in batches; -> per-frame arrays (see keys below).

    The frame ``ref_frame_index`` (the first processed frame of the whole video, ``cut_frame_left``) is the stabilizer
    reference: processed first on every rank (``set_reference=False`` when the engine already holds it); its own record is only
    kept by the rank that owns it.

    ``get_frames(a, b)`` returns frames [a, b) as a (n, H, W, 3) u8 array: numpy (pageable or pinned) or a CUDA tensor.  With an
    engine that has the asynchronous entry points the loop keeps TWO batches in flight (``gt_extract_batch_async`` / ``gt_wait``):
    batch i+1 is enqueued -- and, for host frames, its H2D copy started (``gt_prefetch_frames``) -- before batch i is read back, so
    the GPU never idles between batches.  ``get_masks(a, b)`` (optional) returns the ``(boxes, counts)`` pair of
    ``Engine.pack_boxes`` used as ORB vehicle masks (the tracker's boxes in the reference, extract.py:166,181); default: this
    batch's own detections.  ``on_batch(b0, b1)`` is called after each batch's outputs are valid (progress / timing hooks).
    ``engine`` may be a list of engines on the same GPU (see the comment in the pipelined loop).
    ``next_range_frames``: host frames of the batch that FOLLOWS this range (chunked / streaming ingest): their H2D copy is started
    during the last batch, so the next ``run_range`` call finds its first batch already on the device (steady-state ingest)."""
    engines = list(engine) if isinstance(engine, (list, tuple)) else [engine]
    engine = engines[0]
    md, row = engine.max_det, engine.row
    n = max(hi - lo, 0)
    res = dict(frame=np.arange(lo, hi, dtype=np.int64), count=np.zeros(n, np.int32), status=np.zeros(n, np.int32), stats=np.zeros((n, 4), np.int32),
               H=np.zeros((n, 9), np.float64), boxes=np.empty((n, md, row), np.float32), boxes_stab=np.empty((n, md, 4), np.float32))
    # (boxes / boxes_stab rows at and beyond a frame's `count` are unspecified: np.empty avoids touching ~40 KB per frame twice)
    kw = dict(conf=conf, iou=iou, agnostic=agnostic, classes=classes, stream=stream)
    ref = None
    if set_reference:
        rm = get_masks(ref_frame_index, ref_frame_index + 1) if get_masks else None
        for e in engines:        # every engine (and every rank) computes the reference features itself: deterministic kernels, identical result
            r = e.extract_batch(get_frames(ref_frame_index, ref_frame_index + 1), first_is_reference=True, mask_boxes=rm, **kw) if rm is not None \
                else e.extract_batch(get_frames(ref_frame_index, ref_frame_index + 1), first_is_reference=True, **kw)
            ref = ref or {k: v.copy() for k, v in r.items()}
    if pipelined is None:
        pipelined = hasattr(engine, "wait")
    starts = list(range(lo, hi, batch))

    def store(o, b0, b1):
        s, k = slice(b0 - lo, b1 - lo), b1 - b0
        res["count"][s], res["status"][s], res["stats"][s], res["H"][s] = o["counts"][:k], o["status"][:k], o["stats"][:k], o["H"][:k]
        res["boxes"][s], res["boxes_stab"][s] = o["boxes"][:k], o["boxes_stab"][:k]
        if on_batch:
            on_batch(b0, b1)

    if not pipelined:
        out = engine.alloc_outputs()
        for b0 in starts:
            b1 = min(b0 + batch, hi)
            mk = dict(mask_boxes=get_masks(b0, b1)) if get_masks else {}
            store(engine.extract_batch(get_frames(b0, b1), out=out, **mk, **kw), b0, b1)
    else:
        # Several engines (handles on the same GPU, each with its own workspaces and streams) take the batches round-robin: while one
        # batch is in its low-occupancy stabiliser tail (NMS, selection, RANSAC: 16-128 blocks) the other engine's convolution CTAs
        # fill the SMs.  Measured +6.8 % with two engines on the final round-2 build (3,491 -> 3,730 frames/s; bench.py runs two).
        ne = len(engines)
        for e in engines:
            if getattr(e, "_pipeline_outs", None) is None:   # two pinned output sets per engine, allocated once
                e._pipeline_outs = [e.alloc_outputs(pinned=True), e.alloc_outputs(pinned=True)]
        is_host = lambda f: isinstance(f, np.ndarray) or (hasattr(f, "is_cuda") and not f.is_cuda)
        nxt = get_frames(starts[0], min(starts[0] + batch, hi)) if starts else None
        pending = []                        # (engine, ticket, outputs, b0, b1, frames kept alive, masks kept alive)
        for i, b0 in enumerate(starts):
            b1 = min(b0 + batch, hi)
            e = engines[i % ne]
            cur = nxt
            nxt = get_frames(starts[i + 1], min(starts[i + 1] + batch, hi)) if i + 1 < len(starts) else next_range_frames
            if nxt is not None and is_host(nxt):
                en = engines[(i + 1) % ne]
                if hasattr(en, "prefetch"):
                    en.prefetch(nxt, deferred=(ne == 1))   # one engine: the copy of batch i+1 starts right after batch i's own small uploads
            mk = get_masks(b0, b1) if get_masks else None
            o, t = e.extract_batch(cur, out=e._pipeline_outs[(i // ne) % 2], mask_boxes=mk, sync=False, **kw)
            pending.append((e, t, o, b0, b1, cur, mk))
            if len(pending) > ne:
                pe, pt, po, p0, p1, _, _ = pending.pop(0)
                pe.wait(pt)
                store(po, p0, p1)
        for pe, pt, po, p0, p1, _, _ in pending:
            pe.wait(pt)
            store(po, p0, p1)
    if ref is not None and lo <= ref_frame_index < hi:   # the reference frame maps to itself: identity, boxes unchanged (extract.py:176-179)
        i = ref_frame_index - lo
        res["count"][i], res["status"][i], res["stats"][i] = ref["counts"][0], 0, ref["stats"][0]
        res["H"][i] = np.eye(3).ravel()
        res["boxes"][i], res["boxes_stab"][i] = ref["boxes"][0], ref["boxes_stab"][0]
    return res


def _record_dtype(c_max: int, row: int) -> np.dtype:
    """One frame's record as a packed structure (what travels in the gather)."""
    return np.dtype([("frame", "<i8"), ("count", "<i4"), ("status", "<i4"), ("stats", "<i4", (4,)), ("H", "<f8", (9,)),
                     ("boxes", "<f4", (c_max, row)), ("boxes_stab", "<f4", (c_max, 4))])


def gather_records(local: Dict[str, np.ndarray], rank: int, world: int, device=None) -> Optional[Dict[str, np.ndarray]]:
    """Gathers every rank's per-frame records on rank 0 (returns None elsewhere): ONE small all_gather (frames per rank, largest
    per-frame detection count) and ONE gather of packed per-frame structures -- boxes trimmed to the largest count over all ranks
    (real payload ~3-10 KB / frame instead of max_det rows), H kept f64 and frame numbers i64 because the bytes travel, not floats.
    NCCL when ``device`` is a CUDA device (the GPU box), gloo on the CPU (tests)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist

    dev = device if device is not None else torch.device("cpu")
    n_local = len(local["frame"])
    meta = torch.tensor([n_local, int(local["count"].max()) if n_local else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.cpu() for m in metas]
    n_max = max(int(m[0]) for m in metas)
    c_max = max(1, max(int(m[1]) for m in metas))
    row = local["boxes"].shape[2]
    dt = _record_dtype(c_max, row)
    rec = np.zeros(max(n_max, 1), dt)
    for key in ("frame", "count", "status", "stats", "H"):
        rec[key][:n_local] = local[key]
    rec["boxes"][:n_local] = local["boxes"][:, :c_max]
    rec["boxes_stab"][:n_local] = local["boxes_stab"][:, :c_max]
    t = torch.from_numpy(rec.view(np.uint8).reshape(len(rec), dt.itemsize)).to(dev)
    parts = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, parts, dst=0)
    if rank != 0:
        return None
    chunks = [p.cpu().numpy().reshape(-1).view(dt)[: int(metas[r][0])] for r, p in enumerate(parts)]
    allrec = np.concatenate(chunks, 0)
    # boxes / boxes_stab come back trimmed to (n, c_max, .): every consumer reads rows [:count] only
    return {key: np.ascontiguousarray(allrec[key]) for key in ("frame", "count", "status", "stats", "H", "boxes", "boxes_stab")}


class GmcFromHomography:
    """Drop-in for ``BOTSORT.gmc`` (ultralytics ``trackers/utils/gmc.py``; ``gmc_method: sparseOptFlow`` in
    /root/reference/geotrax/cfg/default.yaml:374) fed by the stabilizer's homographies instead of sparse optical flow on the frames.

    BoT-SORT's global-motion compensation wants the previous-frame -> current-frame 2x3 affine.  The hot path already has, for every
    frame, H_t : frame t -> reference; so prev -> cur = H_cur^-1 . H_prev (its perspective row is ~1e-6 on BEV drone footage and is
    dropped).  The sharded driver therefore needs no pixels on rank 0 -- SURVEY.md section 8f "tracker-side GMC input"."""

    def __init__(self):
        self.prev_H: Optional[np.ndarray] = None
        self.cur_H: Optional[np.ndarray] = None

    def set_frame(self, H: Optional[np.ndarray]) -> None:
        self.cur_H = None if H is None else np.asarray(H, np.float64).reshape(3, 3)

    def apply(self, raw_frame=None, detections=None) -> np.ndarray:
        out = np.eye(2, 3)
        if self.prev_H is not None and self.cur_H is not None:
            M = np.linalg.solve(self.cur_H, self.prev_H)      # inv(H_cur) @ H_prev
            out = (M[:2, :] / M[2, 2]).copy()
        if self.cur_H is not None:
            self.prev_H = self.cur_H
        return out

    def reset_params(self) -> None:
        self.prev_H = self.cur_H = None


def replay_tracks(rec: Dict[str, np.ndarray], tracker, warp_boxes: Callable[[np.ndarray, np.ndarray], np.ndarray], ref_frame_index: int,
                  obb: bool = False, frame_hw: Tuple[int, int] = (2160, 3840)) -> Tuple[np.ndarray, np.ndarray]:
    """Rank 0: sequential tracker over the gathered detections, then the stabilizer's box warp with the stored H.

    -> (tracks (N,12) f32 [frame,id,x,y,w,h,xs,ys,ws,hs,cls,conf], transforms (F,10) f64 [frame, H row-major]) -- the arrays
    extract.py:273-293 ``aggregate_results`` builds; untracked rows (id -1) are dropped as extract.py:287 does.

    ``tracker``: anything with ultralytics' ``update(det, img, feats) -> rows [x1,y1,x2,y2,id,score,cls,det_idx]`` contract -- the real
    BOTSORT / BYTETracker (``tracker.make_tracker``) or the stand-in.  ``det`` is a numpy-backed ``Boxes`` / ``OBB`` as in ultralytics.
    A tracker that owns a ``gmc`` object (BoT-SORT) gets it replaced by ``GmcFromHomography``: no frame pixels are needed here."""
    from .results import OBB, Boxes

    gmc = None
    if hasattr(tracker, "gmc"):
        gmc = GmcFromHomography()
        tracker.gmc = gmc
    order = np.argsort(rec["frame"], kind="stable")
    tracks, transforms = [], []
    for i in order:
        f, n = int(rec["frame"][i]), int(rec["count"][i])
        H = rec["H"][i].reshape(3, 3) if int(rec["status"][i]) == 0 else None
        if gmc is not None:
            gmc.set_frame(np.eye(3) if f == ref_frame_index else H)
        if f != ref_frame_index and H is not None:
            transforms.append(np.concatenate([[float(f)], H.ravel()])[None])
        if n == 0:
            continue
        d = np.ascontiguousarray(rec["boxes"][i, :n])
        rows = np.asarray(tracker.update(OBB(d, frame_hw) if obb else Boxes(d, frame_hw), None, None))
        if len(rows) == 0:
            continue
        if obb and rows.shape[1] == 9:       # ultralytics rows for rotated boxes: [x, y, w, h, angle, id, score, cls, det_idx]
            c, s = np.abs(np.cos(rows[:, 4])), np.abs(np.sin(rows[:, 4]))
            xywh = np.stack([rows[:, 0], rows[:, 1], rows[:, 2] * c + rows[:, 3] * s, rows[:, 2] * s + rows[:, 3] * c], 1).astype(np.float32)
            tid, score, cls = rows[:, 5], rows[:, 6], rows[:, 7]
        else:
            b = np.asarray(rows[:, :4], np.float32)
            xywh = np.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1).astype(np.float32)
            tid, score, cls = rows[:, 4], rows[:, 5], rows[:, 6]
        stab = xywh if (f == ref_frame_index or H is None) else warp_boxes(H, xywh)
        ids = tid.astype(np.uint16).astype(np.float32)          # extract.py:162 casts ids to uint16
        tracks.append(np.concatenate([np.full((len(xywh), 1), f, np.float32), ids[:, None], xywh, stab.astype(np.float32),
                                      cls[:, None].astype(np.uint8).astype(np.float32), score[:, None].astype(np.float32)], 1))
    t = np.concatenate(tracks, 0).astype(np.float32) if tracks else np.empty((0, 12), np.float32)
    t = t[t[:, 1] != -1] if t.size else t
    tr = np.concatenate(transforms, 0) if transforms else np.empty((0, 10))
    return t, tr


def run_flight(engine, get_frames: Callable[[int, int], np.ndarray], n_frames: int, rank: int = 0, world: int = 1, first_frame: int = 0,
               batch: int = 16, tracker=None, gather_device=None, **det_kw):
    """Whole sharded job: local range -> gather -> (rank 0) tracker replay.  Returns (tracks, transforms) on rank 0, else None.

    ``tracker``: a tracker object, a tracker yaml path / ``'greedy-iou'`` for ``tracker.make_tracker``, or None (ultralytics' default
    BoT-SORT; raises when ultralytics is absent -- see ``tracker.make_tracker``)."""
    from .tracker import make_tracker

    lo, hi = frame_ranges(n_frames, world, first_frame)[rank]
    local = run_range(engine, get_frames, lo, hi, first_frame, batch=batch, **det_kw)
    rec = gather_records(local, rank, world, gather_device)
    if rank != 0:
        return None
    trk = tracker if hasattr(tracker, "update") else make_tracker(tracker)
    return replay_tracks(rec, trk, engine.warp_boxes, first_frame, obb=(engine.row == 7), frame_hw=(engine.cfg.frame_h, engine.cfg.frame_w))
