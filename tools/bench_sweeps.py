#!/usr/bin/env python
"""Measurements SURVEY.md section 8(d) asks for beyond the default bench line (written as JSON lines; run on the GPU box):

  config 2 sweep : decode + NMS time at ~150 / ~1,000 / 30,000 (max_nms) candidates per frame x {agnostic, class-aware}
                   (the > 4096-candidate second pass included), batch 16, 3840x2160
  config 3 stress: stabilise-only with 132 and with 1,000 mask boxes per frame; and the `stable` preset (CLAHE, full-res, 4000 / 8000)

    python tools/bench_sweeps.py > profiles/sweeps_r2.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geotrax_b200  # noqa: E402
from geotrax_b200 import synth, weights  # noqa: E402

HW, IMGSZ, B = (2160, 3840), 1920, 16


def time_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    flight = synth.make_flight(B, HW[0], HW[1], seed=100)
    frames = torch.from_numpy(np.stack(flight[0])).cuda()
    sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=HW, imgsz=IMGSZ, cls_bias=-4.4)
    eng = geotrax_b200.Engine(frame_hw=HW, imgsz=IMGSZ, nc=4, max_batch=B)
    eng.load_weights(weights.fold(sd))
    eng.preprocess(frames)
    # ---- config 2: candidates / frame set by the confidence threshold on the same raw head
    for conf in (0.30, 0.25, 0.12, 0.02, 0.0005):
        for agnostic in (True, False):
            post = []
            def run():
                eng.detect(B, conf=conf, iou=0.7, agnostic=agnostic, classes=[0, 1, 2, 3])
                post.append(eng.stage_times()["postprocess"])
            ms = time_ms(run)
            bx, cnt = eng.detect(B, conf=conf, iou=0.7, agnostic=agnostic, classes=[0, 1, 2, 3])
            cand = eng.candidate_counts(B)
            print(json.dumps(dict(sweep="config2", conf=conf, agnostic=agnostic, candidates_per_frame=float(cand.mean()), candidates_max=int(cand.max()),
                                  kept_per_frame=float(cnt.mean()), detect_ms_per_step=ms, postprocess_ms_per_step=float(np.median(post[-5:])),
                                  second_pass_frames=int((np.minimum(cand, 30000) > 4096).sum()))), flush=True)
    # ---- config 3: mask stress
    rng = np.random.default_rng(0)
    for nbox in (132, 1000):
        boxes = flight[1] if nbox == 132 else [synth.make_boxes(nbox, HW[0], HW[1], rng) for _ in range(B)]
        eng.preprocess(frames[:1])
        eng.set_reference(0, boxes[0])
        eng.preprocess(frames)
        ms = time_ms(lambda: eng.stabilize(B, boxes))
        H, status, stats = eng.stabilize(B, boxes)
        print(json.dumps(dict(sweep="config3", preset="default", mask_boxes_per_frame=nbox, stabilize_ms_per_step=ms, frames_per_s=B / ms * 1e3,
                              homographies_ok=int((status == 0).sum()), matches=float(stats[:, 2].mean()), inliers=float(stats[:, 3].mean()))), flush=True)
    eng.close()
    # ---- config 3 heavy case: the `stable` preset
    eng = geotrax_b200.Engine(frame_hw=HW, imgsz=IMGSZ, nc=4, max_batch=B, clahe=1, downsample_ratio=1.0, max_features=4000, filter_ratio=0.8)
    boxes = flight[1]
    eng.preprocess(frames[:1])
    eng.set_reference(0, boxes[0])
    def run():
        eng.preprocess(frames)
        eng.stabilize(B, boxes)
    ms = time_ms(run, reps=3, warm=1)
    H, status, stats = eng.stabilize(B, boxes)
    print(json.dumps(dict(sweep="config3", preset="stable (CLAHE, full-res working image, 4000 / 8000 key points, ratio 0.8)", mask_boxes_per_frame=132,
                          preprocess_plus_stabilize_ms_per_step=ms, frames_per_s=B / ms * 1e3, homographies_ok=int((status == 0).sum()),
                          keypoints_ref=int(stats[0, 0]), keypoints_cur=float(stats[:, 1].mean()), matches=float(stats[:, 2].mean()),
                          inliers=float(stats[:, 3].mean()))), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
