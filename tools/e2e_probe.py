"""Probe of the end-to-end (host frames) step: copy alone, compute alone, overlapped.  python tools/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import geotrax_b200
from geotrax_b200 import synth, weights
B = 16
eng = geotrax_b200.Engine(frame_hw=(2160, 3840), imgsz=1920, nc=4, max_batch=B)
eng.load_weights(weights.fold(weights.random_state_dict(4, "detect", seed=0, cls_bias=-4.4)))
fl = synth.make_flight(4, 2160, 3840, seed=100)
frames = np.stack([fl[0][i % 4] for i in range(B)])
boxes = [fl[1][i % 4] for i in range(B)]
pin = [torch.from_numpy(frames).pin_memory(), torch.from_numpy(frames.copy()).pin_memory()]
dev = torch.from_numpy(frames).cuda()
mask = eng.pack_boxes(boxes)
mask_dev = (torch.from_numpy(mask[0]).cuda(), torch.from_numpy(mask[1]).cuda())
eng.extract_batch(dev[:1], first_is_reference=True, classes=[0, 1, 2, 3], mask_boxes=eng.pack_boxes(boxes[:1]))
out = eng.alloc_outputs(pinned=True)
def sync(): torch.cuda.synchronize()
for _ in range(3): eng.extract_batch(dev, classes=[0, 1, 2, 3], mask_boxes=mask_dev, out=out)
sync(); t=time.perf_counter()
for _ in range(5): eng.extract_batch(dev, classes=[0, 1, 2, 3], mask_boxes=mask_dev, out=out)
sync(); print("compute only ms/step", (time.perf_counter()-t)/5*1e3)
d2 = torch.empty_like(dev)
sync(); t=time.perf_counter()
for i in range(5): d2.copy_(pin[i%2], non_blocking=True)
sync(); print("torch H2D copy only ms/step", (time.perf_counter()-t)/5*1e3)
sync(); t=time.perf_counter()
for i in range(5):
    eng.prefetch(pin[i%2]); eng.preprocess(pin[i%2])
sync(); print("prefetch+preprocess ms/step", (time.perf_counter()-t)/5*1e3)
# serial host path (no prefetch)
sync(); t=time.perf_counter()
for i in range(5): eng.extract_batch(pin[i%2], classes=[0, 1, 2, 3], mask_boxes=mask_dev, out=out)
sync(); print("host frames, no prefetch ms/step", (time.perf_counter()-t)/5*1e3)
eng.prefetch(pin[0]); sync(); t=time.perf_counter()
for i in range(6):
    t0=time.perf_counter()
    if i < 5: eng.prefetch(pin[(i+1)%2])
    t1=time.perf_counter()
    eng.extract_batch(pin[i%2], classes=[0, 1, 2, 3], mask_boxes=mask_dev, out=out)
    print("  step", i, "prefetch call ms", (t1-t0)*1e3, "extract ms", (time.perf_counter()-t1)*1e3)
sync(); print("host frames, prefetch overlapped ms/step", (time.perf_counter()-t)/6*1e3)
# --- variants closer to bench.py
mask_pin = tuple(torch.from_numpy(a).pin_memory() for a in mask)
def run(label, mb, stream):
    eng.prefetch(pin[0]); sync(); t=time.perf_counter()
    for i in range(6):
        if i < 5: eng.prefetch(pin[(i+1)%2])
        eng.extract_batch(pin[i%2], classes=[0, 1, 2, 3], mask_boxes=mb, out=out, stream=stream)
    sync(); print(label, (time.perf_counter()-t)/6*1e3)
run("overlapped, pinned host masks ms/step", mask_pin, None)
run("overlapped, pageable host masks ms/step", mask, None)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
run("overlapped, device masks, torch stream ms/step", mask_dev, ts.cuda_stream)
run("overlapped, pinned masks, torch stream ms/step", mask_pin, ts.cuda_stream)
