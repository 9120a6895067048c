"""Experiment: does alternating batches between TWO engines (two handles, two streams, separate workspaces) let the conv stack of batch
i+1 fill the SMs while batch i is in its low-occupancy stabiliser tail?  Prints frames/s for 1 engine (two tickets in flight) and 2 engines."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import geotrax_b200
from geotrax_b200 import synth, weights
HW, IMGSZ, B, STEPS = (2160, 3840), 1920, 16, 40
sd = weights.fold(weights.random_state_dict(4, "detect", seed=0, frame_hw=HW, imgsz=IMGSZ, cls_bias=-4.4))
fl = synth.make_flight(B, HW[0], HW[1], seed=100)
frames = torch.from_numpy(np.stack(fl[0])).cuda()
def make():
    e = geotrax_b200.Engine(frame_hw=HW, imgsz=IMGSZ, nc=4, max_batch=B)
    e.load_weights(sd)
    e.extract_batch(frames[:1], first_is_reference=True, classes=[0, 1, 2, 3], mask_boxes=e.pack_boxes(fl[1][:1]))
    return e
engs = [make(), make()]
masks = [tuple(torch.from_numpy(a).cuda() for a in e.pack_boxes(fl[1])) for e in engs]
outs = [[e.alloc_outputs(pinned=True), e.alloc_outputs(pinned=True)] for e in engs]
def run(n_eng):
    pend = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(STEPS):
        k = i % n_eng
        e = engs[k]
        o, t = e.extract_batch(frames, classes=[0, 1, 2, 3], out=outs[k][(i // n_eng) % 2], mask_boxes=masks[k], sync=False)
        pend.append((e, t))
        if len(pend) > (2 if n_eng == 1 else 2):
            pe, pt = pend.pop(0); pe.wait(pt)
    for pe, pt in pend: pe.wait(pt)
    torch.cuda.synchronize()
    return STEPS * B / (time.perf_counter() - t0)
for n in (1, 2, 1, 2):
    run(n)
    print(f"{n} engine(s): {run(n):.1f} frames/s")
