"""Runs a few steps of the 4K hot path for ncu (launch list / full capture).

    python tools/profile_step.py [steps] [dtype]
    ncu --profile-from-start off ... python tools/profile_step.py 1      # only the LAST step is inside cudaProfilerStart/Stop

The per-layer conv kernels come from the shipped table (csrc/conv_tune.inc): a profiled run uses exactly the kernels of a normal run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import geotrax_b200
from geotrax_b200 import synth, weights
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp16"
B = 16
eng = geotrax_b200.Engine(frame_hw=(2160, 3840), imgsz=1920, nc=4, max_batch=B, act_dtype=dtype)
eng.load_weights(weights.fold(weights.random_state_dict(4, "detect", seed=0, cls_bias=-4.4)))
fl = synth.make_flight(4, 2160, 3840, seed=100)
frames = np.stack([fl[0][i % 4] for i in range(B)])
boxes = [fl[1][i % 4] for i in range(B)]
dev = torch.from_numpy(frames).cuda()
mask = eng.pack_boxes(boxes)
eng.extract_batch(dev[:1], first_is_reference=True, classes=[0, 1, 2, 3], mask_boxes=eng.pack_boxes(boxes[:1]))
out = eng.alloc_outputs()
for i in range(steps + 1):
    if i == steps:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    eng.extract_batch(dev, classes=[0, 1, 2, 3], mask_boxes=mask, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", eng.launch_count(), eng.stage_times())
