"""Raw-head error vs the fp32 oracle and conv-stack time at 3840x2160 for the current GT_SILU_TANH_PX setting (one-MUFU SiLU experiment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # diagnostic (uses the oracle as the checker, hence under tests/)
import numpy as np, torch
import geotrax_b200
from geotrax_b200 import synth, weights
from oracle import prepost
from oracle.yolov8 import YOLOv8
HW, IMGSZ, B = (2160, 3840), 1920, 16
eng = geotrax_b200.Engine(frame_hw=HW, imgsz=IMGSZ, nc=4, max_batch=B)
sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=HW, imgsz=IMGSZ, cls_bias=-4.4)
eng.load_weights(weights.fold(sd))
fl = synth.make_flight(3, HW[0], HW[1], seed=100)[0]
frames = np.stack([fl[1 + (i % 2)] for i in range(B)])
dev = torch.from_numpy(frames).cuda()
eng.preprocess(dev)
ms = []
for _ in range(6):
    eng.detect(B, conf=0.25)
    ms.append(eng.conv_stack_stats()[0])
raw = eng.raw_head(2)
m = YOLOv8(4).eval(); m.load_state_dict(sd, strict=False)
rels = []
with torch.no_grad():
    for i in range(2):
        _, ref = m(prepost.preprocess([frames[i]], IMGSZ))
        ref = ref.permute(0, 2, 1).numpy()[0]
        rels.append(float(np.linalg.norm(raw[i] - ref) / np.linalg.norm(ref)))
print(f"GT_SILU_TANH_PX={os.environ.get('GT_SILU_TANH_PX', '0')}: raw head rel L2 {rels[0]:.3e} {rels[1]:.3e}; conv stack {np.median(ms[2:]):.3f} ms / 16 frames")
