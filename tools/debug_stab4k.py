import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import geotrax_b200
from geotrax_b200 import synth, weights
eng = geotrax_b200.Engine(frame_hw=(2160, 3840), imgsz=1920, nc=4, max_batch=16)
sd = weights.random_state_dict(4, "detect", seed=0, cls_bias=-4.4)
eng.load_weights(weights.fold(sd))
frames, boxes, Hs = synth.make_flight(16, 2160, 3840, seed=100)
fr = np.stack(frames)
o = eng.extract_batch(fr[:1], first_is_reference=True, classes=[0,1,2,3], mask_boxes=eng.pack_boxes(boxes[:1]))
print("ref: status", o["status"][:1], "stats", o["stats"][:1].tolist(), "H", o["H"][0])
o = eng.extract_batch(fr, classes=[0,1,2,3], mask_boxes=eng.pack_boxes(boxes))
for i in range(16):
    H = o["H"][i].reshape(3,3)
    pts = np.array([[100,100,1],[3700,100,1],[1920,1080,1],[100,2000,1],[3700,2000,1.0]])
    a, b = pts @ H.T, pts @ Hs[i].T
    err = np.linalg.norm(a[:, :2]/a[:, 2:] - b[:, :2]/b[:, 2:], axis=1).mean()
    print(i, "status", int(o["status"][i]), "stats", o["stats"][i].tolist(), "dets", int(o["counts"][i]), "err vs truth %.3f" % err)
print(eng.stage_times())
