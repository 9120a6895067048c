"""compute-sanitizer target: one small pass of every kernel family (default geometry, general geometry, stabilizer).
    compute-sanitizer --tool memcheck python tools/san_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import geotrax_b200
from geotrax_b200 import weights, synth
os.environ.setdefault("GT_SWAP", "0")     # no autotune (hundreds of launches under the sanitizer); GT_SWAP=1..5 covers the other kernels
hw, imgsz = (512, 768), 384
eng = geotrax_b200.Engine(frame_hw=hw, imgsz=imgsz, nc=4, max_batch=2, max_det=300, max_features=500)
sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=hw, imgsz=imgsz, cls_bias=-4.0)
eng.load_weights(weights.fold(sd))
fl = synth.make_flight(3, hw[0], hw[1], seed=1, n_vehicles=12)
frames = np.stack(fl[0])
out0 = eng.extract_batch(frames[:1], first_is_reference=True, conf=0.05, mask_boxes=eng.pack_boxes(fl[1][:1]))
out = eng.extract_batch(frames[1:3], conf=0.05, mask_boxes=eng.pack_boxes(fl[1][1:3]))
print("default geometry:", out["counts"].tolist(), out["status"].tolist(), out["stats"].tolist())
# stand-alone matchers: Hamming on the tensor cores (ragged sizes) and the L2 registration matcher + robust fit
rng = np.random.default_rng(3)
for nq, nt in [(130, 257), (1, 1), (700, 1999)]:
    i_, d_ = eng.match(rng.integers(0, 256, (nq, 32), dtype=np.uint8), rng.integers(0, 256, (nt, 32), dtype=np.uint8))
for nq, nt in [(3, 2), (300, 517)]:
    t_ = np.abs(rng.normal(size=(nt, 128))).astype(np.float32); q_ = np.abs(rng.normal(size=(nq, 128))).astype(np.float32)
    i2, d2 = eng.match_l2(q_, t_)
print("matchers:", i_[:2].tolist(), i2[:2].tolist())
eng.close()
# general geometry: odd frame size, bilinear letterbox and working image
hw2 = (375, 667)
eng = geotrax_b200.Engine(frame_hw=hw2, imgsz=384, nc=4, max_batch=2, max_det=100, max_features=300)
eng.load_weights(weights.fold(weights.random_state_dict(4, "detect", seed=0, frame_hw=hw2, imgsz=384, cls_bias=-4.0)))
fr = np.random.default_rng(0).integers(0, 256, (2,) + hw2 + (3,), dtype=np.uint8)
eng.extract_batch(fr[:1], first_is_reference=True, conf=0.05)
o = eng.extract_batch(fr, conf=0.05)
print("general geometry:", o["counts"].tolist(), o["status"].tolist())
eng.close()
