import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import geotrax_b200
from geotrax_b200 import weights, synth
hw, imgsz = (512, 768), 384
eng = geotrax_b200.Engine(frame_hw=hw, imgsz=imgsz, nc=4, max_batch=2, max_det=300, max_features=500)
sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=hw, imgsz=imgsz, cls_bias=-4.0)
eng.load_weights(weights.fold(sd))
frames = np.stack(synth.make_flight(2, hw[0], hw[1], seed=1, n_vehicles=12)[0])
eng.preprocess(frames)
print(eng.detect(2, conf=0.05)[1])
