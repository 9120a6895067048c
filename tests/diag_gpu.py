"""Scratch diagnostics run by hand on the GPU box (`python tests/diag_gpu.py detector|...`): CUDA path vs the oracle with verbose output.
Lives under tests/ because it imports oracle/ (test infrastructure); pytest does not collect it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch, cv2
import geotrax_b200
from geotrax_b200 import weights, synth
from conftest import bf16_bits_to_f32
from oracle import prepost
from oracle.yolov8 import YOLOv8

def detector():
    hw, imgsz = (512, 768), 384
    eng = geotrax_b200.Engine(frame_hw=hw, imgsz=imgsz, nc=4, max_batch=2, max_det=300, max_features=500)
    sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=hw, imgsz=imgsz, cls_bias=-4.0)
    eng.load_weights(weights.fold(sd))
    frames = np.stack(synth.make_flight(2, 512, 768, 3, n_vehicles=20)[0])
    eng.preprocess(frames)
    boxes, counts, keep = eng.detect(2, conf=0.05, want_keep=True)
    m = YOLOv8(4).eval(); m.load_state_dict(sd, strict=False)
    taps = {}
    with torch.no_grad():
        dec, ref = m(prepost.preprocess(list(frames), imgsz), taps)
    for layer in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 18, 19, 21):
        got = eng.act_to_f32(eng.feature(layer, 2))
        r = taps[str(layer)].permute(0, 2, 3, 1).numpy() if str(layer) in taps else None
        if r is None: continue
        print("layer", layer, got.shape, "rel", np.linalg.norm(got - r) / np.linalg.norm(r), "absmax ref", np.abs(r).max(), "max err", np.abs(got - r).max())
    raw = eng.raw_head(2); rr = ref.permute(0, 2, 1).numpy()
    print("raw rel", np.linalg.norm(raw - rr) / np.linalg.norm(rr), "box part", np.linalg.norm(raw[..., :64] - rr[..., :64]) / np.linalg.norm(rr[..., :64]),
          "cls part", np.linalg.norm(raw[..., 64:] - rr[..., 64:]) / np.linalg.norm(rr[..., 64:]))
    for lvl, (o, n) in enumerate(((0, 1536), (1536, 384), (1920, 96))):
        a, b = raw[:, o:o + n], rr[:, o:o + n]
        print(" level", lvl, "rel", np.linalg.norm(a - b) / np.linalg.norm(b))
    # decode comparison on identical raw
    rawt = torch.from_numpy(raw).permute(0, 2, 1).contiguous()
    head = YOLOv8(4).model[22]
    d2 = head.decode(rawt, [(32, 48), (16, 24), (8, 12)])
    outs, idxs = prepost.non_max_suppression(d2, 0.05, 0.7, [0, 1, 2, 3], True, 300, nc=4, return_idxs=True)
    for b in range(2):
        n = int(counts[b]); print("img", b, "gpu kept", n, "oracle kept", len(outs[b]))
        print(" gpu keep[:8]", keep[b, :8], "\n ora keep[:8]", idxs[b][:8].numpy())
        print(" gpu rows[:3]\n", boxes[b, :3], "\n ora rows[:3] (letterbox px)\n", outs[b][:3].numpy())
        a = int(keep[b, 0]); print(" raw cls logits at anchor", a, raw[b, a, 64:])

def orb():
    eng = geotrax_b200.Engine(frame_hw=(1080, 1920), imgsz=960, nc=4, max_batch=2, max_det=300, max_features=2000)
    fr = synth.make_flight(2, 1080, 1920, seed=11, n_vehicles=40)[0]
    g = cv2.cvtColor(fr[1], cv2.COLOR_BGR2GRAY); g = cv2.resize(g, (960, 540), interpolation=cv2.INTER_LINEAR)
    eng.orb_detect(g[None], None)
    kp, desc = eng.keypoints(0, 0)
    ref_kp, ref_desc = cv2.ORB_create(nfeatures=2000).detectAndCompute(g, None)
    info = eng.orb_level_info()
    prev = g
    for lvl, (w, h, qc, qr) in enumerate(info):
        if lvl: prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR_EXACT)
        fast = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(prev, None)
        fin = [(int(k.pt[0]), int(k.pt[1]), k.response) for k in fast if 31 <= k.pt[0] < w - 31 and 31 <= k.pt[1] < h - 31]
        sc = np.float32(1.2) ** 0; s = float(np.float32(pow(1.2, lvl)))
        mine = kp[kp[:, 5] == lvl]
        mine_xy = set((int(round(x / s)), int(round(y / s))) for x, y in mine[:, :2])
        ref_l = [k for k in ref_kp if k.octave == lvl]
        ref_xy = set((int(round(k.pt[0] / s)), int(round(k.pt[1] / s))) for k in ref_l)
        fast_xy = set((x, y) for x, y, _ in fin)
        cx, cy, csc = eng.fast_candidates(0, 0, lvl)
        cand = {(int(a), int(b)): int(c) for a, b, c in zip(cx, cy, csc)}
        fd = {(x, y): int(r) for x, y, r in fin}
        both = set(cand) & set(fd)
        print(f"   gpu FAST candidates {len(cand)} (dups {len(cx) - len(cand)}) cv2 {len(fd)} common {len(both)} score mismatches {sum(cand[p] != fd[p] for p in both)}",
              "only gpu", sorted(set(cand) - set(fd))[:4], "only cv", sorted(set(fd) - set(cand))[:4])
        if lvl == 0:
            print("    sample (pos, gpu, cv):", [(p, cand[p], fd[p]) for p in sorted(both)[:12]])
            import collections
            print("    diff histogram gpu-cv:", sorted(collections.Counter(cand[p] - fd[p] for p in both).items())[:20])
        print(f"level {lvl} {w}x{h} quota {qc}: cv2 FAST in-border {len(fin)} | gpu kept {len(mine)} cv2 kept {len(ref_l)} common {len(mine_xy & ref_xy)} | gpu in FAST set {len(mine_xy & fast_xy)}")
        if len(mine):
            rs = {(int(round(k.pt[0] / s)), int(round(k.pt[1] / s))): k for k in ref_l}
            shown = 0
            for r in mine:
                key = (int(round(r[0] / s)), int(round(r[1] / s)))
                if key in rs and shown < 3:
                    k = rs[key]; print("   ", key, "resp gpu", r[4], "cv", k.response, "angle gpu", r[3], "cv", k.angle); shown += 1
            print("    gpu resp range", mine[:, 4].min(), mine[:, 4].max(), "cv resp range", min(k.response for k in ref_l), max(k.response for k in ref_l))

if __name__ == "__main__":
    which = sys.argv[1:] or ["detector", "orb"]
    if "detector" in which: detector()
    if "orb" in which: orb()
