"""Hardware check of conv variant 6 (CTA-pair swapped kernel, conv_sw_kernel<1>): gt_conv2d against torch on cases ordered from the
simplest (one pair, one tile, one k-block) to the general ones, then the whole detector forced onto the variant against the
single-CTA swapped kernel.  Run under `timeout` on the GPU box: `GT_SWAP=6 timeout -s KILL 90 python tools/pair_check.py`.
On a mismatch the error is broken down by 128-channel block and by image-row band so that a swapped pixel half or channel half shows."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GT_SWAP", "6")

import geotrax_b200  # noqa: E402
from geotrax_b200 import weights  # noqa: E402

CASES = [
    # (B, H, W, cin, cout, k, stride, act, residual)
    (1, 16, 16, 64, 256, 1, 1, False, False),    # one pair, one tile, one k-block, no activation
    (1, 16, 16, 256, 256, 1, 1, True, False),    # four k-blocks
    (2, 64, 96, 128, 256, 1, 1, True, False),    # 48 tiles: ring and accumulator phases wrap
    (1, 34, 60, 256, 512, 1, 1, True, False),    # two pair tiles along cout, ragged spatial tiles
    (1, 34, 60, 128, 256, 3, 1, True, True),     # 3x3, residual
    (2, 68, 120, 128, 256, 3, 2, True, False),   # stride 2
    (16, 34, 60, 256, 256, 3, 1, True, False),   # more tiles than pairs, 36 k-blocks
]


def main():
    hw, imgsz = (512, 768), 384
    eng = geotrax_b200.Engine(frame_hw=hw, imgsz=imgsz, nc=4, max_batch=16, max_det=300, max_features=500, act_dtype="fp16")
    sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=hw, imgsz=imgsz, cls_bias=-4.0)
    eng.load_weights(weights.fold(sd))
    print("engine ready, conv kernels:", eng.conv_kernel_info() if hasattr(eng, "conv_kernel_info") else "?", flush=True)
    rnd = lambda a: eng.act_to_f32(eng.f32_to_act(a))
    bad = 0
    for case in CASES:
        B, H, W, cin, cout, k, s, act, use_res = case
        g = torch.Generator().manual_seed(abs(hash(case)) % (2 ** 31))
        x = rnd(torch.randn(B, H, W, cin, generator=g).numpy())
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).numpy()
        b = torch.randn(cout, generator=g).numpy() * 0.1
        ref = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(rnd(w)), torch.from_numpy(b), s, k // 2)
        if act:
            ref = torch.nn.functional.silu(ref)
        res_bits = None
        if use_res:
            res = rnd(torch.randn(*ref.permute(0, 2, 3, 1).shape, generator=g).numpy())
            ref = ref + torch.from_numpy(res).permute(0, 3, 1, 2)
            res_bits = eng.f32_to_act(res)
        ref = ref.permute(0, 2, 3, 1).numpy()
        print("case", case, "...", end=" ", flush=True)
        got = eng.act_to_f32(eng.conv2d(eng.f32_to_act(x), w, b, k, s, act, res_bits, out_f32=False))
        err = np.abs(got - ref)
        scale = max(1e-6, np.abs(ref).max())
        e = err.max() / scale
        ok = e < 2e-3
        print("max-normalised error %.3e %s" % (e, "ok" if ok else "MISMATCH"), flush=True)
        if not ok:
            bad += 1
            per_c = [err[..., c:c + 128].max() / scale for c in range(0, cout, 128)]
            print("   by 128-channel block:", " ".join("%.2e" % v for v in per_c))
            rows = err.max(axis=(0, 2, 3)) / scale
            print("   by image row       :", " ".join("%.0e" % v for v in rows[:40]))
            cols = err.max(axis=(0, 1, 3)) / scale
            print("   by image column    :", " ".join("%.0e" % v for v in cols[:40]))
            z = (got == 0).mean()
            print("   fraction of exact zeros in the output: %.3f" % z, flush=True)
    # whole detector: every cout % 256 == 0 layer on the pair kernel vs the engine's plan on the single-CTA swapped kernel
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (2,) + hw + (3,), dtype=np.uint8)
    eng.preprocess(frames)
    eng.detect(2, conf=0.25, iou=0.7, agnostic=True, classes=[0, 1, 2, 3])
    raw6 = eng.raw_head(2).copy()
    eng.close()
    os.environ["GT_SWAP"] = "1"
    eng1 = geotrax_b200.Engine(frame_hw=hw, imgsz=imgsz, nc=4, max_batch=16, max_det=300, max_features=500, act_dtype="fp16")
    eng1.load_weights(weights.fold(sd))
    eng1.preprocess(frames)
    eng1.detect(2, conf=0.25, iou=0.7, agnostic=True, classes=[0, 1, 2, 3])
    raw1 = eng1.raw_head(2)
    rel = np.linalg.norm(raw6 - raw1) / np.linalg.norm(raw1)
    print("whole net, pair vs single-CTA swapped: rel L2 %.3e" % rel, flush=True)
    eng1.close()
    if bad or not rel < 2e-3:
        print("PAIR CHECK FAILED")
        return 1
    print("PAIR CHECK OK")
    return 0


if __name__ == "__main__":
    sys.exit(main())
