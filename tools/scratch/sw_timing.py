"""Debug: per-phase clock64 totals of conv_sw_kernel (library built with -DGT_SW_TIMING).  python tools/scratch/sw_timing.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import geotrax_b200
def run(swap, B, H, W, cin, cout, k, res):
    os.environ["GT_SWAP"] = str(swap)
    eng = geotrax_b200.Engine(frame_hw=(512, 768), imgsz=384, nc=4, max_batch=2, act_dtype="fp16")
    lib = eng.lib
    x = (np.random.randn(B, H, W, cin).astype(np.float32))
    xb = eng.f32_to_act(x)
    w = (np.random.randn(cout, cin, k, k) / (cin * k * k) ** 0.5).astype(np.float32)
    b = np.zeros(cout, np.float32)
    r = eng.f32_to_act(np.random.randn(B, H, W, cout).astype(np.float32)) if res else None
    buf = (ctypes.c_ulonglong * 128)()
    eng.conv2d(xb, w, b, k, 1, True, r)
    lib.gt_debug_sw_timing(buf)
    eng.conv2d(xb, w, b, k, 1, True, r)
    lib.gt_debug_sw_timing(buf)
    a = np.array(buf[:], dtype=np.float64).reshape(16, 8)
    nsm = 148
    tiles = a[2:10, 6].mean() / 1.0
    print(f"swap={swap} {cin}->{cout} k{k} {H}x{W} B{B} res={res}: tiles(all CTAs)={a[2,6]:.0f}")
    print("  MMA warp per tile: tempty wait %.0f  full wait %.0f  total %.0f issue %.0f" % (a[1, 0] / a[1, 6], a[1, 1] / a[1, 6], a[1, 2] / a[1, 6], a[1, 3] / a[1, 6]))
    for wv in range(2, 10):
        n = a[wv, 6]
        print(f"  epi warp {wv} (q={wv&3} h={(wv-2)>>2}) per tile: pre {a[wv,0]/n:.0f} tfull-wait {a[wv,1]/n:.0f} A {a[wv,2]/n:.0f} bar1 {a[wv,3]/n:.0f} B {a[wv,4]/n:.0f} fence+bar2 {a[wv,5]/n:.0f}")
    eng.close()
run(1, 4, 272, 480, 32, 32, 3, False)
run(2, 4, 272, 480, 32, 32, 3, False)
run(1, 4, 136, 240, 64, 64, 3, False)
run(2, 4, 136, 240, 64, 64, 3, False)
run(1, 4, 68, 120, 128, 128, 3, False)
