"""f-3 measurement: gt_match_l2 (tcgen05 candidate pass + exact re-rank) on RootSIFT-like descriptors, device-resident, CUDA events;
cv2.BFMatcher(NORM_L2).knnMatch on a bounded sample of the same queries on the host cores beside it.  One JSON line per size.

    python tools/bench_registration.py [--sizes 50000,250000] [--cpu-queries 256]
"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2, numpy as np, torch
import geotrax_b200


def rootsift_like(rng, n):
    d = rng.gamma(0.6, 18.0, (n, 128)).astype(np.float32)
    d[rng.random((n, 128)) < 0.35] = 0
    d = np.minimum(np.floor(d), 255).astype(np.float32)
    d[:, 0] += 1
    d /= d.sum(1, keepdims=True) + 1e-8
    return np.sqrt(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="50000,250000")
    ap.add_argument("--cpu-queries", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    eng = geotrax_b200.Engine(frame_hw=(256, 384), imgsz=192, nc=1, max_batch=16, max_det=16, max_features=500, ransac_max_iter=10000)
    rng = np.random.default_rng(0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {}
    for n in [int(s) for s in a.sizes.split(",")]:
        train = rootsift_like(rng, n)
        query = rootsift_like(rng, n)
        m = n // 2
        src = rng.choice(n, m, replace=False)
        noisy = np.maximum(train[src] + rng.normal(0, 0.01, (m, 128)).astype(np.float32), 0)
        query[:m] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
        q, t = torch.from_numpy(query).cuda(), torch.from_numpy(train).cuda()
        idx = torch.empty((n, 2), dtype=torch.int32, device="cuda")
        dist = torch.empty((n, 2), dtype=torch.float32, device="cuda")
        call = lambda: eng._ck(eng.lib.gt_match_l2(eng.h, q.data_ptr(), n, t.data_ptr(), n, 128, idx.data_ptr(), dist.data_ptr(), None))
        call(); call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter(); call(); ts.append((time.perf_counter() - t0) * 1e3)    # the call synchronises its stream
        ms = float(np.median(ts))
        hit = float((idx[:m, 0].cpu().numpy() == src).mean())
        # CPU: the same train set, a bounded sample of the queries, all host threads OpenCV uses
        sl = np.linspace(0, n - 1, a.cpu_queries).astype(np.int64)
        t0 = time.perf_counter()
        ref = cv2.BFMatcher(cv2.NORM_L2).knnMatch(query[sl], train, k=2)
        cpu_s = time.perf_counter() - t0
        ri = np.array([[p[0].trainIdx, p[1].trainIdx] for p in ref]); rd = np.array([[p[0].distance, p[1].distance] for p in ref])
        gi, gd = idx.cpu().numpy()[sl], dist.cpu().numpy()[sl]
        flops = 2.0 * n * n * 128
        rec = dict(what="gt_match_l2 (f-3 registration matcher), device-resident descriptors", n_query=n, n_train=n, ms=ms, tflops=flops / (ms * 1e-3) * 1e-12,
                   frac_of_bf16_peak=flops / (ms * 1e-3) * 1e-12 / float(peaks.get("bf16_tflops", 1671.7)), true_match_recall=hit,
                   cpu=dict(kind="cv2.BFMatcher(NORM_L2).knnMatch", queries=int(a.cpu_queries), seconds=cpu_s, extrapolated_seconds_full=cpu_s * n / a.cpu_queries,
                            threads=cv2.getNumThreads()),
                   parity_on_sample=dict(index_pairs_equal=float((gi == ri).all(1).mean()), max_abs_dist_diff=float(np.abs(gd - rd).max())))
        print(json.dumps(rec), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
