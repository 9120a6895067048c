"""Per-level and per-layer gap between the measured conv launches and their roofline bound, from a GT_TUNE_LOG capture.

    python tools/tune_gap.py profiles/tune_log_r1_pair.txt [MEASURED_PEAKS.json]

The log line of a layer carries the winner's time (us per 16-frame launch), TFLOP/s and GB/s, i.e. its FLOPs and algorithmic bytes;
bound = max(FLOPs / bf16 peak, bytes / HBM peak) with the measured sustained peaks (defaults: 1413.3 TFLOP/s, 6463 GB/s)."""
import collections
import json
import re
import sys


def main():
    path = sys.argv[1]
    tf_peak, gb_peak = 1413.3, 6463.0
    if len(sys.argv) > 2:
        pk = json.load(open(sys.argv[2]))
        tf_peak = float(pk.get("bf16_tflops_sustained", tf_peak))
        gb_peak = float(pk.get("hbm_gbs", gb_peak))
    pat = re.compile(r"op\s+(\d+) src\s+(\d+) cin\s+(\d+) cout\s+(\d+) k (\d) s (\d) out\s+(\d+)x(\d+)\s+(.*)->\s+(\S+)\s+([\d.]+) TFLOP/s\s+([\d.]+) GB/s")
    rows = []
    for line in open(path):
        m = pat.search(line)
        if not m:
            continue
        op, _, cin, cout, k, s, H, W, times, best, tf, gb = m.groups()
        t = dict(zip(times.split()[0::2], map(float, times.split()[1::2])))
        us, tf, gb = t[best], float(tf), float(gb)
        bound = max(tf * us / tf_peak, gb * us / gb_peak)
        rows.append(dict(op=int(op), H=int(H), W=int(W), us=us, bound=bound, best=best, cin=int(cin), cout=int(cout), k=int(k), s=int(s), tf=tf, gb=gb))
    lev = collections.OrderedDict()
    for r in rows:
        e = lev.setdefault((r["H"], r["W"]), [0, 0.0, 0.0])
        e[0] += 1
        e[1] += r["us"]
        e[2] += r["bound"]
    print("| level (output H x W) | launches | us / step | roofline bound | gap |\n|---|---|---|---|---|")
    for (H, W), (n, us, bd) in lev.items():
        print(f"| {H}x{W} | {n} | {us:.0f} | {bd:.0f} | {us - bd:.0f} |")
    print(f"| **all** | {len(rows)} | {sum(r['us'] for r in rows):.0f} | {sum(r['bound'] for r in rows):.0f} | {sum(r['us'] - r['bound'] for r in rows):.0f} |")
    print("\n| op | layer | variant | us | bound | gap | TFLOP/s | GB/s |\n|---|---|---|---|---|---|---|---|")
    for r in sorted(rows, key=lambda r: r["bound"] - r["us"])[:12]:
        print(f"| {r['op']} | {r['cin']}->{r['cout']} {r['k']}x{r['k']} s{r['s']} @{r['H']}x{r['W']} | {r['best']} | {r['us']:.1f} | {r['bound']:.1f} | {r['us'] - r['bound']:.1f} | {r['tf']:.0f} | {r['gb']:.0f} |")


if __name__ == "__main__":
    main()
