"""Per-kernel share of one 16-frame step from an ncu launch list (gpu__time_duration.sum).

    python tools/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/summary.md

A step starts at a 16-frame preprocess_half_kernel launch (grid 8100); step_index counts those (default: the 5th = the first
timed step of `bench.py --steps 2 --warmup 3`).  ncu serialises launches and runs them cold, so the absolute sum is larger than
the CUDA-event step time bench.py reports; the SHARES are what is compared."""
import collections, csv, sys

def main():
    path = sys.argv[1]
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    starts = [i for i, r in enumerate(rows) if "preprocess_half" in r[4] and r[8].startswith("(8100")]
    a = starts[want]
    b = starts[want + 1] if want + 1 < len(starts) else len(rows)
    agg = collections.OrderedDict()
    for r in rows[a:b]:
        name = r[4].split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()
        e = agg.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += float(r[-1].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches / step | us / step (ncu, serialised) | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100 % |")

if __name__ == "__main__":
    main()
