#!/usr/bin/env python
"""Reduces an `ncu --set full` report to one CSV row per launch with the counters DESIGN.md / the judge read:

    ncu -i gpurun_out/step_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_reduce.py /tmp/raw.csv > profiles/step_full_r2.csv

Columns: kernel, grid, block, duration (us), DRAM read / write (MB), DRAM throughput (% of peak), tensor-pipe activity (% of cycles:
sm__pipe_tensor_subpipe_hmma_cycles_active / sm__inst_executed_pipe_tensor*), issue slots busy (%), achieved occupancy (%), registers,
dynamic + static shared memory (KB), L2 hit rate (%)."""
import csv
import sys

WANT = [
    ("duration_us", ["gpu__time_duration.sum"], 1e-3),
    ("dram_read_MB", ["dram__bytes_read.sum"], None),
    ("dram_write_MB", ["dram__bytes_write.sum"], None),
    ("dram_pct", ["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"], 1),
    ("tensor_pipe_pct", ["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                         "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"], 1),
    ("issue_slots_pct", ["sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active"], 1),
    ("occupancy_pct", ["sm__warps_active.avg.pct_of_peak_sustained_active"], 1),
    ("regs", ["launch__registers_per_thread"], 1),
    ("smem_dyn_KB", ["launch__shared_mem_per_block_dynamic"], None),
    ("smem_static_KB", ["launch__shared_mem_per_block_static"], None),
    ("l2_hit_pct", ["lts__t_sector_hit_rate.pct"], 1),
    ("sm_throughput_pct", ["sm__throughput.avg.pct_of_peak_sustained_elapsed"], 1),
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {n: i for i, n in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(["id", "kernel", "grid", "block"] + [w[0] for w in WANT])
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        line = [r[col["ID"]], name, r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]]
        for key, names, scale in WANT:
            val = ""
            for n in names:
                if n in col and r[col[n]] != "":
                    v = to_float(r[col[n]])
                    if v is None:
                        continue
                    u = units[col[n]].lower()
                    if scale is None:      # bytes -> MB / KB
                        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "": 1}.get(u.replace("/block", ""), 1)
                        v = v * mult / (1e6 if key.endswith("MB") else 1e3)
                    elif key == "duration_us":
                        mult = {"nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1e-3)
                        v = v * mult
                    val = f"{v:.2f}"
                    break
            line.append(val)
        out.writerow(line)


if __name__ == "__main__":
    main()
