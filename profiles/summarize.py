"""Turns the raw ncu outputs a gpurun call brought back (gpurun_out/) into the small
tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv profiles/r01_launches_upwind512.md
    python profiles/summarize.py full gpurun_out/prof_upwind_r01.ncu-rep profiles/r01_ncu_upwind3d_tma_512.md \
        --traffic-key upwind3d_tma_kernel:512x512x512
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1)
        a = agg.setdefault(r["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write(f"# ncu launch list summary ({os.path.basename(src)})\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over one bench.py command; per-launch times are\n"
                 "cold-cache and serialised, so read the SHARES.\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k[:110]}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.1%} |\n")
    print(open(dst).read())


def full(src, dst, traffic_key=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full summary ({os.path.basename(src)})\n\n")
        fh.write(f"kernel: `{data[0][hdr.index('Kernel Name')][:160]}`  \nlaunches captured: {len(data)}\n\n")
        fh.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n")
        fh.write("|---|---|" + "---:|" * len(data) + "\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                fh.write(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        tr = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in data]
        fh.write(f"\nDRAM traffic per launch (read+write): {sum(tr) / len(tr) / 1e9:.4f} GB (mean of {len(tr)})\n")
    if traffic_key:
        tj_path = os.path.join(HERE, "traffic.json")
        tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
        tj[traffic_key] = sum(tr) / len(tr)
        json.dump(tj, open(tj_path, "w"), indent=1, sort_keys=True)
    print(open(dst).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst)
    else:
        key = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
        full(src, dst, key)
