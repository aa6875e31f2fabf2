#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python tools/ncu_summary.py rep   gpurun_out/r01a/prof_fill.ncu-rep   profiles/r01_fill_raster_kernel.csv
    python tools/ncu_summary.py list  gpurun_out/r01a/launches_fill.csv   profiles/r01_fill_launches.csv
"""
import csv, io, subprocess, sys, collections

METRICS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_membar.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
]


def rep(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [m for m in METRICS if m in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for m in cols:
            i = hdr.index(m)
            w.writerow([m, units[i]] + [r[i] for r in rows[2:]])
    print(f"wrote {dst}: {len(rows) - 2} launches x {len(cols)} metrics")


def launch_list(src, dst):
    text = open(src).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = (r["Kernel Name"], r["Block Size"], r["Grid Size"])
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        v = float(r["Metric Value"].replace(",", "")) / (1.0 if r["Metric Unit"] in ("ns", "nsecond") else 1e-3)
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    total = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "block", "grid", "launches", "total_ns", "avg_ns", "min_ns", "max_ns", "share_of_listed_time"])
        for (k, b, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, b, g, a[0], round(a[1]), round(a[1] / a[0]), round(a[2]), round(a[3]), round(a[1] / total, 4)])
    print(f"wrote {dst}: {len(agg)} distinct kernels, {len(rows)} launches")


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
