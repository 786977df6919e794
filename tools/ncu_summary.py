"""Summarise ncu outputs into small text files for profiles/ (developer tool).

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv       -> per-kernel time shares
  python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep            -> key metrics per launch
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "launch__grid_size", "launch__block_size",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("<unnamed>::", "")
    return name.split("(")[0]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        key = (short(r[ki]), r[gi], r[bi])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total:.1f} us in total "
          "(ncu: cold-cache, serialised; compare shares)")
    print(f"{'kernel':50s} {'grid':>14s} {'block':>12s} {'n':>5s} {'total_us':>10s} {'mean_us':>9s} {'share':>6s}")
    for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:50s} {g:>14s} {b:>12s} {n:5d} {t:10.1f} {t / n:9.1f} {t / total:6.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    # instructions per pipe (ALU / FMA / LSU / XU ...) and pipe utilisation, whatever this ncu calls them
    cols += [(k, i) for i, k in enumerate(hdr)
             if re.search(r"inst_executed_pipe_[a-z_]+\.(sum|avg\.pct_of_peak_sustained_active)$", k)
             and not re.search(r"_pred_|_op_|tensor|_fp64|fp16|tma|tmem|uniform_|ipa|tex", k)]
    seen = collections.OrderedDict()
    for r in data:
        seen.setdefault((short(r[ki]), r[hdr.index("Grid Size")]), r)   # first launch of each shape
    print(f"# {path}: ncu --set full, one launch per kernel/grid shape")
    for (k, g), r in seen.items():
        print(f"\n## {k}  grid={g} block={r[hdr.index('Block Size')]}")
        for name, idx in cols:
            print(f"{name:85s} {r[idx]:>16s} {units[idx]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
