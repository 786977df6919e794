"""From a summary written by tools/ncu_summary.py (tools/gpu_ncu2.sh) make the two committed
digests of an ncu capture: the per-kernel table (time, DRAM traffic, instructions, pipes) and
profiles/roofline_traffic.json, the DRAM bytes per launch that bench.py reports as
roofline.traffic when its sources match.

    python tools/profile_tables.py gpurun_out/ncu_full_summary_<tag>.txt <tag>
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def sections(path):
    cur = None
    for line in open(path):
        m = re.match(r"## (.*?)\s+grid=(\(.*?\)) block=", line)
        if m:
            cur = {"kernel": m.group(1), "grid": m.group(2)}
            yield cur
        elif cur is not None and line.strip() and not line.startswith("#"):
            parts = line.split()
            try:
                value = float(parts[1])
            except (IndexError, ValueError):
                continue
            cur[parts[0]] = value * UNITS.get(parts[2] if len(parts) > 2 else "", 1.0)


def main():
    path, tag = sys.argv[1], sys.argv[2]
    secs = list(sections(path))
    sha = bench.source_sha16()
    lines = [f"# ncu --set full --clock-control none, one launch per kernel and grid (tools/gpu_ncu2.sh {tag});",
             "# shapes of tools/time_kernels.py: 32768 channels x 8320 baselines unless the grid says otherwise.",
             "# us: gpu__time_duration; dram: bytes read + written per launch; winst: warp instructions "
             "(x 32 / 272.6e6 = per visibility);",
             "# issue / alu / fma / lsu / xu: % of peak sustained active (issue slots, pipes).",
             f"# flagger sources {sha}",
             f"{'kernel':48s} {'grid':>16s} {'us':>8s} {'dram_MB':>9s} {'GB/s':>7s} {'winst_M':>8s} "
             f"{'issue':>6s} {'alu':>5s} {'fma':>5s} {'lsu':>5s} {'xu':>5s}"]
    for s in secs:
        us = s.get("gpu__time_duration.sum", 0.0)
        dram = s.get("dram__bytes_read.sum", 0.0) + s.get("dram__bytes_write.sum", 0.0)
        s["dram"] = dram
        pipe = lambda n: s.get(f"sm__inst_executed_pipe_{n}.avg.pct_of_peak_sustained_active", 0.0)  # noqa: E731
        lines.append(f"{s['kernel']:48s} {s['grid']:>16s} {us:8.1f} {dram / 1e6:9.1f} {dram / max(us, 1e-9) / 1e3:7.0f} "
                     f"{s.get('smsp__inst_executed.sum', 0) / 1e6:8.1f} "
                     f"{s.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):6.1f} "
                     f"{pipe('alu'):5.1f} {pipe('fma'):5.1f} {pipe('lsu'):5.1f} {pipe('xu'):5.1f}")
    table = os.path.join(ROOT, "profiles", f"{tag}_kernel_table.txt")
    with open(table, "w") as f:
        f.write("\n".join(lines) + "\n")

    def pick(prefix, grid):
        for s in secs:
            if s["kernel"].startswith(prefix) and s["grid"] == grid:
                return int(round(s["dram"], -3))
        raise SystemExit(f"no launch of {prefix} with grid {grid} in {path}")

    # the launches of the fused flagger (32768 x 8320): the background filter over the whole dump,
    # the others those of one 2368-baseline chunk
    record = {"tag": f"{tag} (profiles/{tag}_ncu_full.txt)", "source_sha16": sha,
              "kernels": {"dataflow_kernel": pick("dataflow_kernel", "(592, 1, 1)"),
                          "bg13_kernel": pick("bg13_kernel<0, 0, 1", "(260, 128, 1)"),
                          "madnz_stream_kernel": pick("madnz_stream_kernel", "(2368, 1, 1)"),
                          "threshold_sum_kernel": pick("threshold_sum_kernel<1, 128, 1>", "(1332, 1, 1)"),
                          "expand_flags_kernel": pick("expand_flags_kernel", "(10, 128, 1)")},
              "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch; bg13_kernel covers the whole "
                      "8320-baseline dump, the other launches are those of a 2368-baseline chunk"}
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as f:
        json.dump(record, f, indent=1)
    print(table)
    print(json.dumps(record["kernels"]))


if __name__ == "__main__":
    main()
