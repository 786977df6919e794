#!/bin/bash
# ncu capture of the 2-D flagger's per-baseline kernel: raw metrics (pipes, stalls) and the
# source lines with the most samples.  Usage: bash tools/gpu_ncu_td.sh <tag> [n_time n_freq n_bl]
tag=${1:-td}; shift
shape=${@:-16 4096 1024}
out=gpurun_out; mkdir -p $out
CMD="python tools/time_twodflag.py $shape --reps 1"
TK_OUT=$out/td_plain_$tag.json timeout 600 $CMD > $out/td_plain_$tag.log 2>&1 &&
TK_OUT=$out/td_ncu_$tag.json timeout 1200 ncu --set full --import-source on --clock-control none \
    -k "regex:twod_baseline_kernel" -s 1 -c 1 -f -o /tmp/prof_td_$tag $CMD > $out/td_ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -2 $out/td_ncu_$tag.log
ncu -i /tmp/prof_td_$tag.ncu-rep --page raw --csv > /tmp/td_raw.csv 2>/dev/null
python - "$out/td_raw_$tag.txt" <<'PY'
import csv, sys
rows = list(csv.reader(open('/tmp/td_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("pipe", "issue", "stall", "duration", "dram__bytes", "inst_executed.sum", "registers", "occupancy",
        "warps_active", "eligible", "l1tex__t_sector_hit", "lts__t_sector_hit", "shared")
with open(sys.argv[1], "w") as f:
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in keep) and "pct_of_peak_sustained_elapsed" not in h:
            f.write(f"{h:100s} {v} {u}\n")
PY
ncu -i /tmp/prof_td_$tag.ncu-rep --page source --csv --print-source sass,cuda > /tmp/td_src.csv 2>/dev/null || \
ncu -i /tmp/prof_td_$tag.ncu-rep --page source --csv > /tmp/td_src.csv 2>/dev/null
python - "$out/td_src_$tag.txt" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open('/tmp/td_src.csv')))
hi = next(i for i, r in enumerate(rows) if any("Sampling" in c for c in r))
hdr = rows[hi]
idx = {}
for i, c in enumerate(hdr):
    idx.setdefault(c.strip(), i)
c_line, c_cuda, c_samp, c_inst = idx["Line No"], 1, idx["# Samples"], idx["Instructions Executed"]
c_sass = 3
c_bar, c_long, c_short, c_wait = idx["stall_barrier"], idx["stall_long_sb"], idx["stall_short_sb"], idx["stall_wait"]
per_line = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
ops = collections.Counter(); opi = collections.Counter()
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    try:
        s_, i_ = int(r[c_samp] or 0), int(r[c_inst] or 0)
    except (ValueError, IndexError):
        continue
    cuda_row = bool(r[c_cuda].strip()) or not r[c_sass].strip()
    if cuda_row:
        tot_s += s_; tot_i += i_
        key = (r[c_line], r[c_cuda].strip()[:110])
        e = per_line[key]
        e[0] += s_; e[1] += i_
        for j, c in enumerate((c_bar, c_long, c_short)):
            try: e[2 + j] += int(r[c] or 0)
            except ValueError: pass
        continue
    sass = r[c_sass].split()
    op = (sass[1] if sass and sass[0].startswith('@') and len(sass) > 1 else (sass[0] if sass else "?"))
    op = ".".join(op.split(".")[:2])
    ops[op] += s_; opi[op] += i_
tot_s = tot_s or 1; tot_i = tot_i or 1
with open(sys.argv[1], "w") as f:
    f.write(f"total samples {tot_s}, warp instructions {tot_i}\n")
    f.write("\nby SASS opcode (instructions %, samples %):\n")
    for op, i_ in opi.most_common(40):
        f.write(f"  {op:22s} {100*i_/tot_i:6.2f} {100*ops[op]/tot_s:6.2f}\n")
    f.write("\nCUDA lines by instructions (instr %, samples %, of which barrier / long scoreboard / short scoreboard %):\n")
    for (ln, src), e in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:90]:
        f.write(f"  {100*e[1]/tot_i:6.2f} {100*e[0]/tot_s:6.2f} ({100*e[2]/tot_s:5.2f} {100*e[3]/tot_s:5.2f} {100*e[4]/tot_s:5.2f})  {ln:>5s} {src}\n")
    f.write("\nCUDA lines by samples:\n")
    for (ln, src), e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:50]:
        f.write(f"  {100*e[1]/tot_i:6.2f} {100*e[0]/tot_s:6.2f} ({100*e[2]/tot_s:5.2f} {100*e[3]/tot_s:5.2f} {100*e[4]/tot_s:5.2f})  {ln:>5s} {src}\n")
PY
ls -la /tmp/prof_td_$tag.ncu-rep
