#!/bin/bash
# ncu only: launch list + one full capture of the flagger's kernels through bench.py
tag=${1:-r01}
out=gpurun_out; mkdir -p $out
KREGEX='regex:bg13_kernel|madnz_stream_kernel|threshold_sum_kernel|expand_flags_kernel'
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv \
    --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$KREGEX" -s 24 -c 8 \
    -f -o $out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"; tail -2 $out/ncu_full_$tag.log
