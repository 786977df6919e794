#!/bin/bash
# ncu of every exported kernel at the BASELINE.json helper configurations (one launch per kernel and
# shape): the plain run first, then one --set full capture, summarised on the box (the report
# itself only comes back if it is small).
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
CMD="python tools/time_kernels.py --reps 0 --chunks=0 --sweep --dataflow"
TK_OUT=$out/tk_plain_$tag.json timeout 600 $CMD > $out/plain_$tag.log 2>&1 &&
TK_OUT=$out/tk_ncu_$tag.json timeout 1500 ncu --set full --clock-control none \
    -k "regex:bg13|bgw_|bg_generic|transpose_|madnz_|threshold_|expand_flags|percentile5|maskedsum|dataflow" \
    -f -o /tmp/prof_$tag $CMD > $out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"; tail -3 $out/ncu_full_$tag.log
python tools/ncu_summary.py full /tmp/prof_$tag.ncu-rep > $out/ncu_full_summary_$tag.txt 2> $out/ncu_summary_$tag.err
ls -la /tmp/prof_$tag.ncu-rep
if [ $(stat -c %s /tmp/prof_$tag.ncu-rep) -lt 40000000 ]; then cp /tmp/prof_$tag.ncu-rep $out/; fi
ls -la $out | tail -8
