#!/bin/bash
# ncu of every exported kernel at the BASELINE.json helper configurations (one launch per kernel and
# shape): first the plain run, then the launch list, then one --set full capture.
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
CMD="python tools/time_kernels.py --reps 1 --chunks=0 --sweep --dataflow"
TK_OUT=$out/tk_plain_$tag.json timeout 600 $CMD > $out/plain_$tag.log 2>&1 &&
TK_OUT=$out/tk_ncu_$tag.json timeout 1500 ncu --set full --clock-control none --import-source on \
    -k "regex:bg13|bgw_|bg_generic|transpose_|madnz_|threshold_|expand_flags|percentile5|maskedsum|dataflow" \
    -f -o $out/prof_$tag $CMD > $out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"; tail -3 $out/ncu_full_$tag.log
ls -la $out | tail -5
