#!/bin/bash
# Source-level ncu capture (per-line instruction counts and stall samples) of a few kernels.
# Usage: bash tools/gpu_ncu_src.sh <tag> <kernel regex> [launch count]
tag=${1:-src}; rx=${2:-madnz_stream_kernel}; n=${3:-1}
out=gpurun_out; mkdir -p $out
CMD="python tools/time_kernels.py --reps 0 --chunks=0"
TK_OUT=$out/tk_plain_$tag.json timeout 600 $CMD > $out/plain_$tag.log 2>&1 &&
TK_OUT=$out/tk_ncu_$tag.json timeout 900 ncu --set full --import-source on --clock-control none \
    -k "regex:$rx" -c $n -f -o $out/prof_$tag $CMD > $out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -2 $out/ncu_$tag.log; ls -la $out/prof_$tag.ncu-rep
