#!/bin/bash
# One GPU visit of round 2: parity tests, smoke, per-kernel timings, the bench line.
# Usage (on the GPU box): bash tools/gpu_r02.sh <tag> [tests|nobench|...]
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $out/smi_$tag.txt
if [ "$2" != "notests" ]; then
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -6 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1
echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log
fi
TK_OUT=$out/time_kernels_$tag.json timeout 600 python tools/time_kernels.py --reps 7 --chunks=0 2>&1 | grep -v "^$" > $out/time_kernels_$tag.log
echo "time_kernels rc=$?"; cat $out/time_kernels_$tag.log
if [ "$2" != "nobench" ]; then
timeout 1200 python bench.py $BENCH_ARGS > $out/bench_$tag.json 2> $out/bench_$tag.err
echo "bench rc=$?"; cat $out/bench_$tag.json; tail -5 $out/bench_$tag.err
fi
