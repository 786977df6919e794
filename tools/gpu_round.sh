#!/bin/bash
# One GPU visit: parity tests, smoke, both bench arms, then the ncu launch list and one
# full capture of the flagger's kernels (profiling recipe of B200_PROFILING.md).
# Usage (on the GPU box): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_$tag.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -5 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1
echo "smoke rc=$?"; tail -3 $out/smoke_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
echo "bench rc=$?"; cat $out/bench_$tag.json; tail -5 $out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "bench ref rc=$?"; cat $out/bench_ref_$tag.json
if [ "$2" == "ncuonly" ] || [ "$2" != "noncu" ]; then
KREGEX='regex:bg13_kernel|madnz_stream_kernel|threshold_sum_kernel|expand_flags_kernel'
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv \
    --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$KREGEX" -s 24 -c 8 \
    -f -o $out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"; tail -3 $out/ncu_full_$tag.log
fi
ls -la $out
