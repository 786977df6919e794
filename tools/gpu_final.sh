#!/bin/bash
# The evidence of one tag in one GPU visit: parity tests, smoke, per-kernel timings, the bench line
# (parity / cfg5 / twodflag blocks), the reference arm, ncu of every kernel, the launch list of the
# bench command and the 2-D flagger's kernel under ncu.   Usage: bash tools/gpu_final.sh <tag>
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
BENCH_ARGS="--twodflag" bash tools/gpu_r02.sh $tag
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "bench ref rc=$?"; cat $out/bench_ref_$tag.json
bash tools/gpu_ncu2.sh $tag
KREGEX='regex:bg13_kernel|madnz_stream_kernel|threshold_sum_kernel|expand_flags_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv \
    --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-cfg5 \
    > $out/ncu_launches_$tag.log 2>&1
echo "ncu launches rc=$?"
bash tools/gpu_ncu_td.sh ${tag#r02}
