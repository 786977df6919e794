#!/bin/bash
# Quick GPU visit while iterating on kernels: C-ABI parity tests + per-kernel timings.
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cabi.py -q --timeout 300 -x > $out/pytest_quick.log 2>&1
echo "pytest rc=$?"; tail -15 $out/pytest_quick.log
for t in ${TS_LIST:-256}; do
  echo "== KSP_TS_THREADS=$t"
  KSP_TS_THREADS=$t timeout 300 python tools/time_kernels.py --reps 5 ${TK_ARGS} 2>&1 | grep -v "^$" | tail -22
done
