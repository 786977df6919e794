#!/bin/bash
# The warp-specialised background kernel: parity, then timings with and without it.
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -q --timeout 300 -x > $out/pytest_ws.log 2>&1
echo "pytest rc=$?"; tail -4 $out/pytest_ws.log
for ws in 1 0; do
  echo "== KSP_BG_WS=$ws"
  KSP_BG_WS=$ws TK_OUT=$out/tk_ws$ws.json timeout 300 python tools/time_kernels.py --reps 7 --chunks=0 2>&1 | grep "^background\|^flagger_fused" | grep -v scratch
done
