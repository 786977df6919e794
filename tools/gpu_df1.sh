#!/bin/bash
# First GPU visit of the dataflow flagger: its parity tests, then timings against the chunked form
# and a sweep of the schedule's lags / ring depth (each configuration is a process: the library
# reads them once).
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_df1.txt
timeout 900 python -m pytest tests/test_gpu_dataflow.py -q --timeout 300 -x > $out/pytest_df1.log 2>&1
echo "pytest dataflow rc=$?"; tail -15 $out/pytest_df1.log
TK_OUT=$out/tk_df_default.json timeout 300 python tools/time_kernels.py --only-fused --chunks=-1,2368 --reps 7 2>&1 | grep -v "^$" | tail -8
for cfg in "1 2 2 0" "2 2 2 0" "3 3 3 0" "4 4 4 0" "3 3 3 24" "6 6 6 0"; do
  set -- $cfg
  echo "== L1=$1 L2=+$2 L3=+$3 ring=$4"
  env KSP_DF_L1=$1 KSP_DF_L2=$2 KSP_DF_L3=$3 $( [ "$4" != "0" ] && echo KSP_DF_RING=$4 ) TK_OUT=$out/tk_df_$1_$2_$3_$4.json \
     timeout 300 python tools/time_kernels.py --only-fused --chunks=-1 --reps 7 2>&1 | grep "flagger_dataflow" | tail -1
done
