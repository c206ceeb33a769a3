#!/bin/bash
# r3c: in-process interleaved A/B of the launch-time knobs (median of 7)
mkdir -p gpurun_out
export AB_REPS=7
timeout 600 python tools/gpu_diag.py ab original 16 "new:" "pf0:tc_prefetch=0" "pf2:tc_prefetch=2" "ar0:tc_ar=0" "ar0pf0:tc_ar=0,tc_prefetch=0" "ar2:tc_ar_min_chunks=2" "old:tc_ar=0,tc_prefetch=0,tc_xf_trunc=0" > gpurun_out/r3c_ab_orig16.log 2>&1; echo "rc=$?"
cat gpurun_out/r3c_ab_orig16.log | grep -v "^decoder.np\|^decoder.hv"
timeout 600 python tools/gpu_diag.py ab fast 32 "new:" "pf0:tc_prefetch=0" "pf2:tc_prefetch=2" "ar0:tc_ar=0" "ar0pf0:tc_ar=0,tc_prefetch=0" "ar2:tc_ar_min_chunks=2" "old:tc_ar=0,tc_prefetch=0,tc_xf_trunc=0" > gpurun_out/r3c_ab_fast32.log 2>&1; echo "rc=$?"
cat gpurun_out/r3c_ab_fast32.log | grep -v "^decoder.np\|^decoder.hv"
