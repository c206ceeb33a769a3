#!/bin/bash
# r3f: AR residual region sets (bytes in flight) A/B
mkdir -p gpurun_out
export AB_REPS=7
timeout 300 python tools/gpu_diag.py tc original > gpurun_out/r3f_tc_orig.log 2>&1; echo "tc orig rc=$?"; grep "e2e\|ERROR\|rror" gpurun_out/r3f_tc_orig.log | head -4
HVN_OPTS="tc_ar_min_wst=2" timeout 300 python tools/gpu_diag.py tc fast > gpurun_out/r3f_tc_fast.log 2>&1; echo "tc fast (min_wst 2) rc=$?"; grep "e2e\|ERROR\|rror" gpurun_out/r3f_tc_fast.log | head -4
timeout 600 python tools/gpu_diag.py ab original 16 "new:" "w2:tc_ar_min_wst=2" "n2:tc_ar_nres=2" "n1:tc_ar_nres=1" "ar0:tc_ar=0" > gpurun_out/r3f_ab_orig16.log 2>&1; echo "rc=$?"
grep "^layer\|conv3.weight\|TOTAL" gpurun_out/r3f_ab_orig16.log
timeout 600 python tools/gpu_diag.py ab fast 32 "new:" "w2:tc_ar_min_wst=2" "n2:tc_ar_nres=2" "n1:tc_ar_nres=1" "ar0:tc_ar=0" > gpurun_out/r3f_ab_fast32.log 2>&1; echo "rc=$?"
grep "^layer\|conv3.weight\|TOTAL" gpurun_out/r3f_ab_fast32.log
