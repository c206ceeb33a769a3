#!/bin/bash
# r3g: row-stacked grouped conv (k_conv_rs): self test + A/B
mkdir -p gpurun_out
export AB_REPS=5
timeout 300 python tools/gpu_diag.py tc original v > gpurun_out/r3g_tc_orig.log 2>&1; echo "tc orig rc=$?"; grep "e2e\|ERROR\|rror\|rel " gpurun_out/r3g_tc_orig.log | head -8
grep "dense.units.[07].conv2" gpurun_out/r3g_tc_orig.log | head -6 | cut -c1-200
timeout 300 python tools/gpu_diag.py tc fast v > gpurun_out/r3g_tc_fast.log 2>&1; echo "tc fast rc=$?"; grep "e2e\|ERROR\|rror\|rel " gpurun_out/r3g_tc_fast.log | head -8
grep "dense.units.[03].conv2" gpurun_out/r3g_tc_fast.log | head -6 | cut -c1-200
timeout 600 python tools/gpu_diag.py ab original 16 "new:" "rs0:tc_rowstack=0" > gpurun_out/r3g_ab_orig16.log 2>&1; echo "rc=$?"
grep "^layer\|decoder.tp.*conv2\|TOTAL" gpurun_out/r3g_ab_orig16.log
timeout 600 python tools/gpu_diag.py ab fast 32 "new:" "rs0:tc_rowstack=0" > gpurun_out/r3g_ab_fast32.log 2>&1; echo "rc=$?"
grep "^layer\|decoder.tp.*conv2\|TOTAL" gpurun_out/r3g_ab_fast32.log
