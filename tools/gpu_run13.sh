#!/bin/bash
# r3d: XF early raw-slot release + raw cursor: self test, interleaved A/B, CNN tests
mkdir -p gpurun_out
export AB_REPS=7
timeout 300 python tools/gpu_diag.py tc original > gpurun_out/r3d_tc_orig.log 2>&1; echo "tc orig rc=$?"; grep "e2e\|ERROR\|rror" gpurun_out/r3d_tc_orig.log | head -4
timeout 300 python tools/gpu_diag.py tc fast > gpurun_out/r3d_tc_fast.log 2>&1; echo "tc fast rc=$?"; grep "e2e\|ERROR\|rror" gpurun_out/r3d_tc_fast.log | head -4
timeout 600 python tools/gpu_diag.py ab original 16 "new:" "early0:tc_xf_early=0" "pf0:tc_prefetch=0" "pf2:tc_prefetch=2" "pf6:tc_prefetch=6" > gpurun_out/r3d_ab_orig16.log 2>&1; echo "rc=$?"
grep -v "^decoder.np\|^decoder.hv" gpurun_out/r3d_ab_orig16.log
timeout 600 python tools/gpu_diag.py ab fast 32 "new:" "early0:tc_xf_early=0" "pf0:tc_prefetch=0" "pf2:tc_prefetch=2" "pf6:tc_prefetch=6" > gpurun_out/r3d_ab_fast32.log 2>&1; echo "rc=$?"
grep -v "^decoder.np\|^decoder.hv" gpurun_out/r3d_ab_fast32.log
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q > gpurun_out/r3d_cnn_tests.log 2>&1; echo "cnn tests rc=$?"; tail -3 gpurun_out/r3d_cnn_tests.log
