#!/bin/bash
# r3j: lean epilogue write-out (per-tile output offsets): self test, A/B, CNN tests
mkdir -p gpurun_out
export AB_REPS=5
timeout 300 python tools/gpu_diag.py tc original > gpurun_out/r3j_tc_orig.log 2>&1; echo "tc orig rc=$?"; grep "e2e\|ERROR\|rror" gpurun_out/r3j_tc_orig.log | head -4
timeout 300 python tools/gpu_diag.py tc fast > gpurun_out/r3j_tc_fast.log 2>&1; echo "tc fast rc=$?"; grep "e2e\|ERROR\|rror" gpurun_out/r3j_tc_fast.log | head -4
timeout 600 python tools/gpu_diag.py ab original 16 "new:" "lean0:tc_lean_epi=0" > gpurun_out/r3j_ab_orig16.log 2>&1; echo "rc=$?"
grep -v "^decoder.np\|^decoder.hv" gpurun_out/r3j_ab_orig16.log | head -70
timeout 600 python tools/gpu_diag.py ab fast 32 "new:" "lean0:tc_lean_epi=0" > gpurun_out/r3j_ab_fast32.log 2>&1; echo "rc=$?"
grep "^layer\|TOTAL\|decoder.tp.u2" gpurun_out/r3j_ab_fast32.log
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q > gpurun_out/r3j_cnn_tests.log 2>&1; echo "cnn tests rc=$?"; tail -3 gpurun_out/r3j_cnn_tests.log
