#!/bin/bash
# r3i: variant agreement test + compute-sanitizer over the new kernels (k_conv_ar, k_conv_rs<3/5>, XF producer changes)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 300 python -m pytest tests/test_cnn_gpu.py -x -q -k "variants" > gpurun_out/r3i_variants_test.log 2>&1; echo "variants test rc=$?"; tail -3 gpurun_out/r3i_variants_test.log
HVN_SAN_MODE=original timeout 240 $S --tool racecheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r3_sanitizer_racecheck_original.log 2>&1
echo "racecheck original rc=$? : $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r3_sanitizer_racecheck_original.log | tail -1)"
HVN_SAN_CNN_ONLY=1 timeout 200 $S --tool racecheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r3_sanitizer_racecheck_fast.log 2>&1
echo "racecheck fast rc=$? : $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r3_sanitizer_racecheck_fast.log | tail -1)"
HVN_SAN_MODE=original timeout 200 $S --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r3_sanitizer_memcheck_original.log 2>&1
echo "memcheck original rc=$? : $(grep -E 'ERROR SUMMARY' gpurun_out/r3_sanitizer_memcheck_original.log | tail -1)"
HVN_SAN_CNN_ONLY=1 timeout 200 $S --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r3_sanitizer_memcheck_fast.log 2>&1
echo "memcheck fast rc=$? : $(grep -E 'ERROR SUMMARY' gpurun_out/r3_sanitizer_memcheck_fast.log | tail -1)"
HVN_SAN_MODE=original timeout 200 $S --tool synccheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r3_sanitizer_synccheck_original.log 2>&1
echo "synccheck original rc=$? : $(grep -E 'ERROR SUMMARY' gpurun_out/r3_sanitizer_synccheck_original.log | tail -1)"
