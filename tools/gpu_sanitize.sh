#!/bin/bash
# compute-sanitizer over the end-to-end target: memcheck (both modes), racecheck / synccheck / initcheck (fast mode).
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 $S --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1)"
done
HVN_SAN_MODE=original timeout 900 $S --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_memcheck_original.log 2>&1
echo "memcheck original rc=$? : $(grep -E 'ERROR SUMMARY' gpurun_out/r2_sanitizer_memcheck_original.log | tail -1)"
HVN_SAN_MODE=original timeout 900 $S --tool racecheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_racecheck_original.log 2>&1
echo "racecheck original rc=$? : $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r2_sanitizer_racecheck_original.log | tail -1)"
