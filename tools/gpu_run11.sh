#!/bin/bash
# r3b: L2 prefetch cursor, AR with two residual regions: self test + A/B layer timing
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py tc original > gpurun_out/r3b_tc_orig.log 2>&1; echo "tc orig rc=$?"; grep "e2e\|worst\|rel \|ERROR\|rror" gpurun_out/r3b_tc_orig.log | head -8
timeout 300 python tools/gpu_diag.py tc fast > gpurun_out/r3b_tc_fast.log 2>&1; echo "tc fast rc=$?"; grep "e2e\|ERROR\|rror" gpurun_out/r3b_tc_fast.log | head -4
for cfg in "new:" "pf0:tc_prefetch=0" "pf8:tc_prefetch=8" "ar4:tc_ar_min_chunks=4" "ar0:tc_ar=0" "nres1:tc_ar_nres=1"; do
  tag=${cfg%%:*}; opts=${cfg#*:}
  HVN_OPTS="$opts" timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r3b_layers_orig16_$tag.log 2>&1
  echo "$tag: $(grep 'conv_tc ' gpurun_out/r3b_layers_orig16_$tag.log | tail -1) | $(grep 'cnn total' gpurun_out/r3b_layers_orig16_$tag.log)"
done
python tools/layer_diff.py gpurun_out/r3b_layers_orig16_{new,pf0,pf8,ar4,ar0,nres1}.log
for cfg in "new:" "pf0:tc_prefetch=0" "ar0:tc_ar=0"; do
  tag=${cfg%%:*}; opts=${cfg#*:}
  HVN_OPTS="$opts" timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r3b_layers_fast32_$tag.log 2>&1
  echo "$tag fast: $(grep 'conv_tc ' gpurun_out/r3b_layers_fast32_$tag.log | tail -1) | $(grep 'cnn total' gpurun_out/r3b_layers_fast32_$tag.log)"
done
python tools/layer_diff.py gpurun_out/r3b_layers_fast32_{new,pf0,ar0}.log
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q > gpurun_out/r3b_cnn_tests.log 2>&1; echo "cnn tests rc=$?"; tail -3 gpurun_out/r3b_cnn_tests.log
