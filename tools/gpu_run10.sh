#!/bin/bash
# r3a: A-resident conv3 (k_conv_ar) + truncating XF split: per-layer self test, A/B layer timing, CNN parity tests
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py tc original > gpurun_out/r3a_tc_orig.log 2>&1; echo "tc orig rc=$?"; grep -v "^\[" gpurun_out/r3a_tc_orig.log | tail -14
timeout 300 python tools/gpu_diag.py tc fast > gpurun_out/r3a_tc_fast.log 2>&1; echo "tc fast rc=$?"; grep "e2e\|worst\|rel " gpurun_out/r3a_tc_fast.log | head -8
for cfg in "new:" "base:tc_ar=0,tc_xf_trunc=0" "aronly:tc_xf_trunc=0" "xfonly:tc_ar=0"; do
  tag=${cfg%%:*}; opts=${cfg#*:}
  HVN_OPTS="$opts" timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r3a_layers_orig16_$tag.log 2>&1
  echo "$tag: $(grep 'conv_tc ' gpurun_out/r3a_layers_orig16_$tag.log | tail -1) | $(grep 'cnn total' gpurun_out/r3a_layers_orig16_$tag.log)"
done
for cfg in "new:" "base:tc_ar=0,tc_xf_trunc=0"; do
  tag=${cfg%%:*}; opts=${cfg#*:}
  HVN_OPTS="$opts" timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r3a_layers_fast32_$tag.log 2>&1
  echo "$tag fast: $(grep 'conv_tc ' gpurun_out/r3a_layers_fast32_$tag.log | tail -1) | $(grep 'cnn total' gpurun_out/r3a_layers_fast32_$tag.log)"
done
python tools/layer_classes.py gpurun_out/r3a_layers_orig16_new.log gpurun_out/r3a_layers_orig16_base.log 2>&1 | tail -40
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q > gpurun_out/r3a_cnn_tests.log 2>&1; echo "cnn tests rc=$?"; tail -3 gpurun_out/r3a_cnn_tests.log
