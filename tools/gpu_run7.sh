#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py tests/test_postproc_gpu.py -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2d_tests.log
timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r2d_layers_orig16.log 2>&1
timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r2d_layers_fast32.log 2>&1
tail -7 gpurun_out/r2d_layers_orig16.log; tail -7 gpurun_out/r2d_layers_fast32.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_orig256.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f frac %.3f cnn %.1f pp %.2f conv0 %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_classes']['cnn_total_ms'], d['kernel_classes']['postproc']['ms'], d['kernel_classes']['conv0']['ms']), d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
/usr/local/cuda/bin/compute-sanitizer --tool initcheck --print-limit 5 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_initcheck.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r2_sanitizer_initcheck.log; grep "Uninitialized\|Host Frame: hvn\|at .*k_" gpurun_out/r2_sanitizer_initcheck.log | sort | uniq -c | sort -rn | head -8
