#!/bin/bash
# round-2 GPU session 3: XF in-place transform (3/4 stages), per-channel weight exponents, range guard: tests + layers + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -25 gpurun_out/r2b_tests.log
if [ $rc -ne 0 ]; then HVN_OPTS=stem_tc=0 timeout 600 python -m pytest tests/test_cnn_gpu.py -m gpu -q > gpurun_out/r2b_tests_nostem.log 2>&1; echo "no-stem rc=$?"; tail -8 gpurun_out/r2b_tests_nostem.log; fi
timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r2b_layers_orig16.log 2>&1
timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r2b_layers_fast32.log 2>&1
tail -7 gpurun_out/r2b_layers_orig16.log; tail -7 gpurun_out/r2b_layers_fast32.log
timeout 300 python tools/gpu_diag.py tc original > gpurun_out/r2b_tc_orig.log 2>&1; tail -12 gpurun_out/r2b_tc_orig.log
timeout 300 python tools/gpu_diag.py ppprof > gpurun_out/r2b_ppprof.log 2>&1; tail -30 gpurun_out/r2b_ppprof.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_orig256.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f frac %.3f cnn %.1f pp %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_classes']['cnn_total_ms'], d['kernel_classes']['postproc']['ms']))
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
