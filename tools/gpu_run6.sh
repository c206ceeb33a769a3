#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r2c_layers_orig16.log 2>&1
timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r2c_layers_fast32.log 2>&1
tail -7 gpurun_out/r2c_layers_orig16.log; tail -7 gpurun_out/r2c_layers_fast32.log
HVN_OPTS=tc_res_tma_max_chunks=2 timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r2c_layers_orig16_rt2.log 2>&1; tail -6 gpurun_out/r2c_layers_orig16_rt2.log | head -1
HVN_OPTS=tc_res_tma_max_chunks=8 timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r2c_layers_orig16_rt8.log 2>&1; tail -6 gpurun_out/r2c_layers_orig16_rt8.log | head -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_orig256.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --workload fast64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_fast64.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f frac %.3f cnn %.1f pp %.2f conv0 %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_classes']['cnn_total_ms'], d['kernel_classes']['postproc']['ms'], d['kernel_classes']['conv0']['ms']), d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
bash tools/gpu_sanitize.sh
