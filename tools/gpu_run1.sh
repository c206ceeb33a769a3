#!/bin/bash
# round-2 GPU session 1: tests, per-layer times (both modes), the new bench lines, the north-star reference arm, chunk sweep
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests1.log
timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r2a_layers_orig16.log 2>&1
timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r2a_layers_fast32.log 2>&1
tail -7 gpurun_out/r2a_layers_orig16.log; tail -7 gpurun_out/r2a_layers_fast32.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_orig256.log 2>&1; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench_orig256.log
timeout 600 python bench.py --workload fast64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_fast64.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref_cuda.log 2>&1; echo "ref rc=$?"; tail -c 1500 gpurun_out/r2a_ref_cuda.log
for c in 64 128; do
  timeout 300 python bench.py --steps 3 --warmup 2 --chunk $c --no-cpu-baseline > gpurun_out/r2a_bench_orig256_c$c.log 2>&1; echo "chunk $c rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f frac %.3f cnn %.1f pp %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_classes']['cnn_total_ms'], d['kernel_classes']['postproc']['ms']))
    except Exception as e: print(f, 'ERR', e)
PY
