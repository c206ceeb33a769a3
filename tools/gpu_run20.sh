#!/bin/bash
# r3k: final check of the committed build: bench lines first (short), then the full GPU suite and smoke
mkdir -p gpurun_out
timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_bench_orig256.log 2>&1; echo "bench rc=$?"
timeout 100 python bench.py --workload fast64 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_bench_fast64.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3k_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f frac %.4f ach %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved']), d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r3k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3k_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3k_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3k_smoke.log
