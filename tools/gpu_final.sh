#!/bin/bash
# what the driver does at round end, on one GPU: GPU tests, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2g_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2g_smoke.log
/usr/bin/time -v timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; grep "Elapsed (wall" gpurun_out/r2g_bench.err
/usr/bin/time -v timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_ref.log 2> gpurun_out/r2g_ref.err; echo "ref rc=$?"; grep "Elapsed (wall" gpurun_out/r2g_ref.err
python - <<'PY'
import json
a=json.loads([x for x in open('gpurun_out/r2g_bench.log') if x.startswith('{')][-1]); b=json.loads([x for x in open('gpurun_out/r2g_ref.log') if x.startswith('{')][-1])
print('ours value %.1f e2e %.1f frac %.3f | reference %.1f | ratio e2e %.2f' % (a['value'], a['e2e']['value'], a['roofline']['frac'], b['value'], a['e2e']['value']/b['value']))
print('cpu_baseline', a.get('cpu_baseline',{}).get('value'), a.get('cpu_baseline',{}).get('sample'))
print('roofline_postproc', json.dumps(a['roofline_postproc'])[:600])
print('clocks', a['clocks'], 'traffic', a['roofline']['traffic'])
PY
