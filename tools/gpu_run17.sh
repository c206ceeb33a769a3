#!/bin/bash
# r3h: final single-GPU session of round 2, session 2: GPU tests, smoke, bench lines, reference arm, launch list, traffic, RS capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3h_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3h_bench_orig256.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --workload fast64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3h_bench_fast64.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r3h_ref.log 2>&1; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3h_bench_*.log')) + ['gpurun_out/r3h_ref.log']:
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'frac', d.get('roofline',{}).get('frac'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_dram_bytes.csv python tools/ncu_target.py 16 original > gpurun_out/r3_ncu_traffic.log 2>&1; echo "traffic rc=$?"
gzip -f gpurun_out/r3_launches.csv
mkdir -p /tmp/ncu_reps
for spec in "2 rs5_u3dense_u1conv2" "20 rs5_u2dense_u0conv2"; do
  set -- $spec
  rep=/tmp/ncu_reps/r3_conv_$2
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_conv_rs -s $1 -c 1 -f -o $rep python tools/ncu_target.py 16 original > gpurun_out/ncu_r3_$2.log 2>&1; echo "$2 rc=$?"
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/r3_conv_$2.raw.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r3_conv_$2.source.csv.gz
done
timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r3h_layers_orig16.log 2>&1
timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r3h_layers_fast32.log 2>&1
python tools/layer_classes.py gpurun_out/r3h_layers_orig16.log gpurun_out/r3h_layers_fast32.log
du -sh gpurun_out
