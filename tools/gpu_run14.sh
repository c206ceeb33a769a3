#!/bin/bash
# r3e: full GPU tests, smoke, bench (default orig256, fast64, chunk sweep), launch list + DRAM traffic + captures of the new kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3e_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3e_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3e_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3e_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3e_bench_orig256.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --workload fast64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_bench_fast64.log 2>&1; echo "bench rc=$?"
for c in 16 64; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $c > gpurun_out/r3e_bench_orig256_chunk$c.log 2>&1; echo "bench chunk $c rc=$?"; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opt tc_ar=0 --opt tc_prefetch=0 --opt tc_xf_trunc=0 --opt tc_xf_early=0 > gpurun_out/r3e_bench_orig256_old.log 2>&1; echo "bench old rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3e_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f frac %.3f cnn %.1f pp %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_classes']['cnn_total_ms'], d['kernel_classes']['postproc']['ms']), d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_dram_bytes.csv python tools/ncu_target.py 16 original > gpurun_out/r3_ncu_traffic.log 2>&1; echo "traffic rc=$?"
gzip -f gpurun_out/r3_launches.csv
bash tools/ncu_capture.sh r3 16 original "128 2 0 1 0 0 55 xf128_d2u1conv1" "64 2 0 1 0 0 2 xf64_d0u1conv1"
for spec in "15 ar_d2u1conv3" "5 ar_d1u1conv3" "1 ar_d0u1conv3"; do
  set -- $spec
  rep=/tmp/ncu_reps/r3_conv_$2
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_conv_ar -s $1 -c 1 -f -o $rep python tools/ncu_target.py 16 original > gpurun_out/ncu_r3_$2.log 2>&1; echo "$2 rc=$?"
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/r3_conv_$2.raw.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r3_conv_$2.source.csv.gz
done
du -sh gpurun_out
