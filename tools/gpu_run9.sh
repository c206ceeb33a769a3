#!/bin/bash
# single-GPU session: full GPU tests (new flood kernel), N=1 references of the sharded workloads, per-layer logs, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2f_tests.log
bash tools/gpu_run5.sh
timeout 300 python tools/gpu_diag.py layers original 16 > gpurun_out/r2f_layers_orig16.log 2>&1
timeout 300 python tools/gpu_diag.py layers fast 32 > gpurun_out/r2f_layers_fast32.log 2>&1
tail -7 gpurun_out/r2f_layers_orig16.log | head -2; tail -7 gpurun_out/r2f_layers_fast32.log | head -2
timeout 300 python tools/gpu_diag.py ppprof > gpurun_out/r2f_ppprof.log 2>&1; grep "total\|pp stats\|k_watershed" gpurun_out/r2f_ppprof.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_orig256.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --workload fast64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_fast64.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f frac %.3f cnn %.1f pp %.2f conv0 %.2f ppnuc %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_classes']['cnn_total_ms'], d['kernel_classes']['postproc']['ms'], d['kernel_classes']['conv0']['ms'], d['roofline_postproc']['on_nuclei_maps']['ms_per_step']), d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_dram_bytes.csv python tools/ncu_target.py 16 original > gpurun_out/r2_ncu_traffic.log 2>&1; echo "traffic rc=$?"
gzip -f gpurun_out/r2_launches.csv
bash tools/ncu_capture.sh r2 16 original \
  "128 2 0 1 0 0 55 xf128_d2u1conv1" \
  "64 2 0 1 0 0 2 xf64_d0u1conv1" \
  "64 2 1 0 1 0 10 rt64_d0u1conv3" \
  "64 2 1 0 1 0 15 rt64_d2u1conv3" \
  "128 3 1 0 0 0 2 res128_d3u1conv3" \
  "128 3 0 0 0 0 25 plain128" \
  "32 8 0 0 0 1 40 halo32_dense_conv2" \
  "128 3 0 0 0 1 6 halo128_conva"
mkdir -p /tmp/ncu_reps; timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_conv0_tc -s 1 -c 1 -f -o /tmp/ncu_reps/r2_conv0_tc python tools/ncu_target.py 16 original > gpurun_out/ncu_r2_conv0_tc.log 2>&1; echo "conv0_tc capture rc=$?"
ncu -i /tmp/ncu_reps/r2_conv0_tc.ncu-rep --page raw --csv > gpurun_out/r2_conv0_tc.raw.csv 2>/dev/null
ncu -i /tmp/ncu_reps/r2_conv0_tc.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_conv0_tc.source.csv.gz
du -sh gpurun_out
