#!/bin/bash
# round-2 multi-GPU session (gpurun --gpus N): bench at N ranks, tile4k (configs[3]) and a reduced WSI (configs[4]) sharded
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_orig256_n$N.log 2>&1; echo "bench n$N rc=$?"
timeout 600 $TR bench.py --workload tile4k --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_tile4k_n$N.log 2>&1; echo "tile4k n$N rc=$?"
timeout 900 $TR bench.py --workload wsi40k --size ${WSI_SIZE:-12000} --gpus $N > gpurun_out/r2_wsi_n$N.log 2>&1; echo "wsi n$N rc=$?"
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_tile_driver.py -m gpu -q -k two_gpus > gpurun_out/r2_test_2gpu.log 2>&1; echo "2gpu test rc=$?"; tail -3 gpurun_out/r2_test_2gpu.log; fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_*_n*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f e2e %.1f ms/step %.1f n_gpus %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['n_gpus']), d.get('inst_map_sha1', d.get('json_sha1','')), d.get('instances',''))
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1200:])
PY
