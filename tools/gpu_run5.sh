#!/bin/bash
# single-GPU references for the sharded workloads: tile4k and the reduced WSI at N=1 (hashes must match the N-rank runs)
mkdir -p gpurun_out
timeout 600 python bench.py --workload tile4k --steps 2 --warmup 1 > gpurun_out/r2_tile4k_n1.log 2>&1; echo "tile4k n1 rc=$?"
timeout 900 python bench.py --workload wsi40k --size ${WSI_SIZE:-12000} > gpurun_out/r2_wsi_n1.log 2>&1; echo "wsi n1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_*_n1.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print(f, 'value %.1f ms/step %.1f' % (d['value'], d['ms_per_step']), d.get('inst_map_sha1', d.get('json_sha1','')), d.get('instances',''))
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1200:])
PY
