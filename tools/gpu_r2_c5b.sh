#!/bin/bash
python -m pytest tests/test_gpu_batch.py tests/test_theta_level.py -m gpu -x -q 2>&1 | tail -3
for c in c5 c4; do
  SMCB_BATCH_CHUNK=0 python tools/c3_probe.py $c 2 >> gpurun_out/r2_c5_probe_c.jsonl 2>> gpurun_out/r2_c5_probe_c.err
  python tools/c3_probe.py $c 2 >> gpurun_out/r2_c5_probe_c.jsonl 2>> gpurun_out/r2_c5_probe_c.err
done
python - <<PY
import json
for l in open('gpurun_out/r2_c5_probe_c.jsonl'):
    d=json.loads(l); print(d['config'],d['chunk_env'],'wall %.1f ms'%(1e3*d['wall_s']), d['theta_sha'], {k:round(v,1) for k,v in d['breakdown_ms'].items()})
PY
