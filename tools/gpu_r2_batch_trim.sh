#!/bin/bash
python -m pytest tests/test_gpu_batch.py tests/test_theta_level.py -m gpu -x -q 2>&1 | tail -3
rm -f gpurun_out/r2_chunk_trim.jsonl
for spec in "2 4096 100 512,148" "0 1024 100 512" "1 2048 100 1024"; do
  python tools/chunk_probe.py $spec 0,auto >> gpurun_out/r2_chunk_trim.jsonl 2>> gpurun_out/r2_chunk.err
done
python - <<PY
import json
for l in open('gpurun_out/r2_chunk_trim.jsonl'):
    d=json.loads(l); print(d['kind'],d['N'],d['M'],d['chunk'],'sys %.3f mn %.3f'%(d['systematic_ms'],d['multinomial_ms']), d['systematic_sha'])
PY
python tools/c3_probe.py c5 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', d['wall_s'], d['theta_sha'])"
