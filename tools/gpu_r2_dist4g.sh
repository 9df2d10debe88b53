#!/bin/bash
N=$(nvidia-smi -L | wc -l)
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 tools/c5_steps_probe.py 12 2>gpurun_out/r2_c5_steps.err | grep '^{' > gpurun_out/r2_c5_steps_n${N}_b.jsonl
python - <<PY
import json
rows=[json.loads(l) for l in open('gpurun_out/r2_c5_steps_n${N}_b.jsonl')]
for r in rows:
    if r['rank']==0: print(r['run'], r['wall_s'], r['slowest'][:2], r['plain_sum_s'])
PY
tail -3 gpurun_out/r2_c5_steps.err
