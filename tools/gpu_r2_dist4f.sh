#!/bin/bash
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 tools/c5_steps_probe.py 10 2>gpurun_out/r2_c5_steps.err | grep '^{' > gpurun_out/r2_c5_steps_n${N}.jsonl
python - <<PY
import json
rows=[json.loads(l) for l in open('gpurun_out/r2_c5_steps_n${N}.jsonl')]
for r in rows:
    if r['rank'] in (0,1): print(r['rank'], r['run'], r['wall_s'], r['slowest'][:3], r['plain_sum_s'])
PY
tail -3 gpurun_out/r2_c5_steps.err
