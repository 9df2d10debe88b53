#!/bin/bash
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 4 --warmup 1 2> gpurun_out/r2_bench_n${N}_final.err | grep '^{' > gpurun_out/r2_bench_n${N}_final.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n${N}_final.json'))
print(d['n_gpus'], 'value %.1f G'%(d['value']/1e9), 'ms %.1f'%d['ms_per_step'], 'single %.3f s'%d['single_gpu_same_workload']['wall_s'], 'sha', d['run']['theta_sha'])
print(d['run'].get('walls_s'), d['run'].get('device_spans_s'))
PY
