#!/bin/bash
# the driver's scaling contract at N GPUs (torchrun): the reference arm first, then the sharded bench line
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2> gpurun_out/r2_bench_n${N}_ref.err | grep '^{' > gpurun_out/r2_bench_n${N}_reference_arm.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 3 --warmup 1 2> gpurun_out/r2_bench_n${N}_final.err | grep '^{' > gpurun_out/r2_bench_n${N}_final.json
tail -c 600 gpurun_out/r2_bench_n${N}_final.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n${N}_final.json'))
print(d['n_gpus'], 'value %.1f G'%(d['value']/1e9), 'ms %.1f'%d['ms_per_step'], 'single %.3f s'%d['single_gpu_same_workload']['wall_s'], 'sha', d['run']['theta_sha'], d['single_gpu_same_workload']['theta_sha'])
r=json.load(open('gpurun_out/r2_bench_n${N}_reference_arm.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
