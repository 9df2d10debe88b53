#!/bin/bash
# 8 GPUs: the sharded bench line (config 5 headline, configs 3/4, replicas, single-GPU reference of the same workload)
set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 3 --warmup 1 2> gpurun_out/r2_bench_n8_a.err | grep '^{' > gpurun_out/r2_bench_n8_a.json
tail -c 1500 gpurun_out/r2_bench_n8_a.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8_a.json'))
print(json.dumps({k:v for k,v in d.items() if k not in ('config','roofline')},indent=1)[:7000])
"
