#!/bin/bash
# 2 GPUs: NCCL parity of the θ-sharded device sampler against one rank, then the sharded bench line
set -x
nvidia-smi -L
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_multi_gpu_test.log
cat gpurun_out/r2_multi_gpu_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err
tail -c 2500 gpurun_out/r2_bench_n2_a.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2_a.json'))
print(json.dumps({k:v for k,v in d.items() if k not in ('config',)},indent=1)[:9000])
"
