#!/bin/bash
# 8 GPUs after the dynamic chunk scheduling: NCCL parity on 2 of them, then the sharded bench line (config 5 headline)
nvidia-smi -L | wc -l
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_multi_gpu_test_b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 3 --warmup 1 2> gpurun_out/r2_bench_n8_b.err | grep '^{' > gpurun_out/r2_bench_n8_b.json
tail -c 800 gpurun_out/r2_bench_n8_b.err
SMCB_BATCH_CHUNK=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 3 --warmup 1 2> gpurun_out/r2_bench_n8_b_static.err | grep '^{' > gpurun_out/r2_bench_n8_b_static.json
