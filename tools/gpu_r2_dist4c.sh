#!/bin/bash
N=$(nvidia-smi -L | wc -l)
rm -f gpurun_out/r2_c5_runs_n${N}_b.jsonl
for m in nvml smi1000 0; do
SMI=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 tools/c5_runs_probe.py 5 2>gpurun_out/r2_c5_runs.err | grep '^{' | grep '"rank": 0' >> gpurun_out/r2_c5_runs_n${N}_b.jsonl
done
cat gpurun_out/r2_c5_runs_n${N}_b.jsonl; tail -3 gpurun_out/r2_c5_runs.err
