#!/bin/bash
# round 2, first measurement of the device-resident sampler: θ-level + ABI tests, then the N=1 bench line
set -x
python -m pytest tests/test_theta_level.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_theta.log
cat gpurun_out/r2_theta.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -c 3000 gpurun_out/r2_bench_n1_a.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1_a.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])
print(json.dumps(d.get('smc2'),indent=1)[:6000])
print(d.get('config1_latency'))
"
