#!/bin/bash
# whole GPU suite + the N=1 bench line
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_full_b.log; cat gpurun_out/r2_pytest_gpu_full_b.log
python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_n1_c.err | grep '^{' > gpurun_out/r2_bench_n1_c.json
tail -c 1500 gpurun_out/r2_bench_n1_c.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1_c.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])
print(json.dumps(d.get('multinomial'),indent=1))
print(d.get('config1_latency'))
for k in ('c3','c4','c5'):
    c=d['smc2'][k]; print(k,{q:c.get(q) for q in ('wall_s','s_per_plain_step','s_per_rejuvenation_step','breakdown_ms')})
"
