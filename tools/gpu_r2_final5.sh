#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -1 gpurun_out/r2_smoke_final.log
python bench.py 2> gpurun_out/r2_bench_n1_final.err | grep '^{' > gpurun_out/r2_bench_n1_final.json; tail -c 300 gpurun_out/r2_bench_n1_final.err
python bench.py --impl reference --steps 2 --warmup 1 2> /dev/null | grep '^{' > gpurun_out/r2_bench_reference_arm_final.json
python bench.py --impl reference --gpus 8 --steps 1 --warmup 1 2> /dev/null | grep '^{' > gpurun_out/r2_bench_reference_arm_n8_final.json
