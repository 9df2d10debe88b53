#!/bin/bash
# whole GPU suite; the N=1 bench line; its launch list under ncu; ncu --set full of the dynamically scheduled batch kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_full_d.log; tail -3 gpurun_out/r2_pytest_gpu_full_d.log
python bench.py 2> gpurun_out/r2_bench_n1_d.err | grep '^{' > gpurun_out/r2_bench_n1_d.json; tail -c 600 gpurun_out/r2_bench_n1_d.err
python bench.py --impl reference --steps 2 --warmup 1 2> /dev/null | grep '^{' > gpurun_out/r2_bench_reference_arm_d.json
python bench.py --steps 2 --warmup 3 --no-smc2 --no-cpu --no-f32 2> /dev/null | grep '^{' > gpurun_out/r2_bench_same_command_plain_d.json && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches_bench_d.csv python bench.py --steps 2 --warmup 3 --no-smc2 --no-cpu --no-f32 > gpurun_out/ncu_d.log 2>&1
python tools/batch_xsmem_probe.py 2 512 4096 40 > gpurun_out/plain_b3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:batch_kernel -c 1 -o gpurun_out/r2_ncu_batch_dyn_ucsv4096_m512 python tools/batch_xsmem_probe.py 2 512 4096 40 > gpurun_out/ncu_b3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
