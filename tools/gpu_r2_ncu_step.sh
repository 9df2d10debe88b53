#!/bin/bash
python tools/prof_step.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'sum_kernel|anc_hist_kernel|move_kernel|bounds_kernel' -s 4 -c 4 -f -o gpurun_out/r2_ncu_step_src python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out/r2_ncu_step_src.ncu-rep; tail -2 gpurun_out/ncu_step.log
