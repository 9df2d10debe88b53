mkdir -p gpurun_out
python tools/prof_step.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'sum_kernel|prop2_kernel|bounds_kernel' -s 3 -c 3 -f -o gpurun_out/prof_step_v5 python tools/prof_step.py > gpurun_out/ncu8.log 2>&1
tail -3 gpurun_out/ncu8.log
