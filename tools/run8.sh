mkdir -p gpurun_out
python tools/prof_step.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'sum_kernel|anc_kernel|move_kernel|bounds_kernel' -s 4 -c 4 -f -o gpurun_out/prof_step_v9 python tools/prof_step.py > gpurun_out/ncu8.log 2>&1
tail -3 gpurun_out/ncu8.log
