mkdir -p gpurun_out
python tools/prof_step.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'sum_kernel|prop2_kernel' -s 2 -c 2 -o gpurun_out/prof_step_v3 python tools/prof_step.py > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
