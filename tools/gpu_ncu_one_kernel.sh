mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'anc_kernel' -s 2 -c 1 -f -o gpurun_out/prof_anc_v13 python tools/prof_step.py > gpurun_out/ncu8.log 2>&1
tail -2 gpurun_out/ncu8.log
