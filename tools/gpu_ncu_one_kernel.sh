# one `ncu --set full` capture of one kernel of the step: KERNEL=sum_kernel|bounds_kernel|anc_hist_kernel|move_kernel  OUT=name
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-anc_hist_kernel}" -s 2 -c 1 -f -o gpurun_out/${OUT:-prof_one} python tools/prof_step.py > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
