mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
python bench.py > gpurun_out/bench21.log 2>&1; tail -c 3000 gpurun_out/bench21.log
python bench.py --steps 2 --warmup 3 --no-smc2 --no-cpu --no-f32 > gpurun_out/bench21_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-smc2 --no-cpu --no-f32 > gpurun_out/ncu21.log 2>&1
python tools/prof_step.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'sum_kernel|anc_hist_kernel|move_kernel|bounds_kernel' -s 4 -c 4 -f -o gpurun_out/prof_step_v21 python tools/prof_step.py > gpurun_out/ncu21b.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke21.log 2>&1; tail -2 gpurun_out/smoke21.log
