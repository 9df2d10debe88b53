mkdir -p gpurun_out
python tools/probe_fetch.py > gpurun_out/probe_fetch.log 2>&1; cat gpurun_out/probe_fetch.log
python bench.py --steps 2 --warmup 1 --no-smc2 > gpurun_out/bench6_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-smc2 > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log | cut -c1-200
