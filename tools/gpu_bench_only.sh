# GPU job: the contract bench (N=1) with every leg, timed
mkdir -p gpurun_out
start=$(date +%s)
timeout 280 python bench.py > gpurun_out/bench_n1_v22.json 2> gpurun_out/bench_n1_v22.err; echo "rc=$? wall=$(( $(date +%s) - start ))s"
tail -c 3000 gpurun_out/bench_n1_v22.json; tail -5 gpurun_out/bench_n1_v22.err
