# C2 roofline sweep N = 2^20 .. 2^24 (all three resamplers, binary64 and binary32-state tiers) + stepping-API cost
mkdir -p gpurun_out
timeout 600 python tools/quick_bench.py 20 21 22 23 24 > gpurun_out/quick_sweep_v21.jsonl 2>&1
grep -c N gpurun_out/quick_sweep_v21.jsonl
timeout 300 python tools/step_api_bench.py 2>&1 | tee gpurun_out/step_api_v21.log
