mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest1.log
tail -30 gpurun_out/pytest1.log
timeout 600 python tools/quick_bench.py > gpurun_out/quick1.log 2>&1; echo "quick rc=$?" >> gpurun_out/quick1.log
tail -40 gpurun_out/quick1.log
