mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest3.log
tail -30 gpurun_out/pytest3.log
timeout 600 python tools/quick_bench.py 20 22 24 > gpurun_out/quick3.log 2>&1; echo "quick rc=$?" >> gpurun_out/quick3.log
head -12 gpurun_out/quick3.log
