mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest7.log
tail -15 gpurun_out/pytest7.log
timeout 600 python tools/quick_bench.py 20 22 24 2>&1 | grep -v multinomial | head -8 > gpurun_out/quick7.log
cat gpurun_out/quick7.log
