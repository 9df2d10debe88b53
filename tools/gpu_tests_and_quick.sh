mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest7.log
tail -3 gpurun_out/pytest7.log
timeout 300 python tools/quick_bench.py 22 24 2>&1 | grep -v "multinomial\|stratified" | head -4 | cut -c1-400
