# GPU job: the whole -m gpu suite, timed
mkdir -p gpurun_out
start=$(date +%s)
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - start ))s" >> gpurun_out/pytest_full.log
tail -15 gpurun_out/pytest_full.log
