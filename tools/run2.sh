mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest2.log
tail -30 gpurun_out/pytest2.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke2.log
tail -5 gpurun_out/smoke2.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench2.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench2.log
tail -5 gpurun_out/bench2.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench2_ref.log 2>&1
tail -3 gpurun_out/bench2_ref.log
