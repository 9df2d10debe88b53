mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest5.log
tail -5 gpurun_out/pytest5.log
timeout 600 python tools/quick_bench.py 20 22 24 2>&1 | grep -v multinomial | head -8 > gpurun_out/quick5.log
cat gpurun_out/quick5.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench5.log 2>&1; tail -1 gpurun_out/bench5.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['avg_us_per_launch']); print(d['e2e']); print(d['smc2'])"
