# GPU job: the parity tests of the rows added last (guided filter, matrix Kalman, per-θ variances, multivariate IBIS) + their timings
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_widen_guided_kalman.py -m gpu -q > gpurun_out/pytest_widen.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_widen.log
tail -30 gpurun_out/pytest_widen.log
timeout 100 python tools/widen_bench.py > gpurun_out/widen_bench.json 2> gpurun_out/widen_bench.err; tail -c 1500 gpurun_out/widen_bench.json; tail -3 gpurun_out/widen_bench.err
