# GPU job: the parity tests of the rows added last (guided filter, matrix Kalman, per-θ variances, multivariate IBIS, golden vectors)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_widen_guided_kalman.py -m gpu -q > gpurun_out/pytest_widen.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_widen.log
tail -30 gpurun_out/pytest_widen.log
