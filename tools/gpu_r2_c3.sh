#!/bin/bash
for c in c3 c4; do
  for e in 0 auto; do
    if [ $e = auto ]; then unset SMCB_BATCH_CHUNK; else export SMCB_BATCH_CHUNK=$e; fi
    python tools/c3_probe.py $c 4 >> gpurun_out/r2_c3_probe.jsonl 2>> gpurun_out/r2_c3_probe.err
  done
done
unset SMCB_BATCH_CHUNK
python -m pytest tests/test_widen_guided_kalman.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_widen_tests_e.log
python -m pytest tests/test_theta_level.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -4 >> gpurun_out/r2_widen_tests_e.log
