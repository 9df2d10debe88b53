#!/bin/bash
# automatic chunk heuristic across shapes; rewritten trace test; UCSV guided tests; whole suite
rm -f gpurun_out/r2_chunk_auto.jsonl
for spec in "2 4096 100 512,300" "0 1024 100 512,200,600" "1 2048 100 1024,300" "0 301 37 333" "2 1024 60 600" "0 8192 40 200"; do
  python tools/chunk_probe.py $spec 0,auto >> gpurun_out/r2_chunk_auto.jsonl 2>> gpurun_out/r2_chunk.err
done
python -m pytest tests/test_theta_level.py -m gpu -x -q -k "docstring" 2>&1 | tail -15 > gpurun_out/r2_trace_test.log
python -m pytest tests/test_widen_guided_kalman.py -m gpu -x -q -k "ucsv or errors" 2>&1 | tail -15 > gpurun_out/r2_ucsv_guided_test.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_full_c.log
