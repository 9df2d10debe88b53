#!/bin/bash
python -m pytest tests/test_theta_level.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_theta_tests_f.log; cat gpurun_out/r2_theta_tests_f.log
for i in 1 2; do python tools/c3_probe.py c3 5 >> gpurun_out/r2_c3_probe_b.jsonl 2>> gpurun_out/r2_c3_probe_b.err; done
python tools/c3_probe.py c3_multinomial 5 >> gpurun_out/r2_c3_probe_b.jsonl 2>> gpurun_out/r2_c3_probe_b.err
