#!/bin/bash
python -m pytest tests/test_gpu_filter.py -m gpu -x -q -k "multinomial or two_level or fuzz or resample or degenerate" 2>&1 | tail -6 > gpurun_out/r2_mn5_tests.log
python tools/mn_time.py 22 24 > gpurun_out/r2_mn5_time.jsonl 2> gpurun_out/r2_mn5_time.err
python tools/degenerate_bench.py > gpurun_out/r2_mn5_degenerate.jsonl 2>&1
