#!/bin/bash
python -m pytest tests/test_theta_level.py -m gpu -x -q -k "dynamic_scheduling" 2>&1 | tail -6 > gpurun_out/r2_dyn_sampler_tests.log; cat gpurun_out/r2_dyn_sampler_tests.log
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_multi_gpu_tests_final.log; cat gpurun_out/r2_multi_gpu_tests_final.log
