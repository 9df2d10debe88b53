#!/bin/bash
# docstring-trace stage structure against T; half-cloud per CTA timing (would a cluster split pay?)
python tools/trace_probe.py 100,200,400,800 6 > gpurun_out/r2_trace_probe.jsonl 2> gpurun_out/r2_trace_probe.err; tail -3 gpurun_out/r2_trace_probe.err
python tools/batch_occupancy_probe.py 2 2048 40 148,296,592,1024 > gpurun_out/r2_batch_occupancy_ucsv2048.jsonl 2>&1
python tools/batch_occupancy_probe.py 2 1024 40 148,296,592,1184 > gpurun_out/r2_batch_occupancy_ucsv1024.jsonl 2>&1
python -m pytest tests -m gpu -q --deselect tests/test_theta_level.py::test_reference_docstring_trace_is_a_plausible_draw 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_rest.log
