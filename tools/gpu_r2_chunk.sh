#!/bin/bash
# dynamic chunk scheduling: bit-identity + timing; the rewritten docstring-trace test
python tools/chunk_probe.py 2 4096 100 512,148,300 > gpurun_out/r2_chunk_ucsv4096.jsonl 2> gpurun_out/r2_chunk.err
python tools/chunk_probe.py 0 1024 100 512,64,200 >> gpurun_out/r2_chunk_lg1024.jsonl 2>> gpurun_out/r2_chunk.err
python tools/chunk_probe.py 1 2048 100 1024,128,300 >> gpurun_out/r2_chunk_sv2048.jsonl 2>> gpurun_out/r2_chunk.err
python tools/chunk_probe.py 0 301 37 333 0,3,5,auto >> gpurun_out/r2_chunk_ragged.jsonl 2>> gpurun_out/r2_chunk.err
tail -5 gpurun_out/r2_chunk.err
python -m pytest tests/test_theta_level.py -m gpu -x -q -k "docstring" 2>&1 | tail -15 > gpurun_out/r2_trace_test.log
python -m pytest tests -m gpu -x -q --deselect tests/test_theta_level.py::test_reference_docstring_trace_is_a_plausible_draw 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_chunk.log
