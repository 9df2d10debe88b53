#!/bin/bash
# ncu --set full: the two-level multinomial kernels at N=2^24, and batch_kernel on config 5's / config 3's inner shapes
set -x
python tools/mn_probe.py 24 4 > gpurun_out/plain_mn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mn_ -s 3 -c 6 -o gpurun_out/r2_ncu_mn_v2 python tools/mn_probe.py 24 4 > gpurun_out/ncu_mn.log 2>&1
python tools/batch_xsmem_probe.py 2 148 4096 20 > gpurun_out/plain_b1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:batch_kernel -c 1 -o gpurun_out/r2_ncu_batch_ucsv4096_m148 python tools/batch_xsmem_probe.py 2 148 4096 20 > gpurun_out/ncu_b1.log 2>&1
python tools/batch_xsmem_probe.py 0 512 1024 100 > gpurun_out/plain_b2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:batch_kernel -c 1 -o gpurun_out/r2_ncu_batch_lg1024_m512 python tools/batch_xsmem_probe.py 0 512 1024 100 > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_mn.log gpurun_out/ncu_b1.log gpurun_out/ncu_b2.log
