#!/bin/bash
python tools/mn_probe.py 24 4 > gpurun_out/plain_mn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mn_c -s 2 -c 2 -f -o gpurun_out/r2_ncu_mn_src python tools/mn_probe.py 24 4 > gpurun_out/ncu_mn.log 2>&1
ls -la gpurun_out/r2_ncu_mn_src.ncu-rep; tail -2 gpurun_out/ncu_mn.log
