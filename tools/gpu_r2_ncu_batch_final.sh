#!/bin/bash
python tools/batch_xsmem_probe.py 2 512 4096 40 > gpurun_out/plain_b4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:batch_kernel -c 1 -f -o gpurun_out/r2_ncu_batch_dyn_ucsv4096_m512_final python tools/batch_xsmem_probe.py 2 512 4096 40 > gpurun_out/ncu_b4.log 2>&1
cat gpurun_out/plain_b4.log; ls -la gpurun_out/r2_ncu_batch_dyn_ucsv4096_m512_final.ncu-rep
