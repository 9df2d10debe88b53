# GPU job (one GPU): ncu --set full of the batched kernel on config 5's inner shape (UCSV 4096 particles per θ) with the clouds in
# shared memory and in global memory / L2, and of the guided batched kernel — the three launches DESIGN.md §8 asks about first.
mkdir -p gpurun_out
for X in 1 0; do
  SMCB_BATCH_X_SMEM=$X timeout 60 python tools/batch_xsmem_probe.py 2 592 4096 20 > gpurun_out/probe_ucsv_x$X.json 2>&1 || exit 1   # runs clean first
  SMCB_BATCH_X_SMEM=$X timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -c 2 \
    -o gpurun_out/ncu_batch_ucsv4096_x$X python tools/batch_xsmem_probe.py 2 592 4096 20 > gpurun_out/ncu_batch_x$X.log 2>&1
done
timeout 60 python tools/widen_bench.py > gpurun_out/widen_bench.json 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -c 4 \
  -o gpurun_out/ncu_batch_guided python tools/widen_bench.py > gpurun_out/ncu_batch_guided.log 2>&1
ls -la gpurun_out/*.ncu-rep
