mkdir -p gpurun_out
timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -2
( for X in 1 0; do SMCB_BATCH_X_SMEM=$X timeout 40 python tools/batch_xsmem_probe.py 0 1024 8192 40; done
  for X in 1 0; do SMCB_BATCH_X_SMEM=$X timeout 40 python tools/batch_xsmem_probe.py 1 1024 8192 40; done ) 2>&1 | tee gpurun_out/batch_xsmem_probe_v22.jsonl
