# 1-GPU experiment: config 5 (UCSV 4096 θ × 4096) with the clouds in shared memory (1 CTA/SM) against clouds in global memory / L2 (2 CTAs/SM)
mkdir -p gpurun_out
( T=100 python tools/smc2_dist.py c5 2>&1 | tail -1
  SMCB_BATCH_X_SMEM=0 T=100 python tools/smc2_dist.py c5 2>&1 | tail -1 ) | tee gpurun_out/c5_xsmem_v22.jsonl | cut -c1-420
