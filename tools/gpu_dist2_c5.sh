# 2-GPU job: config 5 (UCSV smc² 4096 θ × 4096, T=241) at 1 and 2 GPUs — same θ hash, clouds crossing GPUs after sorted θ-ancestors
mkdir -p gpurun_out
( python tools/smc2_dist.py c5 2>&1 | tail -1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/smc2_dist.py c5 2>&1 | tail -1 ) | tee gpurun_out/dist2_c5_v22.jsonl | cut -c1-700
