mkdir -p gpurun_out
G=${1:-2}
for cfg in tiny c3; do
  python tools/smc2_dist.py $cfg 2>&1 | tail -1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/smc2_dist.py $cfg 2>&1 | tail -1
done
