# 8-GPU runs of the θ-sharded samplers (config 5 UCSV 4096 x 4096, config 4 density-tempered SV, config 3)
mkdir -p gpurun_out
for CFG in c5 c4 c3; do
  env RESAMPLER=${RESAMPLER:-multinomial} python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/smc2_dist.py $CFG 2>&1 | tail -1
done | tee gpurun_out/dist8_v17.log
