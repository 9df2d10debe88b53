mkdir -p gpurun_out
run() { G=$1; shift; if [ "$G" = 1 ]; then "$@" python tools/smc2_dist.py $CFG 2>&1 | tail -1; else "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/smc2_dist.py $CFG 2>&1 | tail -1; fi; }
for CFG in c3 c3big c4 c5; do
  for G in 1 2 4 8; do
    run $G env RESAMPLER=${RESAMPLER:-multinomial}
  done
done
