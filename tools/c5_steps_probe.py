"""Where do the occasional slow runs of the θ-sharded config 5 lose their time?  Per run: wall, and the slowest smc²! calls with their
step index and kind (torchrun --nproc-per-node G tools/c5_steps_probe.py [runs])"""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist
import sequential_monte_carlo_b200 as smc
from sequential_monte_carlo_b200 import bench_smc2 as B

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = smc.TorchComm() if world > 1 else None
ctx = smc.Context(local, 1998)
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(B.CONFIGS["c5"])
model, prior, truth = B._setup(smc, cfg["kind"])
y = smc.simulate(model(truth), cfg["T"], seed=1998)[1]
for rep in range(runs + 2):
    s = smc.SMC(cfg["N"], cfg["M"], model, prior, cfg["chain"], 0.5, seed=1998, ctx=ctx, comm=comm, resampler="systematic", engine="device")
    if world > 1:
        dist.barrier()
    ctx.synchronize()
    t0 = time.perf_counter()
    smc.smc2(s, y)
    steps = []
    for t in range(1, cfg["T"]):
        t1 = time.perf_counter()
        smc.smc2_step(s, y, t, verbose=False)
        steps.append((time.perf_counter() - t1, t, bool(s.rejuvenated)))
    θ = s.θ
    ctx.synchronize()
    wall = time.perf_counter() - t0
    st = s._eng.stats()
    s.close()
    if rep >= 2:
        top = sorted(steps, reverse=True)[:4]
        rej = sorted([(round(a, 4), t) for a, t, r in steps if r], key=lambda v: v[1])
        print(json.dumps({"rank": rank, "run": rep - 2, "wall_s": round(wall, 4), "slowest": [(round(a, 4), t, r) for a, t, r in top],
                          "plain_sum_s": round(sum(a for a, t, r in steps if not r), 4), "rejuvenations": rej}), flush=True)
if world > 1:
    dist.destroy_process_group()
