"""SMC² timing leg of bench.py (BASELINE.json metric part 2: "SMC² 512×1024 s/step at 1–8 GPUs").

Config 3: SMC(1024, 512, lg_mod, lg_prior, 3, 0.5) on T=100 observations of lg_mod([0.5,0.9,0.8]).
θ-particles are sharded across the ranks (strong scaling: total work fixed); reports seconds per
smc²! call split into plain propagation steps and steps that rejuvenated, plus the whole-run wall.
"""
import time

import numpy as np


def run(local, rank, world, N=1024, M=512, T=100, chain=3):
    import torch
    import torch.distributed as dist
    import sequential_monte_carlo_b200 as smc

    def lg_mod(θ):
        return smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1))

    prior = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
    y = smc.simulate(lg_mod([0.5, 0.9, 0.8]), T, seed=1998)[1]
    comm = smc.TorchComm() if world > 1 else None
    ctx = smc.default_context()
    out = {}
    for rep in range(2):   # rep 0 warms up (module load, allocations)
        s = smc.SMC(N, M, lg_mod, prior, chain, 0.5, seed=1998, ctx=ctx, comm=comm)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        smc.smc2(s, y)
        plain, rejuv = [], []
        for t in range(1, T):
            t1 = time.perf_counter()
            smc.smc2_step(s, y, t, verbose=False)
            (rejuv if s.rejuvenated else plain).append(time.perf_counter() - t1)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tm = torch.tensor([wall, sum(plain), sum(rejuv)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        wall, tp, tr = (float(v) for v in tm.tolist())
        pu = s.stats["particle_updates"] * world
        out = {"workload": f"smc² + smc²! t=2..{T}: SMC({N},{M},lg_mod,lg_prior,{chain},0.5), θ sharded over {world} GPU(s), multinomial",
               "scaling": "strong", "wall_s": wall, "s_per_step_mean": wall / T,
               "s_per_plain_step": tp / max(len(plain), 1), "s_per_rejuvenation_step": tr / max(len(rejuv), 1),
               "rejuvenations": len(rejuv), "particle_updates": pu, "particle_updates_per_s": pu / wall,
               "device_ms_in_filters": s.stats["device_ms"], "sweeps": s.stats["sweeps"], "clouds_moved_rank0": s.stats["clouds_moved"],
               "posterior_mean": [float(v) for v in smc.expected_parameters(s).ravel()], "final_ess": float(s.ess)}
        s.close()
    return out
