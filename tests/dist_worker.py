"""One rank of the multi-GPU parity test (tests/test_multi_gpu.py): runs a θ-sharded sampler on this rank's GPU through the
library's own NCCL communicator (smcb_comm_init) and writes what it holds — θ, ω, logZ (replicated) and its slice of the
clouds — to <out>.rank<r>.npz.  Launched by torch.distributed.run; torch.distributed (gloo) only carries the NCCL id."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_sampler(smc, which, ctx, comm, engine="device"):
    if which == "lg":
        pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
        model = lambda θ: smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1))
        y = smc.simulate(model([0.5, 0.9, 0.8]), 40, seed=1998)[1]
        return smc.SMC(256, 64, model, pg, 3, 0.5, seed=11, resampler="systematic", ctx=ctx, comm=comm, engine=engine), y, "smc2"
    if which == "lg_dyn":     # more θ per GPU than resident CTAs (592 of 256 threads): the rejuvenation sweeps are dynamically scheduled
        pg = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
        model = lambda θ: smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1))
        y = smc.simulate(model([0.5, 0.9, 0.8]), 48, seed=1998)[1]
        return smc.SMC(1024, 1400, model, pg, 2, 0.5, seed=5, resampler="systematic", ctx=ctx, comm=comm, engine=engine), y, "smc2"
    if which == "ucsv":
        pg = smc.product_distribution([smc.Uniform(0, 1), smc.Normal(3, 2), smc.Uniform(0, 2), smc.Uniform(0, 2)])
        model = lambda θ: smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1))
        y = smc.simulate(model([0.2, 3.0, 1.0, 1.0]), 24, seed=1998)[1]
        return smc.SMC(512, 32, model, pg, 2, 0.5, seed=3, ctx=ctx, comm=comm, engine=engine), y, "smc2"
    pg = smc.product_distribution([smc.Normal(0, 2), smc.Uniform(-1, 1), smc.LogNormal(-1, 1)])
    model = lambda θ: smc.SV(θ[0], θ[1], θ[2])
    y = smc.simulate(model([-1.0, 0.9, 0.3]), 50, seed=1998)[1]
    return smc.SMC(256, 48, model, pg, 2, 0.5, seed=8, resampler="stratified", ctx=ctx, comm=comm, engine=engine), y, "dt"


def run(smc, s, y, mode):
    rejuv = []
    if mode == "dt":
        smc.density_tempered(s, y, verbose=False)
        rejuv = [ξ for ξ, _ in s.schedule]
    else:
        smc.smc2(s, y)
        for t in range(1, len(y)):
            smc.smc2_step(s, y, t, verbose=False)
            if s.rejuvenated:
                rejuv.append(t)
    return rejuv


def main():
    which, out = sys.argv[1], sys.argv[2]
    import torch.distributed as dist
    import sequential_monte_carlo_b200 as smc
    from sequential_monte_carlo_b200 import smc_samplers as ss
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = smc.Context(device=int(os.environ.get("LOCAL_RANK", rank)), seed=1998)
    comm = ss.NcclComm.from_torch(ctx)
    s, y, mode = build_sampler(smc, which, ctx, comm)
    assert s.engine == "device" and s.Mloc == s.M // world
    rejuv = run(smc, s, y, mode)
    stats = s._eng.stats()
    np.savez(f"{out}.rank{rank}.npz", theta=s.θ, omega=s.ω, logZ=s.logZ, x=s.x, w=s.w, rejuv=np.array(rejuv, np.float64), ess=s.ess,
             moved=stats["clouds_moved"], means=ss.state_means(s))
    s.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
