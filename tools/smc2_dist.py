"""θ-sharded SMC² / density-tempered runs under torchrun; prints timings and a checksum that must not
depend on the number of GPUs.   torchrun --nproc-per-node G tools/smc2_dist.py [config]"""
import hashlib, json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist
import sequential_monte_carlo_b200 as smc

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = smc.TorchComm() if world > 1 else None
ctx = smc.Context(local, 1998)
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"


def lg_mod(θ):
    return smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1))


def ucsv_mod(θ):
    return smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1))


def sv_mod(θ):
    return smc.SV(θ[0], θ[1], θ[2])


lg_prior = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])
ucsv_prior = smc.product_distribution([smc.Uniform(0, 1), smc.Normal(3, 2), smc.Uniform(0, 2), smc.Uniform(0, 2)])
sv_prior = smc.product_distribution([smc.Normal(0, 2), smc.Uniform(-1, 1), smc.LogNormal(-1, 1)])

if cfg == "c3":      # BASELINE config 3
    N, M, T, chain, model, prior, algo = 1024, 512, 100, 3, lg_mod, lg_prior, "smc2"
    y = smc.simulate(lg_mod([0.5, 0.9, 0.8]), T, seed=1998)[1]
elif cfg == "c3big":  # config 3 with 4096 θ (fills 8 GPUs)
    N, M, T, chain, model, prior, algo = 1024, 4096, 100, 3, lg_mod, lg_prior, "smc2"
    y = smc.simulate(lg_mod([0.5, 0.9, 0.8]), T, seed=1998)[1]
elif cfg == "c4":    # BASELINE config 4: density tempered SV 1024 θ × 2048, T=500
    N, M, T, chain, model, prior, algo = 2048, 1024, 500, 3, sv_mod, sv_prior, "dt"
    y = smc.simulate(sv_mod([-1.0, 0.9, 0.3]), T, seed=1998)[1]
elif cfg == "c5":    # BASELINE config 5: UCSV smc² 4096 θ × 4096, T=241
    N, M, T, chain, model, prior, algo = 4096, 4096, int(os.environ.get("T", 241)), 3, ucsv_mod, ucsv_prior, "smc2"
    y = smc.simulate(ucsv_mod([0.2, 3.0, 1.0, 1.0]), T, seed=1998)[1]
else:                # tiny, for the independence check
    N, M, T, chain, model, prior, algo = 128, 64, 40, 2, lg_mod, lg_prior, "smc2"
    y = smc.simulate(lg_mod([0.5, 0.9, 0.8]), T, seed=1998)[1]

resampler = os.environ.get("RESAMPLER", "multinomial")
# GUIDED=1: guided inner filters (locally optimal proposals; LG configs only — SMC(..., proposal=...), docs/SPEC.md §10)
proposal = smc.lg_optimal_proposals if os.environ.get("GUIDED") == "1" and model is lg_mod else None
s = smc.SMC(N, M, model, prior, chain, 0.5, seed=1998, ctx=ctx, comm=comm, resampler=resampler, proposal=proposal)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
plain, rejuv = [], []
if algo == "smc2":
    smc.smc2(s, y)
    for t in range(1, T):
        t1 = time.perf_counter()
        smc.smc2_step(s, y, t, verbose=False)
        (rejuv if s.rejuvenated else plain).append(time.perf_counter() - t1)
else:
    smc.density_tempered(s, y, verbose=False)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
tm = torch.tensor([wall], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
x = s.x   # local clouds
h = hashlib.sha256(np.ascontiguousarray(s.θ).tobytes()).hexdigest()[:16]
hx = hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()[:16]
allhx = [None] * world
if world > 1:
    dist.all_gather_object(allhx, hx)
else:
    allhx = [hx]
if rank == 0:
    pu = s.stats["particle_updates"] * world
    print(json.dumps(dict(cfg=cfg, algo=algo, world=world, N=N, M=M, T=T, resampler=resampler, wall_s=float(tm.item()), s_per_step=float(tm.item()) / T,
                          plain_ms=1e3 * float(np.mean(plain)) if plain else None, rejuv_ms=1e3 * float(np.mean(rejuv)) if rejuv else None,
                          n_rejuv=len(rejuv), stages=len(getattr(s, "schedule", [])), particle_updates=pu, gpups=pu / float(tm.item()) / 1e9,
                          device_ms_rank0=s.stats["device_ms"], clouds_moved_rank0=s.stats["clouds_moved"],
                          theta_sha=h, x_sha_by_rank=allhx, logZ_sum=float(np.sum(s.logZ)), ess=float(s.ess),
                          mean=[float(v) for v in smc.expected_parameters(s).ravel()])), flush=True)
s.close()
if world > 1:
    dist.destroy_process_group()
