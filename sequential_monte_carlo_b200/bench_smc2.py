"""θ-level legs of bench.py: the BASELINE.json configurations that shard over θ, run through the public
sampler API on the device-resident engine (smcb_sampler_*, csrc/smcb_sampler.cu):

  c3  SMC(1024, 512, lg_mod, lg_prior, 3, 0.5), smc² + smc²! over T = 100            (BASELINE configs[2], README.md:88-103)
  c4  density_tempered, stochastic volatility, 1024 θ × 2048 state particles, T = 500  (configs[3])
  c5  smc², 4-parameter UCSV (examples/inflation_example.jl:229-256 shape), 4096 θ × 4096, T = 241  (configs[4])

θ-particles are sharded over the ranks (strong scaling: the job is fixed).  One "step" = one whole run of the sampler over
its series.  Every run reports wall time, particle-updates/s of the whole job, where the time went (device time in the inner
filters, in the all-gathers, in cloud moves, in the θ-level kernels — CUDA events on the library's stream — and the
remainder: host control + launch gaps), and a hash of θ that must not depend on the number of GPUs.
"""
import hashlib
import time

import numpy as np

CONFIGS = {
    "c3": dict(kind="lg", N=1024, M=512, T=100, chain=3, algo="smc2", bytes_per_update=56,
               name="smc² + smc²! t=2..100: SMC(1024, 512, lg_mod, lg_prior, 3, 0.5) (BASELINE.json configs[2])"),
    "c4": dict(kind="sv", N=2048, M=1024, T=500, chain=3, algo="dt", bytes_per_update=56,
               name="density_tempered, stochastic volatility, 1024 θ × 2048 state particles, T=500 (BASELINE.json configs[3])"),
    "c5": dict(kind="ucsv", N=4096, M=4096, T=241, chain=3, algo="smc2", bytes_per_update=88,
               name="smc² + smc²! t=2..241: UCSV (inflation example shape), 4096 θ × 4096 state particles, chain 3 (BASELINE.json configs[4])"),
}


def _setup(smc, kind):
    if kind == "lg":
        model = lambda θ: smc.StateSpaceModel(smc.LinearGaussian(θ[0], 1.0, θ[1], θ[2], 0.0), (1, 1))   # README.md:75-78
        prior = smc.product_distribution([smc.TruncatedNormal(0, 1, -1, 1), smc.LogNormal(), smc.LogNormal()])   # README.md:81-85
        truth = [0.5, 0.9, 0.8]
    elif kind == "sv":
        model = lambda θ: smc.SV(θ[0], θ[1], θ[2])
        prior = smc.product_distribution([smc.Normal(0, 2), smc.Uniform(-1, 1), smc.LogNormal(-1, 1)])
        truth = [-1.0, 0.9, 0.3]
    else:
        model = lambda θ: smc.StateSpaceModel(smc.UCSV(θ[0], θ[1], (θ[2], θ[3])), (3, 1))              # inflation_example.jl:229-232
        prior = smc.product_distribution([smc.Uniform(0, 1), smc.Normal(3, 2), smc.Uniform(0, 2), smc.Uniform(0, 2)])   # :234-239
        truth = [0.2, 3.0, 1.0, 1.0]
    return model, prior, truth


def run_config(name, ctx, comm, rank, world, steps=1, warmup=1, resampler="systematic", barrier=None, T=None, M=None):
    """`warmup` untimed runs, then `steps` timed runs of configuration `name`, θ sharded over `world` ranks.  Returns the
    per-run averages (max over ranks is taken by the caller through `reduce_max`)."""
    import sequential_monte_carlo_b200 as smc
    if name.endswith("_multinomial"):        # the same configuration with the reference's own inner resampling law (particles.jl:117)
        name, resampler = name[: -len("_multinomial")], "multinomial"
    cfg = dict(CONFIGS[name])
    if T:
        cfg["T"] = int(T)
    if M:
        cfg["M"] = int(M)
    model, prior, truth = _setup(smc, cfg["kind"])
    y = smc.simulate(model(truth), cfg["T"], seed=1998)[1]
    walls, spans, plains, rejuvs, stats_sum, out = [], [], [], [], None, {}
    if world > 1:
        warmup = max(warmup, 2)        # the first sharded run also builds NCCL's point-to-point channels: never the profiled one
    for rep in range(warmup + steps):
        s = smc.SMC(cfg["N"], cfg["M"], model, prior, cfg["chain"], 0.5, seed=1998, ctx=ctx, comm=comm, resampler=resampler, engine="device")
        s._eng.set_profiling(rep == warmup - 1)     # the breakdown comes from the last (untimed) warm-up run: the events cost ~10 µs per step
        if barrier:
            barrier()
        ctx.synchronize()
        t0 = time.perf_counter()
        plain, rejuv = [], []
        if cfg["algo"] == "smc2":
            smc.smc2(s, y)
            for t in range(1, cfg["T"]):
                t1 = time.perf_counter()
                smc.smc2_step(s, y, t, verbose=False)
                (rejuv if s.rejuvenated else plain).append(time.perf_counter() - t1)
        else:
            smc.density_tempered(s, y, verbose=False)
        θ = s.θ                       # the D2H read of the result (θ, ω, logZ) is inside the timed region
        ctx.synchronize()
        wall = time.perf_counter() - t0
        st = s._eng.stats()
        if rep == warmup - 1:
            prof = {k: st[k] for k in ("filter_ms", "allgather_ms", "exchange_ms", "theta_ms")}
            prof_wall = wall
        elif rep >= warmup:
            walls.append(wall)
            spans.append(st["span_ms"])
            plains += plain
            rejuvs += rejuv
            st = {k: v for k, v in st.items() if k != "span_ms"}
            stats_sum = st if stats_sum is None else {k: stats_sum[k] + st[k] for k in st}
        out = {"theta_sha": hashlib.sha256(np.ascontiguousarray(θ).tobytes()).hexdigest()[:16], "logZ_sum": float(np.sum(s.logZ)),
               "final_ess": float(s.ess), "posterior_mean": [float(v) for v in smc.expected_parameters(s).ravel()],
               "stages": len(getattr(s, "schedule", [])), "final_N": int(s.N)}
        s.close()
    n = len(walls)
    pu_local = stats_sum["particle_updates"] / n
    dev = prof                      # CUDA-event breakdown of the last warm-up run (same seed, same work)
    wall = float(np.mean(walls))      # the bench contract times exactly `steps` runs: the mean; median, min and every run ride along
    out.update({
        "workload": cfg["name"] + f", {resampler} inner resampling, θ sharded over {world} GPU(s)", "config": name, "algo": cfg["algo"],
        "N": cfg["N"], "M": cfg["M"], "T": cfg["T"], "scaling": "strong", "runs_timed": n, "wall_s": wall, "wall_s_median": float(np.median(walls)), "wall_s_min": float(np.min(walls)), "walls_s": [float(w) for w in walls],
        "device_spans_s": [1e-3 * float(v) for v in spans], "device_span_s": 1e-3 * float(np.mean(spans)),
        "particle_updates_local": pu_local, "bytes_per_update": cfg["bytes_per_update"],
        "s_per_plain_step": float(np.mean(plains)) if plains else None, "s_per_rejuvenation_step": float(np.mean(rejuvs)) if rejuvs else None,
        "rejuvenations": stats_sum["rejuvenations"] // n, "sweeps": stats_sum["sweeps"] // n, "clouds_received_this_rank": stats_sum["clouds_moved"] // n,
        "stream_syncs": stats_sum["syncs"] // n, "kernel_launches": stats_sum["launches"] // n,
        "breakdown_ms": dict(dev, host_control_and_gaps=1e3 * prof_wall - sum(dev.values()), wall_of_the_profiled_run=1e3 * prof_wall),
    })
    return out


def finish(out, world, reduce_max, reduce_sum):
    """max over ranks of the times, sum over ranks of the work; whole-job particle-updates/s"""
    out["wall_s"] = reduce_max(out["wall_s"])
    out["device_span_s"] = reduce_max(out["device_span_s"])
    out["particle_updates"] = reduce_sum(out.pop("particle_updates_local"))
    out["particle_updates_per_s"] = out["particle_updates"] / out["wall_s"]
    for k in ("s_per_plain_step", "s_per_rejuvenation_step"):
        if out[k] is not None:
            out[k] = reduce_max(out[k])
    out["breakdown_ms"] = {k: reduce_max(v) for k, v in out["breakdown_ms"].items()}
    out["clouds_received_all_ranks"] = int(reduce_sum(float(out.pop("clouds_received_this_rank"))))
    return out
