#!/usr/bin/env python
"""bench.py — headline benchmark of the particle-filter hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--logn 24] [--T 1000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `log_likelihood(N, y, model)` sweep (bootstrap PF, resample every step) of the
univariate LinearGaussian lg_mod([0.5,0.9,0.8]) over T synthetic observations — BASELINE.json
configs[1] at its largest size (N = 2^24, T = 1000) — through the C ABI of libsmcb200.so.

  value  particle-updates/s (N*T per sweep), device time of the sweeps (CUDA events on the library's
         stream; the only inputs, y[T] and 6 parameters, travel as kernel arguments)
  e2e    the same metric through the public Python API `log_likelihood(N, y, model) -> (x, w, logZ)`
         with host buffers: wall clock including the D2H copy of the final cloud and weights
  roofline  one filter step (sum + bounds + ancestor + move kernels) against the measured HBM peak using
         SURVEY.md §8(d)'s 56 algorithmic bytes per particle-update; the per-kernel split (algorithmic
         bytes, µs, GB/s of each of the four launches) rides along in roofline.per_kernel
  cpu_baseline  the CPU oracle (a port of the Julia reference; Julia is not in this image) timed on
         the host cores on a bounded sample

N > 1 GPUs: a single filter does not shard (global scan + gather every step: "replicas only",
DESIGN.md §5) — every rank runs its own filter on its own Philox stream (weak scaling).  The
θ-sharded SMC² numbers (the path that does shard) ride along in the "smc2" object, and the rows built after
the hot path (guided filter, matrix Kalman, per-θ moments: SURVEY §8f) in the "widen" object.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LG_THETA = [0.5, 0.9, 0.8]                       # README.md:21  (A, Q, R)
LG_PARAMS = [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]       # A, B, Q, R, x0, σ0
BYTES_PER_UPDATE = 56
DATA_SEED = 1998


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# algorithmic bytes per particle-update of each launch of one step (LG1D fp64): they add up to the 56 B of SURVEY §8(d)
KERNEL_BYTES = {"scan": 16, "bounds": 0, "anc": 12, "prop": 28}
KERNEL_NAMES = {"scan": "sum_kernel", "bounds": "bounds_kernel", "anc": "anc_hist_kernel", "prop": "move_kernel"}


def cpu_sample(steps, logn_sample=20, T_sample=64):
    """The oracle's log_likelihood (port of particles.jl:132-147; single-threaded like the reference)."""
    from oracle import oracle as o
    o.build()
    N = 1 << logn_sample
    _, y = o.simulate(0, LG_PARAMS, T_sample, DATA_SEED)
    times = []
    for s in range(steps):
        t0 = time.perf_counter()
        o.log_likelihood(0, LG_PARAMS, N, y, o.SYSTEMATIC, seed=DATA_SEED, epoch=s)
        times.append(time.perf_counter() - t0)
    return N * T_sample / statistics.mean(times), f"LG1D N=2^{logn_sample}, T={T_sample}, systematic, oracle/smc_oracle.c (gcc -O2), linear in N*T"


def cpu_rejuvenation_sample(N=1024, T=100, seconds=8.0):
    """The reference's dominant cost beside it (SURVEY §8d ii): one `Threads.@threads for m` sweep of full particle
    filters (rejuvenate!, smc_samplers.jl:112-121) by the CPU oracle, OpenMP over θ with every host thread, on a
    bounded number of θ-particles.  bench.py is one of the few places allowed to execute oracle/ (as the baseline)."""
    from oracle import oracle as o
    o.build()
    threads = o.num_threads()
    P = np.tile(o.params8([0.5, 1.0, 0.9, 0.8, 0.0, 1.0]), (threads, 1))
    _, y = o.simulate(0, P[0], T, 1998)
    M, t1 = threads, 0.0
    while True:                                    # grow the sample until it runs for about `seconds`
        Pm = np.tile(P[:1], (M, 1))
        t0 = time.perf_counter()
        o.batch_log_likelihood(0, Pm, None, N, y, o.MULTINOMIAL, 1998, 0, 0, want_state=False)
        t1 = time.perf_counter() - t0
        if t1 > seconds / 4 or M >= 4096:
            break
        M *= 4
    return {"value": M * N * T / t1, "unit": "particle-updates/s", "cores": threads, "kind": "port",
            "sample": f"one rejuvenation sweep: {M} theta x {N} particles x T={T}, multinomial, oracle/smc_oracle.c OpenMP over theta"}


def run_reference(args, rank):
    if rank != 0:
        return
    v, sample = cpu_sample(max(args.steps, 1) + 0, 20, 32)
    N, T = 1 << args.logn, args.T
    line = {
        "impl": "reference", "metric": "particle-updates/sec (N×T) bootstrap PF", "value": v, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * N * T / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(N, T),
        "cpu_baseline": {"value": v, "unit": "particle-updates/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Julia is not installed in this image and the reference module does not load as shipped (SURVEY F2/F7): the "
                "reference arm is the C port of its algorithm; the reference PF is single-threaded (particles.jl:122)",
    }
    print(json.dumps(line), flush=True)


def workload_config(N, T):
    return {"workload": f"log_likelihood bootstrap PF, univariate LinearGaussian lg_mod([0.5,0.9,0.8]), N=2^{N.bit_length() - 1} "
                        f"particles, T={T}, systematic resampling every step (BASELINE.json configs[1])",
            "N": N, "T": T, "resampler": "systematic", "model": None, "l2": "working set (x ping-pong + logw, 0.4 GB) larger than L2",
            "parallelism": "replicas (one filter per GPU)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--logn", type=int, default=24)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-smc2", action="store_true", help="skip the θ-sharded SMC² leg")
    ap.add_argument("--no-widen", action="store_true", help="skip the guided-filter / matrix-Kalman leg (SURVEY §8f rows)")
    ap.add_argument("--no-f32", action="store_true", help="skip the binary32-state tier leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import sequential_monte_carlo_b200 as smc
    from sequential_monte_carlo_b200 import particles, state_space_models as ssm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsmcb200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, T = 1 << args.logn, args.T
    _, y = smc._lib.simulate(smc.KIND_LG1D, LG_PARAMS, T, DATA_SEED)
    model = ssm.LinearGaussian(LG_THETA[0], 1.0, LG_THETA[1], LG_THETA[2], 0.0)
    particles.set_default_context(smc.Context(local, seed=DATA_SEED + rank))
    ctx = particles.default_context()
    rs = smc.SYSTEMATIC

    # ---- the two arms alternate sweep by sweep inside ONE bracketed region, so that both see the same clocks (a
    # 1 kW part drifts under its power cap over a multi-second run):
    #   kernel arm  device-resident state, CUDA events on the library's stream around the sweep
    #   e2e arm     the public API with host buffers: wall clock of log_likelihood + the D2H read of (x, w)
    for _ in range(args.warmup):
        ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
    xw = particles.log_likelihood(N, y[: min(T, 8)], model, resampler="systematic")   # page-locks the host buffers once
    xw[0].numpy(), xw[1].numpy()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    dev_ms, launches, wall_kernel, wall_e2e = 0.0, 0, 0.0, 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
        ms, n = ctx.timing()
        wall_kernel += time.perf_counter() - t0
        dev_ms += ms["total"]
        launches += n["total"]
        t0 = time.perf_counter()
        x, w, logZ = particles.log_likelihood(N, y, model, resampler="systematic")
        _ = float(logZ) + float(w[0]) + float(x[0])
        wall_e2e += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()

    # ---- per-kernel durations (CUDA events around every launch; same workload, K sweeps)
    ctx.set_profiling(True)
    kms = {"scan": 0.0, "prop": 0.0, "init": 0.0, "stats": 0.0, "bounds": 0.0, "anc": 0.0}
    kn = dict.fromkeys(kms, 0)
    for _ in range(max(1, min(args.steps, 2))):
        ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
        ms, n = ctx.timing()
        for k in kms:
            kms[k] += ms.get(k, 0.0)
            kn[k] += n.get(k, 0)
    ctx.set_profiling(False)

    tmax = torch.tensor([dev_ms, wall_kernel, wall_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_kernel_max, wall_e2e_max = (float(v) for v in tmax.tolist())

    units = N * T * args.steps * world
    value = units / (dev_ms_max * 1e-3)
    e2e = units / wall_e2e_max
    peak, peak_src = measured_peak()
    step_us = {k: (1e3 * kms[k] / kn[k] if kn[k] else None) for k in kms}
    # average duration of one step (its four launches) inside the timed sweeps: CUDA events around the whole sweep on
    # the library's stream, minus the two launches that are not part of a step (init at t=1, and the T-th sum_kernel
    # that only produces the statistics of the final weights).  The per-launch events below add ~2.5 us of event
    # overhead to every launch, so their sum is reported but not used.
    fused_events = sum(v for k, v in step_us.items() if k in ("scan", "bounds", "anc", "prop") and v)
    sweep_us = 1e3 * dev_ms_max / args.steps
    fused = (sweep_us - (step_us["init"] or 0.0) - (step_us["scan"] or 0.0)) / max(T - 1, 1) if T > 1 else None
    achieved = BYTES_PER_UPDATE * N / (fused * 1e-6) / 1e9 if fused else None

    line = {
        "metric": "particle-updates/sec (N×T) bootstrap PF", "value": value, "unit": "particle-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (simulate(), Philox seed 1998)",
        "config": workload_config(N, T),
        "e2e": {"value": e2e, "unit": "particle-updates/s", "h2d_bytes_per_step": int(8 * T + 64),
                "d2h_bytes_per_step": int(16 * N + 8), "ms_per_step": 1e3 * wall_e2e_max / args.steps},
        "gpu_launches": launches,
        "wall_ms_per_step_kernel_arm": 1e3 * wall_kernel_max / args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                     "traffic": load_traffic(), "peak_source": peak_src,
                     "kernel": "one filter step = sum_kernel + bounds_kernel + anc_hist_kernel + move_kernel "
                               "(16 + 0 + 12 + 28 = 56 algorithmic B/particle-update, SURVEY §8d)",
                     "step_us": fused, "step_us_sum_of_per_launch_events": fused_events,
                     "avg_us_per_launch": step_us,
                     "per_kernel": {KERNEL_NAMES[k]: {"bytes_per_update": KERNEL_BYTES[k], "us": step_us[k],
                                                      "GB/s": (KERNEL_BYTES[k] * N / (step_us[k] * 1e-6) / 1e9) if step_us[k] else None}
                                    for k in KERNEL_BYTES}},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_f32:
        # the binary32-state tier of the same workload (docs/SPEC.md §9): two sweeps, device time; SURVEY §8d counts
        # 40 algorithmic B per particle-update for it
        try:
            ctx.set_precision("f32")
            ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
            ms32 = 0.0
            for _ in range(2):
                ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
                ms32 += ctx.timing()[0]["total"]
            v32 = 2 * N * T / (ms32 * 1e-3)
            line["f32_states"] = {"value": v32, "unit": "particle-updates/s", "ms_per_step": ms32 / 2, "dtype": "f32 states, f64 arithmetic",
                                  "algorithmic_bytes_per_update": 40, "roofline_frac": v32 * 40 / (peak * 1e9)}
        except Exception as e:
            line["f32_states"] = {"error": repr(e)}
        finally:
            ctx.set_precision("f64")
    if rank == 0 and world == 1 and not args.no_cpu:
        v, sample = cpu_sample(2, 20, 64)
        line["cpu_baseline"] = {"value": v, "unit": "particle-updates/s", "cores": 1, "kind": "port", "sample": sample}
    if not args.no_smc2:
        try:
            from sequential_monte_carlo_b200 import bench_smc2
            line["smc2"] = bench_smc2.run(local, rank, world)
            if rank == 0 and world == 1 and not args.no_cpu:
                line["smc2"]["cpu_baseline"] = cpu_rejuvenation_sample(1024, 100)
        except Exception as e:  # the headline line must still print
            line["smc2"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_widen:
        try:
            line["widen"] = widen_leg(ctx)
        except Exception as e:  # the headline line must still print
            line["widen"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def widen_leg(ctx, M=512, N=1024, T=100):
    """SURVEY §8(f) rows built after the hot path (device time of the library's own events, ms): the guided filter
    (docs/SPEC.md §10) against the bootstrap filter on config 3's inner shape (512 θ × 1024 particles, T = 100, all θ
    at the true parameters so that the scatter of logZ over θ is the estimator's own), the matrix Kalman filter on 4096
    Hodrick–Prescott models, and the per-θ moments."""
    import sequential_monte_carlo_b200 as smc
    P = np.tile(smc._lib.params8(LG_PARAMS), (M, 1))
    y = smc._lib.simulate(smc.KIND_LG1D, LG_PARAMS, T, 1998)[1]
    lg = smc.LinearGaussian(*LG_PARAMS[:5])
    prop = np.array([[smc.locally_optimal_proposal(lg, yt)] * M for yt in y])
    b = ctx.batch(smc.KIND_LG1D, M, N)
    out = {"workload": f"{M} θ × {N} particles, T={T}, LG1D at the true θ, systematic; one launch per sweep"}
    for name, q in (("bootstrap", None), ("guided", prop)):
        b.log_likelihood(P, y, smc.SYSTEMATIC, 0, proposal=q)          # warm-up (module load)
        t0 = time.perf_counter()
        z = b.log_likelihood(P, y, smc.SYSTEMATIC, 0, proposal=q)
        wall = (time.perf_counter() - t0) * 1e3
        ms = b.timing()[0]
        out[name] = {"ms_per_sweep": ms, "wall_ms_per_sweep": wall, "particle_updates_per_s": M * N * T / (ms * 1e-3), "sd_logZ_over_theta": float(np.std(z)),
                     "mean_logZ": float(np.mean(z))}
    # one large-N guided filter (guided_move_kernel in place of move_kernel, log-weights stored: 56 B per particle-update)
    Ng, yg = 1 << 24, y[:50]
    ctx.guided_log_likelihood(smc.KIND_LG1D, LG_PARAMS, Ng, yg, prop[:50, 0], smc.SYSTEMATIC)
    zg = ctx.guided_log_likelihood(smc.KIND_LG1D, LG_PARAMS, Ng, yg, prop[:50, 0], smc.SYSTEMATIC)
    msg = ctx.timing()[0]["total"]
    out["guided_single_filter"] = {"workload": "LG1D N=2^24, T=50, systematic, locally optimal proposal", "ms_per_sweep": msg,
                                   "particle_updates_per_s": Ng * 50 / (msg * 1e-3), "roofline_frac_56B": Ng * 50 * 56 / (msg * 1e-3) / (measured_peak()[0] * 1e9),
                                   "logZ": zg}
    t0 = time.perf_counter()
    mean, var = b.weighted_moments()
    out["per_theta_moments_ms"] = (time.perf_counter() - t0) * 1e3
    b.close()
    out["kalman_matched_init_logZ"] = float(ctx.kalman_loglik(LG_PARAMS, y, matched_init=True)[0][0])
    yhp = smc._lib.simulate(smc.KIND_LG1D, [1.0, 1.0, 0.05, 1.0, 0.0, 1.0], 241, 1998)[1]
    blocks = np.stack([smc.hodrick_prescott(λ=lam, y=yhp).block() for lam in np.geomspace(1.0, 1e5, 4096)])
    ctx.kalman_mv_loglik(2, blocks, yhp)
    t0 = time.perf_counter()
    ll, _, _ = ctx.kalman_mv_loglik(2, blocks, yhp)
    out["kalman_mv"] = {"workload": "4096 Hodrick–Prescott models (d = 2), T = 241, one launch, wall incl. H2D/D2H",
                        "ms": (time.perf_counter() - t0) * 1e3, "best_lambda": float(np.geomspace(1.0, 1e5, 4096)[int(np.argmax(ll))])}
    return out


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the four launches of one step at N=2^24, from the committed
    `ncu --set full` capture (profiles/traffic.json names the report), or null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_step"]
    except Exception:
        return None


if __name__ == "__main__":
    main()
