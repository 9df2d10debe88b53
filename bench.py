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

N = 1 GPU: the workload above (it fits one GPU; a single filter does not shard: global scan + gather every step,
"replicas only").  The θ-level configurations (BASELINE configs[2..4]) ride along in the "smc2" object,
the rows built after the hot path (guided filter, matrix Kalman, per-θ moments: SURVEY §8f) in "widen".

N > 1 GPUs: the path that shards — BASELINE configs[4], smc² on the 4-parameter UCSV model with 4096 θ ×
4096 state particles, T = 241, θ sharded over the ranks through the library's own NCCL communicator (smcb_comm_init):
STRONG scaling, one "step" = one whole smc² run, value = particle-updates of the whole job / device time (max over
ranks), with the breakdown of where the time went, the same workload on ONE GPU of the same box beside it
("single_gpu_same_workload"), configs[2] and configs[3] sharded the same way ("smc2"), and the replicated
single-filter headline as an extra ("replicas").
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
UCSV_TRUE = [0.2, 0.2, 3.0, 1.0, 1.0]          # γε, γη, x0, log σε, log ση (docs/SPEC.md §4)
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

    def __init__(self, index, enabled=True):
        """index: the GPU to watch, or None for every GPU of the box (ONE poller for a multi-rank job: a poller per rank at 100 ms
        was measured to stall the ranks' launches and synchronisations — config 5 on 4 GPUs: single runs of 1.25-1.6 s among 0.75 s
        ones, profiles/r2_c5_runs_n4_nvidia_smi_polling.jsonl); enabled=False: this rank does not sample."""
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def start(self):
        if not self.enabled:
            return
        try:
            sel = ["-i", str(self.index)] if self.index is not None else []
            self.proc = subprocess.Popen(["nvidia-smi"] + sel + [f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.enabled:
            return None
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
KERNEL_BYTES = {"scan": 16, "bounds": 0, "anc": 12, "prop": 28}   # SURVEY §8d's algorithmic split of the 56 B
# what the LG1D launches really move: the log-weight is two fma of x', so move_kernel does not store it (−8 B) and sum_kernel
# reads x' instead of logw (same 8 B): 48 B per particle-update
KERNEL_BYTES_MOVED_LG1D = {"scan": 16, "bounds": 0, "anc": 12, "prop": 20}
KERNEL_NAMES = {"scan": "sum_kernel", "bounds": "bounds_kernel", "anc": "anc_hist_kernel", "prop": "move_kernel"}


def cpu_sample(steps, logn_sample=20, T_sample=64, style="reference"):
    """The CPU arm, single-threaded like the reference's particle filter (particles.jl:122).
    style="reference": the reference-STYLE port BASELINE.md §3 specifies (oracle/smc_oracle.c: smco_reference_style_log_likelihood —
    alias-table multinomial resampling rebuilt on every step, fresh allocations per step, per-particle sqrt / log σ, xoshiro256++);
    style="det": the parity oracle's log_likelihood (deterministic math + Philox, systematic resampling like the GPU arm)."""
    from oracle import oracle as o
    o.build()
    N = 1 << logn_sample
    _, y = o.simulate(0, LG_PARAMS, T_sample, DATA_SEED)
    times = []
    for s in range(steps):
        t0 = time.perf_counter()
        if style == "reference":
            o.reference_style_log_likelihood(LG_PARAMS, N, y, DATA_SEED + s)
        else:
            o.log_likelihood(0, LG_PARAMS, N, y, o.SYSTEMATIC, seed=DATA_SEED, epoch=s)
        times.append(time.perf_counter() - t0)
    what = ("reference-style port (alias-table multinomial resampling and fresh allocations every step, as particles.jl:107-129 executes)"
            if style == "reference" else "parity oracle (deterministic math + Philox, systematic resampling)")
    return N * T_sample / statistics.mean(times), f"LG1D N=2^{logn_sample}, T={T_sample}, {what}, oracle/smc_oracle.c (gcc -O2), linear in N*T"


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
        o.reference_style_batch(0, Pm, N, y, 1998)
        t1 = time.perf_counter() - t0
        if t1 > seconds / 4 or M >= 4096:
            break
        M *= 4
    t0 = time.perf_counter()
    o.batch_log_likelihood(0, Pm, None, N, y, o.MULTINOMIAL, 1998, 0, 0, want_state=False)
    t2 = time.perf_counter() - t0
    return {"value": M * N * T / t1, "unit": "particle-updates/s", "cores": threads, "kind": "port",
            "sample": f"one rejuvenation sweep: {M} theta x {N} particles x T={T}, reference-style filters (alias-table multinomial, allocations per step), "
                      "oracle/smc_oracle.c OpenMP over theta",
            "parity_oracle_value": M * N * T / t2}


def cpu_sharded_sample(seconds=10.0):
    """The CPU arm of the N > 1 workload (configs[4]): the reference spends >95 % of smc² in rejuvenate!'s `Threads.@threads for m`
    loop of full particle filters (smc_samplers.jl:112-121) — one such sweep of UCSV filters with 4096 particles each over
    T = 60 observations by the CPU oracle, OpenMP over θ with every host thread, multinomial like the reference."""
    from oracle import oracle as o
    o.build()
    threads = o.num_threads()
    P = o.params8([0.2, 0.2, 3.0, 1.0, 1.0])
    N, T = 4096, 60
    _, y = o.simulate(2, P, T, DATA_SEED)
    M, t1 = threads, 0.0
    while True:
        Pm = np.tile(P[None, :], (M, 1))
        t0 = time.perf_counter()
        o.reference_style_batch(2, Pm, N, y, DATA_SEED)      # the reference-style port, as at N = 1
        t1 = time.perf_counter() - t0
        if t1 > seconds / 4 or M >= 4096:
            break
        M *= 4
    return M * N * T / t1, threads, f"one rejuvenation sweep of configs[4]'s inner filters: {M} theta x {N} particles x T={T}, UCSV, reference-style filters " \
                                    f"(alias-table multinomial, allocations per step), oracle/smc_oracle.c OpenMP over theta ({threads} threads)"


def run_reference(args, rank):
    if rank != 0:
        return
    note = ("Julia is not installed in this image and the reference module does not load as shipped (SURVEY F2/F7): the "
            "reference arm is the C port of its algorithm (oracle/smc_oracle.c)")
    if args.gpus > 1:
        # the arm of the sharded workload: every host thread, like Threads.@threads over θ
        vals = [cpu_sharded_sample() for _ in range(max(1, min(args.steps, 2)))]
        v, threads, sample = float(np.mean([a[0] for a in vals])), vals[0][1], vals[0][2]
        from sequential_monte_carlo_b200 import bench_smc2
        line = {
            "impl": "reference", "metric": "particle-updates/sec (N×T) bootstrap PF", "value": v, "unit": "particle-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": sharded_config(args.gpus),
            "extrapolated": True, "measured_on": sample,
            "cpu_baseline": {"value": v, "unit": "particle-updates/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": note + "; the value is the throughput of the sampled sweep, the whole 4096 θ × 4096 × T=241 run was not executed on the CPU",
        }
        print(json.dumps(line), flush=True)
        return
    sample_logn, sample_T = 20, 32
    v, sample = cpu_sample(max(args.steps, 1), sample_logn, sample_T)
    v_det, sample_det = cpu_sample(1, sample_logn, sample_T, style="det")
    N, T = 1 << args.logn, args.T
    line = {
        "impl": "reference", "metric": "particle-updates/sec (N×T) bootstrap PF", "value": v, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (1 << sample_logn) * sample_T / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(N, T),
        "extrapolated": True, "sample_N": 1 << sample_logn, "sample_T": sample_T,
        "ms_per_step_note": f"ms_per_step is the measured time of one sampled sweep (N=2^{sample_logn}, T={sample_T}); the value (particle-updates/s) "
                            "is linear in N·T for a per-particle loop, the full N=2^24 × T=1000 sweep (≈ 2000 s on one core) was not run",
        "cpu_baseline": {"value": v, "unit": "particle-updates/s", "cores": 1, "kind": "port", "sample": sample},
        "cpu_baseline_parity_oracle": {"value": v_det, "unit": "particle-updates/s", "cores": 1, "kind": "port", "sample": sample_det},
        "e2e": {"value": v, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": note + "; the reference PF is single-threaded (particles.jl:122), so one core is all it can use; the arm is the "
                       "reference-style port of BASELINE.md §3 (multinomial by alias table, allocations per step), the parity oracle's "
                       "own speed rides along",
    }
    print(json.dumps(line), flush=True)


def sharded_config(world):
    return {"workload": "smc² + smc²! t=2..241 on the 4-parameter UCSV model (examples/inflation_example.jl shape), 4096 θ × 4096 state "
                        f"particles, chain 3, ESS 0.5 (BASELINE.json configs[4]), θ sharded over {world} GPUs",
            "N": 4096, "M": 4096, "T": 241, "chain": 3, "resampler": "systematic", "model": None,
            "l2": "state clouds 4096 θ × 4096 × 4 × 8 B = 0.5 GB per copy (larger than L2 at 1-2 GPUs); every sweep redraws all clouds",
            "parallelism": f"theta-sharded x{world} (ncclAllGather of M/G log-likelihoods per step, ncclSend/Recv of resampled clouds)"}


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

    if world > 1:
        main_sharded(args, rank, world, local, barrier)
        dist.destroy_process_group()
        return

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
                     "per_kernel": {KERNEL_NAMES[k]: {"bytes_per_update": KERNEL_BYTES[k], "bytes_moved_per_update": KERNEL_BYTES_MOVED_LG1D[k],
                                                      "us": step_us[k],
                                                      "GB/s": (KERNEL_BYTES_MOVED_LG1D[k] * N / (step_us[k] * 1e-6) / 1e9) if step_us[k] else None}
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
        # the binary32-ARITHMETIC tier (docs/SPEC.md §9b: float normals four per Philox block, float4 loads / stores): same workload
        try:
            ctx.set_precision("f32_arith")
            ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
            msa, za = 0.0, 0.0
            for _ in range(2):
                za = ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
                msa += ctx.timing()[0]["total"]
            ctx.set_profiling(True)
            ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y, rs, stream=rank)
            ms_, n_ = ctx.timing()
            ctx.set_profiling(False)
            va = 2 * N * T / (msa * 1e-3)
            line["f32_arithmetic"] = {"value": va, "unit": "particle-updates/s", "ms_per_step": msa / 2, "us_per_filter_step": 1e3 * msa / 2 / T,
                                      "dtype": "f32 normals / model arithmetic / log-weights, u64 CDF", "algorithmic_bytes_per_update": 40,
                                      "roofline_frac": va * 40 / (peak * 1e9), "logZ": za,
                                      "avg_us_per_launch": {KERNEL_NAMES[k]: 1e3 * ms_[k] / max(n_[k], 1) for k in KERNEL_NAMES}}
        except Exception as e:
            line["f32_arithmetic"] = {"error": repr(e)}
        finally:
            ctx.set_precision("f64")
    if rank == 0 and world == 1 and not args.no_cpu:
        v, sample = cpu_sample(2, 20, 64)
        line["cpu_baseline"] = {"value": v, "unit": "particle-updates/s", "cores": 1, "kind": "port", "sample": sample}
        vd, sd_ = cpu_sample(1, 20, 32, style="det")
        line["cpu_baseline_parity_oracle"] = {"value": vd, "unit": "particle-updates/s", "cores": 1, "kind": "port", "sample": sd_}
    # the reference's own resampling law (multinomial: particles.jl:17-19, the API default) on the same workload: two-level
    # draw of docs/SPEC.md §5c (sum -> mn_prep -> mn_count -> mn_cell -> move), against the same 56 B roofline
    try:
        Tm = min(T, 100)
        ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y[:Tm], smc.MULTINOMIAL, stream=rank)
        msm = 0.0
        ctx.set_profiling(True)
        km = {}
        for _ in range(2):
            ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y[:Tm], smc.MULTINOMIAL, stream=rank)
            ms_, n_ = ctx.timing()
            for k in ("scan", "bounds", "anc", "prop"):
                km[k] = km.get(k, 0.0) + 1e3 * ms_[k] / max(n_[k], 1) / 2
        ctx.set_profiling(False)
        for _ in range(2):
            ctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, N, y[:Tm], smc.MULTINOMIAL, stream=rank)
            msm += ctx.timing()[0]["total"]
        vm = 2 * N * Tm / (msm * 1e-3)
        line["multinomial"] = {"workload": f"the headline workload with multinomial resampling (the reference's law and the API default), T={Tm}",
                               "value": vm, "unit": "particle-updates/s", "us_per_step": 1e3 * msm / 2 / Tm, "roofline_frac_56B": vm * BYTES_PER_UPDATE / (peak * 1e9),
                               "vs_systematic": vm / value,
                               "avg_us_per_launch": {"sum_kernel": km.get("scan"), "mn_prep+mn_count": km.get("bounds"), "mn_cell_kernel": km.get("anc"),
                                                     "move_kernel": km.get("prop")}}
        if not args.no_cpu:
            from oracle import oracle as o
            t0 = time.perf_counter()
            o.reference_style_log_likelihood(LG_PARAMS, 1 << 20, y[:16], DATA_SEED)
            line["multinomial"]["cpu_baseline"] = {"value": (1 << 20) * 16 / (time.perf_counter() - t0), "unit": "particle-updates/s", "cores": 1, "kind": "port",
                                                   "sample": "LG1D N=2^20, T=16, multinomial by alias table rebuilt every step (the reference-style port of BASELINE.md §3), oracle/smc_oracle.c"}
            t0 = time.perf_counter()
            o.log_likelihood(0, LG_PARAMS, 1 << 20, y[:16], o.MULTINOMIAL, seed=DATA_SEED, epoch=0)
            line["multinomial"]["cpu_baseline_parity_oracle"] = {"value": (1 << 20) * 16 / (time.perf_counter() - t0), "unit": "particle-updates/s", "cores": 1, "kind": "port",
                                                                 "sample": "LG1D N=2^20, T=16, multinomial (SPEC §5c two-level draw), oracle/smc_oracle.c"}
    except Exception as e:
        line["multinomial"] = {"error": repr(e)}
    if not args.no_smc2:
        try:
            line["smc2"] = smc2_legs(ctx, None, 0, 1, None, ("c3", "c3_multinomial"), steps=5, warmup=2)   # 16 ms runs: cheap to repeat
            line["smc2"].update(smc2_legs(ctx, None, 0, 1, None, ("c4",), steps=3, warmup=1))
            line["smc2"].update(smc2_legs(ctx, None, 0, 1, None, ("c5", "c5_multinomial"), steps=1, warmup=1))
            if not args.no_cpu:
                line["smc2"]["cpu_baseline"] = cpu_rejuvenation_sample(1024, 100)
        except Exception as e:  # the headline line must still print
            line["smc2"] = {"error": repr(e)}
    try:
        line["config1_latency"] = config1_latency(ctx)
    except Exception as e:
        line["config1_latency"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_widen:
        try:
            line["widen"] = widen_leg(ctx)
        except Exception as e:  # the headline line must still print
            line["widen"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _reducers(world):
    if world == 1:
        return (lambda v: float(v)), (lambda v: float(v))
    import torch
    import torch.distributed as dist

    def red(op):
        def f(v):
            t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=op)
            return float(t.item())
        return f
    return red(dist.ReduceOp.MAX), red(dist.ReduceOp.SUM)


def smc2_legs(ctx, comm, rank, world, barrier, names, steps=1, warmup=1):
    """the θ-level configurations through the public sampler API on the device-resident engine"""
    from sequential_monte_carlo_b200 import bench_smc2
    rmax, rsum = _reducers(world)
    out = {}
    for name in names:
        try:
            out[name] = bench_smc2.finish(bench_smc2.run_config(name, ctx, comm, rank, world, steps=steps, warmup=warmup, barrier=barrier),
                                          world, rmax, rsum)
        except Exception as e:
            out[name] = {"error": repr(e)}
    return out


def config1_latency(ctx, reps=20):
    """BASELINE configs[0]: log_likelihood(1024, y, lg_mod([0.5,0.9,0.8])), T = 100 — wall clock of ONE public call with host
    buffers (H2D of y, the whole series in one launch on the batched engine, D2H of x, w, logZ), checked against kalman_filter."""
    import sequential_monte_carlo_b200 as smc
    from sequential_monte_carlo_b200 import particles, state_space_models as ssm
    model = ssm.LinearGaussian(LG_THETA[0], 1.0, LG_THETA[1], LG_THETA[2], 0.0)
    y = smc._lib.simulate(smc.KIND_LG1D, LG_PARAMS, 100, DATA_SEED)[1]
    particles.log_likelihood(1024, y, model)
    ts, zs = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        x, w, z = particles.log_likelihood(1024, y, model)
        _ = float(z) + float(np.asarray(w)[0])
        ts.append(time.perf_counter() - t0)
        zs.append(float(z))
    kal = float(ctx.kalman_loglik(LG_PARAMS, y, matched_init=True)[0][0])
    return {"workload": "log_likelihood(1024, y, lg_mod([0.5,0.9,0.8])), T=100, multinomial (the API default), one public call incl. H2D/D2H "
                        "(BASELINE.json configs[0])", "wall_ms_median": 1e3 * statistics.median(ts), "wall_ms_min": 1e3 * min(ts),
            "particle_updates_per_s": 1024 * 100 / statistics.median(ts), "logZ_mean": float(np.mean(zs)), "logZ_sd": float(np.std(zs)),
            "kalman_matched_init_logZ": kal}


def main_sharded(args, rank, world, local, barrier):
    """N > 1: BASELINE configs[4] θ-sharded over the ranks (strong scaling), see the module docstring"""
    import torch
    import torch.distributed as dist
    import sequential_monte_carlo_b200 as smc
    from sequential_monte_carlo_b200 import bench_smc2, smc_samplers as ss
    ctx = smc.Context(local, seed=DATA_SEED)
    comm = ss.NcclComm.from_torch(ctx)
    rmax, rsum = _reducers(world)
    sampler = ClockSampler(None, enabled=(rank == 0))   # one poller for the whole box, on rank 0
    barrier()
    sampler.start()
    head = bench_smc2.run_config("c5", ctx, comm, rank, world, steps=args.steps, warmup=max(1, min(args.warmup, 3)), barrier=barrier)
    barrier()
    clocks = sampler.stop()
    head = bench_smc2.finish(head, world, rmax, rsum)
    peak, peak_src = measured_peak()
    units = head["particle_updates"]
    value = units / head["device_span_s"]
    filt_s = 1e-3 * head["breakdown_ms"]["filter_ms"]
    line = {
        "metric": "particle-updates/sec (N×T) bootstrap PF", "value": value, "unit": "particle-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": 1e3 * head["device_span_s"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic (simulate(), Philox seed 1998)",
        "config": sharded_config(world),
        "e2e": {"value": units / head["wall_s"], "unit": "particle-updates/s", "h2d_bytes_per_step": 8 * 241,
                "d2h_bytes_per_step": 64 * (241 + 3 * head["rejuvenations"]) + 8 * 4096 * 6, "ms_per_step": 1e3 * head["wall_s"],
                "note": "wall clock of smc²(smc, y) + smc²!(smc, y, t) for t = 2..T through the public sampler API: y from host memory, ess read back "
                        "every step, θ / ω / logZ read back at the end; the sampler object is created outside"},
        "gpu_launches": int(head["kernel_launches"] * args.steps),
        "breakdown_ms_per_step": head["breakdown_ms"],
        "run": {k: head[k] for k in ("theta_sha", "logZ_sum", "final_ess", "posterior_mean", "rejuvenations", "sweeps", "clouds_received_all_ranks",
                                     "s_per_plain_step", "s_per_rejuvenation_step", "stream_syncs", "wall_s", "wall_s_median", "walls_s", "device_span_s", "device_spans_s", "particle_updates")},
        "roofline": {"bound": "hbm", "achieved": 88 * (units / world) / filt_s / 1e9 if filt_s > 0 else None, "peak": peak, "unit": "GB/s",
                     "frac": (88 * (units / world) / filt_s / 1e9 / peak) if filt_s > 0 else None, "traffic": None, "peak_source": peak_src,
                     "kernel": "batch_kernel<ModelUCSV> (one CTA per θ, whole series in one launch): 88 algorithmic B per particle-update "
                               "(SURVEY §8d, UCSV fp64) × this rank's particle-updates ÷ its device time in the inner filters; the clouds are "
                               "block-resident, so the kernel is bound by fp64 issue / barrier latency, not HBM (SURVEY §8d, DESIGN.md §4)"},
        "clocks": clocks,
    }
    # the same workload on ONE GPU of this box, same build, same run (rank 0 alone; the other ranks wait at the barrier)
    if not args.no_smc2:
        if rank == 0:
            try:
                ctx1 = smc.Context(local, seed=DATA_SEED)
                one = bench_smc2.finish(bench_smc2.run_config("c5", ctx1, None, 0, 1, steps=1, warmup=1), 1, float, float)
                line["single_gpu_same_workload"] = {k: one[k] for k in ("wall_s", "device_span_s", "particle_updates", "particle_updates_per_s", "theta_sha",
                                                                          "breakdown_ms", "rejuvenations", "s_per_plain_step", "s_per_rejuvenation_step")}
                line["single_gpu_same_workload"]["value"] = one["particle_updates"] / one["device_span_s"]
                ctx1.close()
            except Exception as e:
                line["single_gpu_same_workload"] = {"error": repr(e)}
        barrier()
        line["smc2"] = smc2_legs(ctx, comm, rank, world, barrier, ("c3", "c4", "c5_multinomial"), steps=2, warmup=1)
    # the single-filter headline as N independent replicas (weak scaling, no collective): a single filter does not shard
    try:
        Nr, Tr = 1 << args.logn, 200
        yr = smc._lib.simulate(smc.KIND_LG1D, LG_PARAMS, Tr, DATA_SEED)[1]
        rctx = smc.Context(local, seed=DATA_SEED + rank)
        rctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, Nr, yr, smc.SYSTEMATIC, stream=rank)
        barrier()
        ms = 0.0
        for _ in range(2):
            rctx.log_likelihood(smc.KIND_LG1D, LG_PARAMS, Nr, yr, smc.SYSTEMATIC, stream=rank)
            ms += rctx.timing()[0]["total"]
        barrier()
        rctx.close()
        line["replicas"] = {"workload": f"log_likelihood LG1D N=2^{args.logn}, T={Tr}, systematic: one independent filter per GPU (weak scaling)",
                            "value": 2 * Nr * Tr * world / (rmax(ms) * 1e-3), "unit": "particle-updates/s", "ms_per_sweep": rmax(ms) / 2}
    except Exception as e:
        line["replicas"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


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
    # the guided move of UCSV (docs/SPEC.md §10b: tempered optimal trend proposal) on config 5's inner shape: 296 independent
    # filters of 4096 particles at the true θ, κ = 1 against the bootstrap filter — device time and the scatter of logZ
    Mu, Nu, Tu = 296, 4096, 100
    Pu = np.tile(smc._lib.params8(UCSV_TRUE), (Mu, 1))
    yu = smc._lib.simulate(smc.KIND_UCSV, UCSV_TRUE, Tu, 1998)[1]
    pu = np.zeros((Tu, Mu, 3))
    pu[:, :, 0], pu[:, :, 2] = 1.0, 1.0
    bu = ctx.batch(smc.KIND_UCSV, Mu, Nu)
    gu = {"workload": f"{Mu} θ × {Nu} particles, T={Tu}, UCSV at the true θ, systematic; κ = 1 (locally optimal trend move) against bootstrap"}
    for name, q in (("bootstrap", None), ("guided", pu)):
        bu.log_likelihood(Pu, yu, smc.SYSTEMATIC, 0, proposal=q)
        z = bu.log_likelihood(Pu, yu, smc.SYSTEMATIC, 0, proposal=q)
        ms = bu.timing()[0]
        gu[name] = {"ms_per_sweep": ms, "particle_updates_per_s": Mu * Nu * Tu / (ms * 1e-3), "sd_logZ_over_theta": float(np.std(z)), "mean_logZ": float(np.mean(z))}
    bu.close()
    out["guided_ucsv"] = gu
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
