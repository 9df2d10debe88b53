"""config 3 / 4 / 5 on one GPU through the device sampler, static against dynamic scheduling of the batched engine:
python tools/c3_probe.py c3 [runs]   (SMCB_BATCH_CHUNK=0 forces one CTA per θ)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequential_monte_carlo_b200 as smc
from sequential_monte_carlo_b200 import bench_smc2 as B
ctx = smc.Context(0, 1998)
name = sys.argv[1]
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
out = B.run_config(name, ctx, None, 0, 1, steps=runs, warmup=2)
print(json.dumps({"config": name, "chunk_env": os.environ.get("SMCB_BATCH_CHUNK", "auto"), **{k: out[k] for k in ("wall_s", "wall_s_min", "device_span_s", "s_per_plain_step", "s_per_rejuvenation_step", "theta_sha", "breakdown_ms", "sweeps", "rejuvenations")}}))
