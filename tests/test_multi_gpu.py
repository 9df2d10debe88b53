"""The θ-sharded samplers on 2 NCCL ranks against the same run on one rank (results do not depend on the number
of GPUs) — real NCCL, real GPUs, through smcb_comm_init / smcb_sampler_*.  Skipped on a box with one GPU; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import numpy as np
import pytest

import sequential_monte_carlo_b200 as smc
from sequential_monte_carlo_b200 import smc_samplers as ss

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("which", ["lg", "ucsv", "sv_dt", "lg_dyn"])
def test_two_nccl_ranks_equal_one_rank(ctx, which, tmp_path):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.dist_worker import build_sampler, run
    s, y, mode = build_sampler(smc, which, ctx, None)
    rejuv = run(smc, s, y, mode)
    θ, ω, z, x, w, means = s.θ, s.ω, s.logZ, s.x, s.w, ss.state_means(s)
    s.close()
    assert len(rejuv) >= 2
    out = str(tmp_path / "dist")
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
           "29611", os.path.join(ROOT, "tests", "dist_worker.py"), which, out]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    M = θ.shape[0]
    moved = 0
    for rank in range(2):
        d = np.load(f"{out}.rank{rank}.npz")
        np.testing.assert_array_equal(d["theta"], θ)            # replicated vectors: bit-identical on every rank
        np.testing.assert_array_equal(d["logZ"], z)
        np.testing.assert_array_equal(d["omega"], ω)
        np.testing.assert_array_equal(d["rejuv"], np.array(rejuv, np.float64))
        lo, hi = rank * M // 2, (rank + 1) * M // 2
        np.testing.assert_array_equal(d["x"], x[lo:hi])         # this rank's slice of the clouds
        np.testing.assert_array_equal(d["w"], w[lo:hi])
        np.testing.assert_array_equal(d["means"], means)        # all-gathered per-θ state means
        moved += int(d["moved"])
    assert moved > 0                                            # some clouds did cross the GPUs
