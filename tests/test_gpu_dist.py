"""Sharded loss head over NCCL on real GPUs: two ranks where the box has them, and a one-rank group on any box --
degenerate as an exchange, but every per-rank CUDA phase (partial forward, statistics merge, partial backward,
gradient return, normalisation backward, the one-call step and its scale launch) runs for real and is held to the
single-process oracle.  The host choreography for world_size 2 is covered on CPU by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_parity():
    n = min(torch.cuda.device_count(), 2)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tools", "dist_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all ranks OK" in r.stdout


@pytest.mark.parametrize("exchange", ["nccl", "symm"])
def test_one_rank_group_runs_every_sharded_phase(exchange):
    env = dict(os.environ, CE_DIST_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1",
                        "--master-addr", "127.0.0.1", "--master-port", "29618" if exchange == "nccl" else "29619",
                        os.path.join(ROOT, "tools", "dist_check.py")],
                       capture_output=True, text=True, timeout=420, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all ranks OK" in r.stdout and "one-call step: OK" in r.stdout
