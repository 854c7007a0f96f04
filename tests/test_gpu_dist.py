"""Multi-GPU parity under pytest: when the box has >= 2 GPUs, two ranks (NCCL) run the node-range sharded
MagNetConv (push exchange on a permuted DSBM graph, halo all-to-all on a locality-ordered graph) and compare
their rows with the single-GPU layer (tools/dist_check.py: <= 2e-6 relative, K = 1 and 2, one shared operand,
repeated calls).  Skipped on single-GPU boxes; bench.py --gpus N additionally carries an in-line parity_check
against the oracle in its JSON line."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["2", "1", "0"])
def test_sharded_layer_matches_single_gpu_world2(engine):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, PGSD_PUSH_ENGINE=engine)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK PASS world=2" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
