"""GPU run of the sharded sumchecks: two ranks sharing cuda:0 (this test box has one GPU), gloo for the
96-byte all-gathers, local rounds through the C-ABI; every round message must equal the unsharded run bit for bit."""
import json
import os
import subprocess
import sys

import pytest

from test_dist_gloo import ROOT, free_port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,log_n", [(2, 12), (4, 9)])
def test_sharded_sumchecks_match_unsharded(world, log_n):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tools", "sharded_sumcheck.py"), "--log-n",
           str(log_n), "--backend", "gloo", "--same-device", "--check", "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert r["messages_match_unsharded"] is True and r["n_gpus"] == world
