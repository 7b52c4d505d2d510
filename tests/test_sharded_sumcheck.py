"""Host logic of the sharded sumchecks (provekit_b200/sharded.py) on CPU: world_size-2 and -4 gloo runs in which
the local round is the ORACLE's round (the checker standing in for the kernel), compared with the unsharded oracle run.
The GPU path (same host logic over pk_zk_sumcheck_round / pk_whir_sumcheck_round) is tools/sharded_sumcheck.py and
tests/test_gpu_sharded.py."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from helpers import rand_fr
from test_dist_gloo import ROOT, free_port

sys.path.insert(0, ROOT)
from provekit_b200 import sharded  # noqa: E402


def test_shard_round_trip():
    rng = np.random.default_rng(1)
    a = rand_fr(rng, 64)
    for world in (1, 2, 4, 8):
        lo = [sharded.shard_low_bits(a, r, world) for r in range(world)]
        hi = [sharded.shard_high_bits(a, r, world) for r in range(world)]
        assert np.array_equal(sharded.unshard_low_bits(lo), a)
        assert np.array_equal(sharded.unshard_high_bits(hi), a)
        # low-bit sharding keeps MSB pairs (i, i + n/2) on one rank; high-bit sharding keeps LSB pairs (2i, 2i+1)
        for r in range(world):
            assert np.array_equal(lo[r][: 32 // world], a[r:32:world]) and np.array_equal(lo[r][32 // world:], a[32 + r::world])
            assert np.array_equal(hi[r][0::2], a[r * 64 // world:(r + 1) * 64 // world:2])


def test_field_sum_is_modular():
    P = sharded.P
    parts = np.zeros((3, 2, 4), np.uint64)
    vals = [[P - 1, 5], [P - 2, 7], [3, P - 12]]
    for r in range(3):
        for j in range(2):
            parts[r, j] = sharded._to_limbs(vals[r][j])
    out = sharded.field_sum(parts)
    assert sharded._to_int(out[0]) == (P - 1 + P - 2 + 3) % P and sharded._to_int(out[1]) == 0


WORKER = """
import sys, json, hashlib, ctypes
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import torch.distributed as dist
import oracle
from helpers import rand_fr, ptr
from provekit_b200 import sharded

orc = oracle.lib()

class OracleBackend:
    # host arrays stand in for device buffers; the oracle round is the local kernel
    def upload(self, a): return np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    def alloc(self, n): return np.zeros((n, 4), np.uint64)
    def download(self, b, n): return b[:n].copy()
    def free(self, b): pass
    def zk_round(self, bufs, log_n, fold, local=False):
        out = np.zeros((3, 4), np.uint64)
        orc.orc_zk_sumcheck_round(*[ptr(b) for b in bufs], log_n, ptr(fold) if fold is not None else None, ptr(out))
        return out
    def whir_round(self, src, dst, log_n, fold, local=False):
        out = np.zeros((3, 4), np.uint64)
        orc.orc_whir_sumcheck_round(ptr(src[0]), ptr(src[1]), log_n, ptr(fold) if fold is not None else None, ptr(out))
        if fold is not None:  # the oracle folds in place; the kernel writes (p_out, w_out)
            h = 1 << (log_n - 1)
            dst[0][:h] = src[0][:h]; dst[1][:h] = src[1][:h]
        return out

def challenge(rnd, sums):
    d = hashlib.sha256(bytes([rnd]) + np.ascontiguousarray(sums).tobytes()).digest()
    v = int.from_bytes(d, "little") % sharded.P
    return sharded._to_limbs(v).reshape(1, 4)

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = sharded.Gather(dist)
log_n = {log_n}
rng = np.random.default_rng(7)
arrs = [rand_fr(rng, 1 << log_n) for _ in range(4)]
be = OracleBackend()
zk = sharded.sharded_zk_sumcheck(be, g, [sharded.shard_low_bits(a, rank, world) for a in arrs], log_n, challenge)
wh = sharded.sharded_whir_sumcheck(be, g, sharded.shard_high_bits(arrs[0], rank, world), sharded.shard_high_bits(arrs[1], rank, world), log_n, challenge)
one = sharded.Gather(None)
zk1 = sharded.sharded_zk_sumcheck(be, one, arrs, log_n, challenge)
wh1 = sharded.sharded_whir_sumcheck(be, one, arrs[0], arrs[1], log_n, challenge)
ok = all(np.array_equal(a, b) for a, b in zip(zk, zk1)) and all(np.array_equal(a, b) for a, b in zip(wh, wh1))
res = np.array([1 if ok and len(zk) == log_n and len(wh) == log_n else 0, g.calls], dtype=np.uint64)
allres = g(res)
if rank == 0:
    print(json.dumps(dict(ok=[int(x[0]) for x in allres], gathers=int(res[1]), world=world)))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world,log_n", [(2, 6), (4, 5), (2, 2)])
def test_gloo_sharded_sumchecks_equal_unsharded(tmp_path, world, log_n):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(WORKER.format(root=ROOT, log_n=log_n)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert r["ok"] == [1] * world and r["world"] == world
    lg = world.bit_length() - 1
    # data-path collectives: one 96-byte all-gather per sharded round, plus the hand-over gathers (4 + 2 arrays)
    assert r["gathers"] == 2 * (log_n - lg) + 6
