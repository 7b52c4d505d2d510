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
           str(log_n), "--backend", "gloo", "--same-device", "--exchange", "gather", "--check", "--steps", "1", "--warmup", "0"]
    for attempt in range(2):  # the rendezvous port is picked, released and re-bound by torchrun: retry once if it was taken
        cmd[cmd.index("--master-port") + 1] = str(free_port())
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
        if out.returncode == 0:
            break
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert r["messages_match_unsharded"] is True and r["n_gpus"] == world


@pytest.mark.parametrize("world,log_n", [(2, 11), (4, 8), (8, 6)])
def test_fused_exchange_virtual_ranks(world, log_n):
    """pk_*_round_sharded with the ranks as threads of one process (one ctx / stream each, all on cuda:0, mailboxes are
    plain device buffers): the kernel-side exchange (peer stores, flag wait, field sum) must reproduce the unsharded
    messages bit for bit, for both sumchecks."""
    import hashlib
    import sys
    import threading

    if os.environ.get("CUDA_LAUNCH_BLOCKING") == "1":
        pytest.skip("the ranks' exchange kernels must run concurrently")

    import numpy as np

    sys.path.insert(0, ROOT)
    import provekit_b200 as pk
    from helpers import rand_fr
    from provekit_b200 import sharded

    def challenge(rnd, sums):
        d = hashlib.sha256(bytes([rnd]) + np.ascontiguousarray(sums).tobytes()).digest()
        return sharded._to_limbs(int.from_bytes(d, "little") % sharded.P).reshape(1, 4)

    class ThreadGather:  # hand-over gather between the threads (the per-round exchange never goes through it)
        def __init__(self, rank, shared):
            self.rank, self.world, self.shared, self.calls, self.bytes = rank, world, shared, 0, 0

        def __call__(self, x):
            self.shared["slots"][self.rank] = np.array(x, copy=True)
            self.shared["bar"].wait()
            out = np.stack(self.shared["slots"])
            self.shared["bar"].wait()
            return out

    rng = np.random.default_rng(21)
    arrs = [rand_fr(rng, 1 << log_n) for _ in range(4)]
    ctxs = [pk.Context(0) for _ in range(world)]
    boxes = [c.shard_mailbox() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.shard_group(r, world, [b.device_ptr for b in boxes])
    shared = {"slots": [None] * world, "bar": threading.Barrier(world)}
    results, errors = [None] * world, []

    def run(r):
        try:
            be = sharded.GpuBackend(ctxs[r], fused=True)
            g = ThreadGather(r, shared)
            zk = sharded.sharded_zk_sumcheck(be, g, [sharded.shard_low_bits(a, r, world) for a in arrs], log_n, challenge)
            wh = sharded.sharded_whir_sumcheck(be, g, sharded.shard_high_bits(arrs[0], r, world),
                                               sharded.shard_high_bits(arrs[1], r, world), log_n, challenge)
            results[r] = (zk, wh)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            shared["bar"].abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errors, errors
    one = sharded.Gather(None)
    be0 = sharded.GpuBackend(ctxs[0])
    zk1 = sharded.sharded_zk_sumcheck(be0, one, arrs, log_n, challenge)
    wh1 = sharded.sharded_whir_sumcheck(be0, one, arrs[0], arrs[1], log_n, challenge)
    for r in range(world):
        zk, wh = results[r]
        assert len(zk) == log_n and len(wh) == log_n
        assert all(np.array_equal(x, y) for x, y in zip(zk, zk1)), f"rank {r} zk"
        assert all(np.array_equal(x, y) for x, y in zip(wh, wh1)), f"rank {r} whir"
    for b in boxes:
        b.free()
    for c in ctxs:
        c.close()


def _run_torchrun(script, world, extra, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "0", os.path.join(ROOT, "tools", script)] + extra
    for attempt in range(2):  # the rendezvous port is picked, released and re-bound by torchrun: retry once if it was taken
        cmd[cmd.index("--master-port") + 1] = str(free_port())
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, OMP_NUM_THREADS="1"))
        if out.returncode == 0:
            break
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("world,log_n", [(2, 12), (4, 14), (8, 16)])
def test_sharded_commit_multi_process_same_device(world, log_n):
    """pk_rs_encode_sharded + pk_merkle_combine_roots across PROCESSES over real CUDA IPC (pk_ipc_export / pk_ipc_open):
    `world` ranks share cuda:0 (the test box has one GPU), every rank stores its columns into the peers' mapped leaf
    blocks, gloo carries the barrier and the 32-byte sub-roots.  The root must equal the single-process pk_commit_batch."""
    r = _run_torchrun("sharded_commit.py", world, ["--log-n", str(log_n), "--same-device", "--check", "--steps", "1", "--warmup", "1"])
    assert r["root_matches_single_gpu"] is True and r["n_gpus"] == world and r["same_device"] is True
    assert r["opening_matches_single_gpu"] is True  # sharded pk_commit_open: rows + ark MultiPath equal the single-GPU opening


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_commit_and_sumchecks_over_nvlink(world):
    """the same on `world` real GPUs (NCCL, NVLink peer stores) — skipped on boxes with fewer GPUs"""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _run_torchrun("sharded_commit.py", world, ["--log-n", "18", "--check", "--steps", "1", "--warmup", "1"])
    assert r["root_matches_single_gpu"] is True and r["opening_matches_single_gpu"] is True and r["n_gpus"] == world
    r = _run_torchrun("sharded_sumcheck.py", world, ["--log-n", "16", "--check", "--steps", "1", "--warmup", "1"])
    assert r["messages_match_unsharded"] is True and r["n_gpus"] == world
