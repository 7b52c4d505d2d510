"""N>1 host logic on CPU: world_size-2 gloo run of bench.py's multi-GPU plumbing (barrier, MAX over ranks,
proof partition).  The data path itself has no collective (proof-level replicas)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_covers_job():
    sys.path.insert(0, ROOT)
    from tools.dist_util import aggregate_throughput, proofs_for_rank
    for total, world in [(10, 1), (10, 2), (7, 4), (3, 8), (16, 8)]:
        parts = [proofs_for_rank(total, r, world) for r in range(world)]
        flat = [x for p in parts for x in p]
        assert flat == list(range(total))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert abs(aggregate_throughput(10, 8, 1000.0) - 80.0) < 1e-9


def test_gloo_world2_max_and_barrier(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys, json
        sys.path.insert(0, {ROOT!r})
        from tools.dist_util import Dist, proofs_for_rank, aggregate_throughput
        d = Dist(backend="gloo")
        assert d.world == 2
        d.barrier()
        mine = 10.0 + 5.0 * d.rank          # rank 1 is the slow one
        worst = d.max(mine)
        total = d.sum(len(proofs_for_rank(9, d.rank, d.world)))
        d.barrier()
        if d.rank == 0:
            print(json.dumps(dict(worst=worst, total=total, value=aggregate_throughput(4, d.world, worst))))
        d.close()
    """))
    port = free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["worst"] == 15.0 and r["total"] == 9.0
    assert abs(r["value"] - 2 * 4 / 0.015) < 1e-6
