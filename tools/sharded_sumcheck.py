"""Sharded sumchecks across the GPUs of one box (SURVEY 8e "Sumcheck/fold"): launched with torchrun, one process per GPU.
zk-sumcheck arrays (a, b, c, eq) are sharded by the low index bits, WHIR (p, w) by the high bits; every round is the
unchanged single-GPU kernel on the local shard plus ONE all-gather of 96 B per rank (NCCL) and a G-term modular add; the
last log2 G rounds run on the gathered 2G survivors.  Host logic: provekit_b200/sharded.py.  Prints one JSON line.

  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/sharded_sumcheck.py --log-n 22 --check
  (--backend gloo --same-device: all ranks on cuda:0 with CPU collectives; used by tests/test_gpu_sharded.py on 1 GPU)
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import provekit_b200 as pk  # noqa: E402
from provekit_b200 import sharded  # noqa: E402
from tools.workload import rand_fr  # noqa: E402


def challenge(rnd, sums):
    """stand-in for the transcript: every rank derives the same fold value from the round message"""
    d = hashlib.sha256(bytes([rnd & 0xFF]) + np.ascontiguousarray(sums).tobytes()).digest()
    return sharded._to_limbs(int.from_bytes(d, "little") % sharded.P).reshape(1, 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=22, help="zk-sumcheck over 4 arrays of 2^log_n; WHIR over 2 of 2^(log_n+1)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    ap.add_argument("--small-collective", default="gloo", choices=["gloo", "nccl"],
                    help="transport of the 96-byte round messages when the world is NCCL")
    ap.add_argument("--exchange", default="fused", choices=["fused", "gather"],
                    help="fused: partial sums exchanged by a kernel over CUDA-IPC peer memory (pk_*_round_sharded); "
                         "gather: host-side all-gather per round")
    ap.add_argument("--same-device", action="store_true", help="all ranks on cuda:0 (single-GPU box test mode)")
    ap.add_argument("--check", action="store_true", help="rank 0 also runs unsharded and compares every round message")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = 0 if args.same_device else local
    torch.cuda.set_device(dev)
    d = None
    if world > 1:
        if args.backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo")
        d = dist
    if d is not None and args.backend == "nccl" and args.small_collective == "nccl":
        gather = sharded.Gather(d, torch.device("cuda", dev))
    elif d is not None and args.backend == "nccl":
        gather = sharded.Gather(d, None, group=d.new_group(backend="gloo"))  # 96-byte messages stay on the host
    else:
        gather = sharded.Gather(d, None)
    ctx = pk.Context(dev)
    fused = args.exchange == "fused" and world > 1
    be = sharded.GpuBackend(ctx, fused=fused)
    peers = []
    if fused:
        mbox = ctx.shard_mailbox()
        handles = [None] * world
        d.all_gather_object(handles, ctx.ipc_export(mbox))
        peers = [mbox.device_ptr if r == rank else ctx.ipc_open(handles[r]) for r in range(world)]
        ctx.shard_group(rank, world, peers)
        d.barrier()  # every mailbox is zeroed and mapped before the first round
    m0, m = args.log_n, args.log_n + 1
    rng = np.random.default_rng(8)
    zk_full = [rand_fr(rng, 1 << m0) for _ in range(4)]
    wh_full = [rand_fr(rng, 1 << m) for _ in range(2)]
    # the local shards live on the device (in a sharded prover they are produced there); every step folds a fresh clone
    zk_loc = [ctx.upload(sharded.shard_low_bits(a, rank, world)) for a in zk_full]
    wh_loc = [ctx.upload(sharded.shard_high_bits(a, rank, world)) for a in wh_full]
    whir_rounds = 4  # one WHIR folding step (FoldingFactor::Constant(4)); afterwards the polynomial is re-committed

    def step():
        zk_in, wh_in = [b.clone() for b in zk_loc], [b.clone() for b in wh_loc]
        ctx.sync()
        if d is not None:
            d.barrier()
        t0 = time.perf_counter()
        zk = sharded.sharded_zk_sumcheck(be, gather, zk_in, m0, challenge)
        ctx.sync()
        t1 = time.perf_counter()
        wh = sharded.sharded_whir_sumcheck(be, gather, wh_in[0], wh_in[1], m, challenge, rounds=whir_rounds)
        ctx.sync()
        t2 = time.perf_counter()
        return zk, wh, (t1 - t0) * 1e3, (t2 - t1) * 1e3

    for _ in range(args.warmup):
        step()
    if d is not None:
        d.barrier()
    zk_ms, wh_ms = [], []
    for _ in range(args.steps):
        zk, wh, a, b = step()
        zk_ms.append(a)
        wh_ms.append(b)
    ms = np.array([min(zk_ms), min(wh_ms)])
    if d is not None:  # a sharded step takes as long as its slowest rank
        t = torch.tensor(ms, dtype=torch.float64, device=torch.device("cuda", dev) if args.backend == "nccl" else "cpu")
        d.all_reduce(t, op=d.ReduceOp.MAX)
        ms = t.cpu().numpy()
    ok = None
    if args.check and rank == 0:
        one = sharded.Gather(None)
        zk1 = sharded.sharded_zk_sumcheck(be, one, zk_full, m0, challenge)
        wh1 = sharded.sharded_whir_sumcheck(be, one, wh_full[0], wh_full[1], m, challenge, rounds=whir_rounds)
        ok = bool(len(zk) == m0 and len(wh) == whir_rounds and all(np.array_equal(x, y) for x, y in zip(zk, zk1))
                  and all(np.array_equal(x, y) for x, y in zip(wh, wh1)))
    if rank == 0:
        n0, n1 = 1 << m0, 1 << m
        zk_bytes = 32 * 4 * (n0 + sum(n0 >> (i - 1) for i in range(1, m0)) + sum(n0 >> i for i in range(1, m0)))
        print(json.dumps({"workload": f"sharded sumchecks: zk over 4 x 2^{m0} (all {m0} rounds), WHIR over 2 x 2^{m} ({whir_rounds} rounds)",
                          "n_gpus": world, "backend": args.backend, "exchange": "fused (NVLink peer stores + device-side sum)" if fused else "host all-gather over " + (args.small_collective if args.backend == "nccl" else "gloo"), "same_device": args.same_device,
                          "zk_sumcheck_ms": float(ms[0]), "whir_sumcheck_ms": float(ms[1]),
                          "zk_alg_gbs": zk_bytes / float(ms[0]) / 1e6,
                          "includes": "all rounds with the shards resident in HBM: local kernel, D2H of the 96-byte partial message, all-gather, modular sum, challenge",
                          "host_collectives": f"{gather.calls} all-gathers, {gather.bytes} B sent per rank over all steps (hand-over of the 2-element shards; plus 96 B per round when --exchange gather)",
                          "messages_match_unsharded": ok}))
    for r, p in enumerate(peers):
        if r != rank:
            ctx.ipc_close(p)
    ctx.close()
    if d is not None:
        d.destroy_process_group()


if __name__ == "__main__":
    main()
