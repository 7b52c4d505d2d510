"""Sharded WHIR commitment across the GPUs of one box (BASELINE config 4): launched with torchrun, one
process per GPU.  Columns of the codeword are sharded for the NTT, rows (Merkle leaves) for hashing; the exchange is
fused into the last NTT pass as NVLink peer stores into CUDA-IPC mapped leaf blocks; the "rows complete" barrier and the
all-gather of the sub-tree roots run as one-warp kernels over peer mailboxes (torch.distributed only exchanges the IPC handles).  Prints one JSON line from rank 0.

  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/sharded_commit.py --log-n 23
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=23, help="coefficients per polynomial (2 polynomials, rate 1/2)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    ap.add_argument("--same-device", action="store_true",
                    help="all ranks on cuda:0 with gloo collectives: real CUDA IPC between processes on a single-GPU box")
    ap.add_argument("--check", action="store_true", help="compare the root with a single-GPU pk_commit_batch on rank 0")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = 0 if args.same_device else local
    backend = "gloo" if args.same_device else args.backend
    torch.cuda.set_device(dev)
    if world > 1:
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo")
    coll_device = torch.device("cuda", dev) if backend == "nccl" else None
    ctx = pk.Context(dev)
    stream = torch.cuda.ExternalStream(ctx.stream)
    log_n, rate, batch = args.log_n, 1, 2
    n = 1 << log_n
    rows = 1 << (log_n + rate - 4)
    w = 16 * batch
    # every rank holds both coefficient vectors (the host uploads the witness to all GPUs); seed shared
    rng = np.random.default_rng(4)
    polys = [ctx.upload(rand_fr(rng, n)) for _ in range(batch)]
    sc = sharded.ShardedCommit(ctx, dist if world > 1 else None, rank, world, polys, log_n, rate, coll_device)
    commit = sc.commit

    for _ in range(args.warmup):
        root = commit()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        root = commit()
    e1.record(stream)
    ctx.sync()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=coll_device if coll_device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ok, open_ok, open_ms = None, None, None
    if args.check:
        # a STIR-sized opening through the sharded tree (every rank takes part), compared with the single-GPU opening
        qrng = np.random.default_rng(99)
        idx = np.unique(qrng.integers(0, rows, size=min(109, rows), dtype=np.uint64))
        if rows >= 4:
            idx = np.unique(np.concatenate([idx, np.array([0, 1, rows // 2, rows - 1], dtype=np.uint64)]))
        t0 = time.perf_counter()
        got = sc.open(idx)
        open_ms = (time.perf_counter() - t0) * 1e3
    if args.check and rank == 0:
        cm = ctx.commit_batch(polys, log_n, rate)
        ok = bool(np.array_equal(sharded.to_montgomery(root), cm.root))  # cm.root is Montgomery, ours canonical
        exp = cm.open(idx)
        open_ok = bool(np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
                       and len(got[3]) == len(exp[3]) and all(np.array_equal(a, b) for a, b in zip(got[3], exp[3])))
        cm.free()
    if rank == 0:
        nbytes_ntt = batch * 96 * n
        nbytes_mrk = 32 * (rows * w + 2 * rows - 1)
        print(json.dumps({"workload": f"sharded commit: {batch} polynomials of 2^{log_n} coefficients, rate 1/2, {rows} leaves x {w}",
                          "n_gpus": world, "same_device": args.same_device, "ms_per_commit": ms, "commits_per_s": 1e3 / ms,
                          "alg_gbs_ntt_plus_merkle": (nbytes_ntt + nbytes_mrk) / ms / 1e6,
                          "exchange": "fused into the last NTT pass (peer stores via CUDA IPC); stream barrier and all-gather of the sub-roots as "
                                      "one-warp kernels over peer mailboxes (host collectives only at set-up, backend " + backend + ")",
                          "root_matches_single_gpu": ok, "opening_matches_single_gpu": open_ok, "open_ms": open_ms,
                          "root": [int(x) for x in root]}))
    if world > 1:
        dist.barrier()
    sc.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
