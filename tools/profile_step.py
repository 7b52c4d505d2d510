"""One warm-up proof, then ONE proof inside cudaProfilerStart/Stop — the command ncu wraps
(`ncu --profile-from-start off ...`).  A number printed under ncu is never a bench value."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import provekit_b200 as pk  # noqa: E402
from tools import workload as wl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="poseidon-1000")
args = ap.parse_args()
from bench import WORKLOADS  # noqa: E402

r1cs = wl.synth_r1cs(**WORKLOADS[args.workload], seed=1)
rnd = wl.randomness(r1cs)
ctx = pk.Context(0)
prover = pk.Prover(ctx, r1cs)
prover.prove(r1cs["witness"], rnd)
prover.upload_inputs(r1cs["witness"], rnd)
l0 = ctx.launches
torch.cuda.profiler.start()
proof = prover.prove_staged()
ctx.sync()
torch.cuda.profiler.stop()
print("launches in profiled proof:", ctx.launches - l0, "proof bytes:", len(proof))
