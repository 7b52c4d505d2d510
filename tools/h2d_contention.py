"""Why the host-mask path (pk_prove_enqueue, 128 MB H2D per proof) is slower than the seeded one: proofs/s with device-drawn masks,
with host masks, and with device-drawn masks next to the same H2D bytes on an unrelated stream.  -> profiles/r02_h2d_contention.jsonl"""
import os, sys, json, time, collections, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import provekit_b200 as pk
from tools import workload as wl
from bench import WORKLOADS
n_fl = int(os.environ.get("N_FL", "16")); steps = 160
r1cs = wl.synth_r1cs(**WORKLOADS["poseidon-1000"], seed=1)
rnd = wl.randomness(r1cs, seed=7)
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t.numpy(), t
witness, wk = pin(r1cs["witness"]); rp, keep = {}, []
for k, v in rnd.items():
    rp[k], t = pin(v); keep.append(t)
ctxs = [pk.Context(0) for _ in range(n_fl)]; provers = [pk.Prover(c, r1cs) for c in ctxs]
seed = bytes(range(32))
def run(enq, steps, bg=None):
    for c in ctxs: c.sync()
    t0 = time.perf_counter(); q = collections.deque()
    for k in range(steps):
        p_ = provers[k % n_fl]
        if len(q) == n_fl: q.popleft().collect()
        if bg: bg()
        enq(p_); q.append(p_)
    while q: q.popleft().collect()
    for c in ctxs: c.sync()
    torch.cuda.synchronize()
    return steps / (time.perf_counter() - t0)
enq_seed = lambda p_: p_.enqueue_seeded(witness, seed)
enq_mask = lambda p_: p_.enqueue(witness, rp)
for p_ in provers: p_.prove_seeded(witness, seed)
run(enq_seed, 3 * n_fl); run(enq_mask, n_fl)
out = {"n_fl": n_fl, "chunk_mb": os.environ.get("PK_H2D_CHUNK_MB")}
out["seeded"] = run(enq_seed, steps)
out["masks"] = run(enq_mask, steps)
# seeded proofs + the same H2D bytes as background traffic on an unrelated stream
side = torch.cuda.Stream()
gw = keep[1].view(torch.int64).reshape(-1); mw = keep[0].view(torch.int64).reshape(-1)
d1 = torch.empty_like(gw, device="cuda"); d2 = torch.empty_like(mw, device="cuda")
def bg():
    with torch.cuda.stream(side):
        d1.copy_(gw, non_blocking=True); d2.copy_(mw, non_blocking=True)
out["seeded_plus_background_h2d"] = run(enq_seed, steps, bg)
print(json.dumps(out))
