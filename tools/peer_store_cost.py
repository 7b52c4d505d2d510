"""Cost of fusing the codeword exchange into the last NTT pass: pk_rs_encode_sharded with local vs remote (peer GPU) leaf
blocks, one process, two GPUs.  python tools/peer_store_cost.py  ->  profiles/r02_peer_store_cost.jsonl"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import provekit_b200 as pk
from tools.workload import rand_fr
# enable peer access both ways through torch (one process, two devices)
a = torch.zeros(1 << 20, device="cuda:0"); b = torch.zeros(1 << 20, device="cuda:1")
b.copy_(a); a.copy_(b); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
print(json.dumps(dict(can_access=[torch.cuda.can_device_access_peer(0, 1), torch.cuda.can_device_access_peer(1, 0)])))
ctx0, ctx1 = pk.Context(0), pk.Context(1)
torch.cuda.set_device(0)
stream = torch.cuda.ExternalStream(ctx0.stream)
def timeit(fn, reps=3):
    fn(); ctx0.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); ctx0.sync(); ts.append(e0.elapsed_time(e1))
    return min(ts)
rng = np.random.default_rng(1)
log_n = 23
poly = ctx0.upload(rand_fr(rng, 1 << log_n))
rows = 1 << (log_n + 1 - 4)
for n_peers in (2, 4, 8):
    per = rows // n_peers
    loc = [ctx0.buffer(per * 32) for _ in range(n_peers)]
    rem = [ctx1.buffer_shared(per * 32) for _ in range(n_peers)]
    for nc in (16, 8, 4):
        t_loc = timeit(lambda: ctx0.rs_encode_sharded(poly, log_n, 1, 0, nc, [x.device_ptr for x in loc], 32, 0))
        ptrs = [loc[0].device_ptr] + [x.device_ptr for x in rem[1:]]
        t_rem = timeit(lambda: ctx0.rs_encode_sharded(poly, log_n, 1, 0, nc, ptrs, 32, 0))
        remote_bytes = (n_peers - 1) * per * nc * 32
        print(json.dumps(dict(k="rs_encode_sharded", n_peers=n_peers, n_cols=nc, ms_local=t_loc, ms_remote=t_rem,
                              remote_mb=remote_bytes / 1e6, extra_ms=t_rem - t_loc,
                              remote_gbs_if_serial=remote_bytes / max(t_rem - t_loc, 1e-6) / 1e6)))
    for x in loc + rem: x.free()
# plain device-to-device copy rate 0 -> 1 (copy engine) and an SM copy kernel (torch) for comparison
src = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0"); dst = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:1")
for name, fn in (("memcpy_peer", lambda: dst.copy_(src, non_blocking=True)),):
    fn(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print(json.dumps(dict(k=name, gbs=(1 << 28) / e0.elapsed_time(e1) / 1e6)))
