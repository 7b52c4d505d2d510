"""Sumchecks sharded across the GPUs of one box ("Sumcheck/fold"): one process per GPU, the only data-path
exchange is an all-gather of three field elements (96 B) per rank and round.

* ProveKit's zk-sumcheck (`sumcheck_fold_map_reduce`, provekit/common/src/utils/sumcheck.rs:16-39) pairs index i with
  i + len/2 (binds the MOST significant variable), so the arrays are sharded by the LOW index bits: rank g owns
  x[(j << log2 G) | g].  A global pair (i, i + N/2) is the local pair (j, j + N/(2G)) of one rank, every round of the
  first m_0 - log2 G is the unchanged single-GPU kernel on the local shard, and the round message is the field sum of
  the ranks' partial sums.
* WHIR's sumcheck (`SumcheckSingle` [whir]) pairs 2i with 2i+1 (binds the LEAST significant variable), so the arrays
  are sharded by the HIGH index bits (contiguous blocks).
* When a shard is down to two elements the 2G survivors are gathered onto every rank and the last log2 G rounds run on
  one GPU (they are 2G..4 elements long).

The local round is injected (`backend`), so the same host logic runs over the CUDA kernels (GpuBackend, the product
path) and, in the CPU gloo tests, over the oracle.  Nothing here computes field arithmetic on the CPU except the
G-term modular sum of the gathered partials (NCCL has no 256-bit modular reduction).
"""
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def log2_exact(n: int) -> int:
    k = n.bit_length() - 1
    if n <= 0 or (1 << k) != n:
        raise ValueError(f"{n} is not a power of two")
    return k


def shard_low_bits(arr: np.ndarray, rank: int, world: int) -> np.ndarray:
    """x[(j << log2 G) | g] for j = 0.. : the zk-sumcheck shard of rank g"""
    return np.ascontiguousarray(arr.reshape(-1, 4)[rank::world])


def shard_high_bits(arr: np.ndarray, rank: int, world: int) -> np.ndarray:
    """contiguous block g of the array: the WHIR-sumcheck shard of rank g"""
    a = arr.reshape(-1, 4)
    per = a.shape[0] // world
    return np.ascontiguousarray(a[rank * per:(rank + 1) * per])


def unshard_low_bits(parts) -> np.ndarray:
    """inverse of shard_low_bits: parts[g][j] -> x[(j << log2 G) | g]"""
    world = len(parts)
    n = parts[0].reshape(-1, 4).shape[0]
    out = np.empty((n * world, 4), np.uint64)
    for g, p in enumerate(parts):
        out[g::world] = p.reshape(-1, 4)
    return out


def unshard_high_bits(parts) -> np.ndarray:
    return np.concatenate([p.reshape(-1, 4) for p in parts], axis=0)


def _to_int(limbs) -> int:
    return sum(int(x) << (64 * i) for i, x in enumerate(limbs))


def _to_limbs(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def field_sum(partials: np.ndarray) -> np.ndarray:
    """(G, k, 4) Montgomery-form partial sums -> (k, 4): addition mod p commutes with the Montgomery map"""
    partials = np.asarray(partials, dtype=np.uint64)
    g, k = partials.shape[0], partials.shape[1]
    out = np.empty((k, 4), np.uint64)
    for j in range(k):
        out[j] = _to_limbs(sum(_to_int(partials[r, j]) for r in range(g)) % P)
    return out


class Gather:
    """all-gather of small uint64 arrays over torch.distributed (nccl on GPUs, gloo in the CPU tests)"""

    def __init__(self, dist=None, device=None, group=None):
        """device: where the collective runs (a cuda device for nccl; None = host tensors, for a gloo group).  The
        messages are 96 B and already on the host (the C-ABI returns them), so a host-side gloo group next to the NCCL
        world avoids an H2D + D2H per round."""
        self.dist, self.device, self.group = dist, device, group
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.calls = self.bytes = 0

    def __call__(self, x: np.ndarray) -> np.ndarray:
        """x: any uint64 array -> (world, *x.shape)"""
        x = np.ascontiguousarray(x, dtype=np.uint64)
        if self.dist is None:
            return x[None]
        import torch
        t = torch.from_numpy(x.view(np.int64).reshape(-1).copy())
        if self.device is not None:
            t = t.to(self.device)
        out = torch.empty(self.world * t.numel(), dtype=torch.int64, device=t.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        self.calls += 1
        self.bytes += x.nbytes
        return out.cpu().numpy().view(np.uint64).reshape((self.world,) + x.shape)


class GpuBackend:
    """local rounds on this rank's GPU through the C-ABI (pk_zk_sumcheck_round / pk_whir_sumcheck_round)"""

    def __init__(self, ctx, fused: bool = False):
        """fused: the ranks' contexts are joined by pk_shard_group_set; a local round then returns the GLOBAL message
        (partial sums exchanged by the kernel over NVLink peer stores) and no host-side collective is needed."""
        self.ctx, self.fused = ctx, fused

    def upload(self, arr):
        """host array -> new device buffer; a device buffer is adopted as is (the sumcheck folds it in place and frees it)"""
        return arr if hasattr(arr, "device_ptr") else self.ctx.upload(arr)

    def alloc(self, n):
        return self.ctx.buffer(n)

    def download(self, buf, n):
        return buf.download(n)

    def free(self, buf):
        buf.free()

    def zk_round(self, bufs, log_n, fold, local=False):
        if self.fused and local:
            return self.ctx.sumcheck_fold_map_reduce_sharded(*bufs, log_n, fold)
        return self.ctx.sumcheck_fold_map_reduce(*bufs, log_n, fold)

    def whir_round(self, src, dst, log_n, fold, local=False):
        f = self.ctx.whir_sumcheck_round_sharded if (self.fused and local) else self.ctx.whir_sumcheck_round
        if fold is None:
            return f(src[0], src[1], log_n)
        return f(src[0], src[1], log_n, fold, dst[0], dst[1])


def sharded_zk_sumcheck(backend, gather: Gather, local_arrays, log_n: int, challenge, rounds: int = None):
    """All `rounds` (default log_n) rounds of the zk-sumcheck over 4 arrays of 2^log_n elements of which this rank holds the
    low-bit shard `local_arrays` (host arrays of 2^log_n / G elements).  challenge(round, sums) -> fold value (4 limbs,
    the transcript's job).  Returns the list of round messages [f(0), f(-1), f(inf)] as (3,4) arrays — identical on
    every rank and identical to the unsharded run."""
    world = gather.world
    lg = log2_exact(world)
    rounds = log_n if rounds is None else rounds
    if log_n - lg < 1:
        raise ValueError("every rank needs at least two elements")
    bufs = [backend.upload(a) for a in local_arrays]
    cur = log_n - lg  # local length = 2^cur
    msgs, fold = [], None
    local = True
    for rnd in range(rounds):
        if local and fold is not None and cur == 1:
            # two elements per rank left: gather the 2G survivors (index = (j << lg) | g) and finish on one GPU
            parts = [gather(backend.download(b, 2)) for b in bufs]
            for b in bufs:
                backend.free(b)
            bufs = [backend.upload(unshard_low_bits(list(p))) for p in parts]
            cur, local = lg + 1, False
        sharded_round = local and world > 1
        part = backend.zk_round(bufs, cur, fold, local=sharded_round)
        if fold is not None:
            cur -= 1
        sums = field_sum(gather(part)) if sharded_round and not getattr(backend, "fused", False) else part
        msgs.append(sums)
        fold = challenge(rnd, sums)
    for b in bufs:
        backend.free(b)
    return msgs


def sharded_whir_sumcheck(backend, gather: Gather, local_p, local_w, log_n: int, challenge, rounds: int = None):
    """WHIR sumcheck rounds [h(0), h(1), h(2)] over (p, w) of 2^log_n elements, sharded by the high index bits."""
    world = gather.world
    lg = log2_exact(world)
    rounds = log_n if rounds is None else rounds
    if log_n - lg < 1:
        raise ValueError("every rank needs at least two elements")
    cur = log_n - lg
    src = (backend.upload(local_p), backend.upload(local_w))
    dst = (backend.alloc(max(1 << (cur - 1), 1)), backend.alloc(max(1 << (cur - 1), 1)))
    msgs, fold = [], None
    local = True
    for rnd in range(rounds):
        if local and fold is not None and cur == 1:
            parts = [gather(backend.download(b, 2)) for b in src]
            for b in src + dst:
                backend.free(b)
            src = tuple(backend.upload(unshard_high_bits(list(p))) for p in parts)
            dst = (backend.alloc(world), backend.alloc(world))
            cur, local = lg + 1, False
        sharded_round = local and world > 1
        part = backend.whir_round(src, dst, cur, fold, local=sharded_round)
        if fold is not None:
            cur -= 1
            src, dst = dst, src
        sums = field_sum(gather(part)) if sharded_round and not getattr(backend, "fused", False) else part
        msgs.append(sums)
        fold = challenge(rnd, sums)
    for b in src + dst:
        backend.free(b)
    return msgs
