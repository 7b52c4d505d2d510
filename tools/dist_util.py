"""Multi-GPU plumbing of bench.py: one process per GPU, proof-level replicas (no data-path collective).
The only cross-rank traffic is the barrier around the timed region and a MAX all-reduce of the elapsed time."""
import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class Dist:
    """Thin wrapper so that the same code runs single-process, under NCCL (GPU) and under gloo (CPU tests)."""

    def __init__(self, backend=None, device=None):
        self.rank, self.world, self.local_rank = env_rank()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch
            import torch.distributed as dist
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kw = {}
            if backend == "nccl":
                self.device = torch.device("cuda", self.local_rank)
                kw["device_id"] = self.device
            dist.init_process_group(backend, **kw)
            self.dist = dist
            self.backend = backend

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, x: float) -> float:
        """maximum of x over all ranks (the timing rule: a multi-GPU step takes as long as its slowest rank)"""
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device if self.backend == "nccl" else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device if self.backend == "nccl" else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def proofs_for_rank(total_proofs: int, rank: int, world: int):
    """Static partition of a job of `total_proofs` independent proofs (strong-scaling mode): contiguous blocks,
    sizes differ by at most one.  bench.py's default is weak scaling (every rank proves `steps` proofs)."""
    base, extra = divmod(total_proofs, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def aggregate_throughput(proofs_per_rank: int, world: int, max_ms: float) -> float:
    """whole-job proofs/s: all ranks' proofs divided by the slowest rank's time"""
    return world * proofs_per_rank / (max_ms / 1e3)
