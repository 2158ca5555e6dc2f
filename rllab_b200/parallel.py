"""Multi-GPU plumbing: one process per GPU (torchrun), lanes sharded by rank, NCCL all-reduce of the small float64
reduction vectors (gradient, FVP, loss/KL scalars, advantage statistics, baseline normal equations) over
NVLink/NVSwitch.  The reference has no counterpart (its update is single-process): this is row (e) of
the scope table.  With world_size == 1 every method is a no-op (no communicator is created)."""
import os


class Comm(object):
    def __init__(self, backend=None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.active = self.world_size > 1
        if self.active and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend=backend, **kw)

    def all_reduce_sum(self, t):
        if self.active:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def all_reduce_max(self, t):
        if self.active:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t

    def barrier(self):
        if self.active:
            self.dist.barrier()

    def shard(self, n_total):
        """Contiguous lane block of this rank: lane i -> GPU floor(i*G/N)."""
        per = n_total // self.world_size
        rem = n_total % self.world_size
        n = per + (1 if self.rank < rem else 0)
        lane0 = self.rank * per + min(self.rank, rem)
        return n, lane0


_default = None


def default_comm():
    global _default
    if _default is None:
        _default = Comm()
    return _default
