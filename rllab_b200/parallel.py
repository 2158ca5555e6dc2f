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
        self._gather_bufs = {}
        self.n_collectives = 0
        self._owns_group = False
        if self.active and not dist.is_initialized():
            self._owns_group = True
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
            self.n_collectives += 1
        return t

    def all_reduce_max(self, t):
        if self.active:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t

    def all_reduce_mixed(self, t, n_sum):
        """In place: t[:n_sum] summed over ranks, t[n_sum:] maximised over ranks, with ONE collective: an all-gather of
        the whole vector followed by a local fixed-order reduction (b200rl_reduce_ranks).  Every reduction vector of an
        iteration is a few KB, so the cost of a collective is its launch latency, not its bytes: gathering world x n
        doubles instead of reducing n costs nothing extra, halves the number of collectives for vectors that carry sums
        and maxima (statistics; loss / KL triples), and makes the summation order rank order by construction -- every
        rank computes bit-identical results whatever algorithm NCCL picks."""
        if not self.active:
            return t
        import torch
        n = t.numel()
        key = (n, t.device, t.dtype)
        buf = self._gather_bufs.get(key)
        if buf is None:
            buf = torch.empty(self.world_size * n, dtype=t.dtype, device=t.device)
            self._gather_bufs[key] = buf
        self.dist.all_gather_into_tensor(buf, t)
        self.n_collectives += 1
        if t.is_cuda:
            from . import ops
            ops.reduce_ranks(buf, self.world_size, n, n_sum, t)
        else:                    # CPU tensors only occur in the gloo tests of this plumbing (no kernels involved)
            g = buf.view(self.world_size, n)
            t[:n_sum] = g[:, :n_sum].sum(0)
            if n_sum < n:
                t[n_sum:] = g[:, n_sum:].max(0).values
        return t

    def barrier(self):
        if self.active:
            self.dist.barrier()

    def close(self):
        """Tear the process group down (NCCL warns at exit otherwise); only if this object created it."""
        if self.active and self._owns_group and self.dist.is_initialized():
            self.dist.destroy_process_group()
            self._owns_group = False

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
