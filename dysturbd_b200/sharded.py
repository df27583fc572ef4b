"""One coupled population over several ranks (one process per GPU, torch.distributed).

Agents are partitioned by community into equal contiguous index ranges (households never straddle a boundary, so phases
B, C, E, F are shard-local).  The only cross-shard coupling of a day is phase D (contagious3.py:230-248): a susceptible
agent on one shard may have an infectious contact on another.  Everything a day needs to know about a remote agent fits
in its one-byte day-start snapshot (infectious?, isolated?, days since infection), so the per-day exchange is ONE
all-gather of `N_glob` bytes (NCCL over NVLink on the GPU box; gloo in the CPU tests), placed between the two halves of
the day (dys_step_begin / dys_step_end).  Random draws are keyed by GLOBAL agent indices, so a sharded run is
bit-identical to the unsharded one.  The per-day integer statistics are sums over shards (an all-reduce of 64 integers).
"""
import numpy as np
import torch
import torch.distributed as dist


class _CudaAlias:
    """a device buffer owned by the C library, viewed as a torch tensor (zero copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class ShardedEngine:
    def __init__(self, pop, params, backend=None, device=None, group=None, **kw):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group, self.pop = group, pop
        if pop.N * self.world != pop.N_glob or pop.lo != self.rank * pop.N:
            raise ValueError("shards must be equal, contiguous and in rank order")
        self.on_gpu = backend is None
        if self.on_gpu:
            from . import engine
            self.sim = engine.Engine(pop, params, device=device, rank=self.rank, world=self.world, **kw)
            ptr, nbytes = self.sim.device_buffer(engine.BUF_INFECTIOUS)
            self.ib = torch.as_tensor(_CudaAlias(ptr, nbytes), device=torch.device("cuda", device))[:pop.N_glob]
            self.stream = torch.cuda.ExternalStream(self.sim.cuda_stream(), device=device)
        else:                                   # tests: a CPU stand-in with the same two-half interface
            self.sim = backend(pop, params)
            self.ib = torch.zeros(pop.N_glob, dtype=torch.uint8)

    def exchange(self):
        """all-gather of the day-start infectious bytes, in place where the library reads them"""
        lo, n = self.pop.lo, self.pop.N
        if self.on_gpu:
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.ib, self.ib[lo:lo + n], group=self.group)
        else:
            mine = torch.from_numpy(self.sim.ib_local().copy())
            parts = [torch.zeros(n, dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            self.ib = torch.cat(parts)
            self.sim.set_ib_global(self.ib.numpy())

    def step(self, day):
        self.sim.step_begin(day)
        self.exchange()
        self.sim.step_end(day)

    def set_open(self, open_flags):
        self.sim.set_open(open_flags)

    def local_stats(self, day):
        return self.sim.stats(day)

    def stats(self, day):
        """global stats block: every entry is a sum over shards"""
        t = torch.from_numpy(np.ascontiguousarray(self.sim.stats(day), dtype=np.int64).copy())
        if self.on_gpu:
            t = t.cuda()
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def state(self):
        return self.sim.state()

    def close(self):
        if hasattr(self.sim, "close"):
            self.sim.close()
