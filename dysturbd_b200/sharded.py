"""One coupled population over several ranks (one process per GPU, torch.distributed).

Agents are partitioned by community into equal contiguous index ranges (households never straddle a boundary, so phases
B, C, E, F are shard-local).  The only cross-shard coupling of a day is phase D (contagious3.py:230-248): a susceptible
agent on one shard may have an infectious contact on another.  Everything a day needs to know about a remote agent fits
in its one-byte day-start snapshot (infectious?, isolated?, days since infection).  Two ways to exchange it:

  mode "peer"  (default on GPUs): the exchange is FUSED into the producing kernel.  Ranks swap CUDA IPC handles once;
               afterwards k_agent stores tomorrow's snapshot bytes straight into every peer GPU's buffer (P2P stores over
               NVLink / NVSwitch) and raises a flag there, and the next day's k_push waits for all flags.  No collective
               call, no extra launch, no host involvement per day.
  mode "nccl": one all-gather of N_glob bytes per day between the two halves of the day (dys_step_begin / dys_step_end),
               in place in the library's buffer.  Also what the CPU tests exercise (gloo, C oracle as the backend).

Random draws are keyed by GLOBAL agent indices, so a sharded run is bit-identical to the unsharded one.  The per-day
integer statistics are sums over shards (an all-reduce of 64 integers when the host needs them).
"""
import numpy as np
import torch
import torch.distributed as dist


class _CudaAlias:
    """a device buffer owned by the C library, viewed as a torch tensor (zero copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class ShardedEngine:
    def __init__(self, pop, params, backend=None, device=None, group=None, mode="peer", **kw):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group, self.pop = group, pop
        if pop.N * self.world != pop.N_glob or pop.lo != self.rank * pop.N:
            raise ValueError("shards must be equal, contiguous and in rank order")
        self.on_gpu = backend is None
        self.mode = mode if self.on_gpu else "nccl"
        if self.on_gpu:
            from . import engine
            self.sim = engine.Engine(pop, params, device=device, rank=self.rank, world=self.world, **kw)
            ptr, nbytes = self.sim.device_buffer(engine.BUF_INFECTIOUS)
            both = torch.as_tensor(_CudaAlias(ptr, nbytes), device=torch.device("cuda", device))
            half = nbytes // 2
            self.ib = [both[:half][:pop.N_glob], both[half:][:pop.N_glob]]      # buffer of day d is d & 1
            self.stream = torch.cuda.ExternalStream(self.sim.cuda_stream(), device=device)
            if self.mode == "peer":
                handles = [None] * self.world
                dist.all_gather_object(handles, (self.sim.peer_export(), (pop.lo, pop.lo + pop.N) + tuple(pop.halo)), group=group)
                ok = 1
                try:
                    self.sim.peer_attach([x[0] for x in handles], [x[1] for x in handles])
                except Exception:                  # no P2P path to some rank: every rank falls back together
                    ok = 0
                t = torch.tensor([ok], device=torch.device("cuda", device))
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
                if int(t.item()) == 0:
                    self.sim.peer_detach()
                    self.mode = "nccl"
                self.sim.sync()
                dist.barrier(group=group)          # nobody steps before every rank's flags are reset and attached
        else:                                   # tests: a CPU stand-in with the same two-half interface
            self.sim = backend(pop, params)

    def exchange(self, day):
        """mode "nccl": all-gather of the day-start infectious bytes, in place where the library reads them"""
        lo, n = self.pop.lo, self.pop.N
        if self.on_gpu:
            buf = self.ib[day & 1]
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(buf, buf[lo:lo + n], group=self.group)
        else:
            mine = torch.from_numpy(self.sim.ib_local().copy())
            parts = [torch.zeros(n, dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            self.sim.set_ib_global(torch.cat(parts).numpy())

    def step(self, day):
        if self.mode == "peer":
            self.sim.step(day)
        else:
            self.sim.step_begin(day)
            self.exchange(day)
            self.sim.step_end(day)

    def step_many(self, first_day, n_days, open_flags=None):
        if self.mode == "peer":
            self.sim.step_many(first_day, n_days, open_flags)
        else:
            for k in range(n_days):
                if open_flags is not None:
                    self.sim.set_open(open_flags[k])
                self.step(first_day + k)

    def set_open(self, open_flags):
        self.sim.set_open(open_flags)

    def set_state(self, *state):
        self.sim.set_state(*state)
        if self.on_gpu:
            self.sim.sync()
            dist.barrier(group=self.group)

    def outputs(self, day):
        return self.sim.outputs(day)

    def sync(self):
        if hasattr(self.sim, "sync"):
            self.sim.sync()

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
