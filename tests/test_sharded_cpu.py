"""The N>1 path on CPU: two gloo ranks, each holding its own shard of ONE coupled synthetic population (generated
shard-locally), exchange the day-start infectious bytes every day through dysturbd_b200.sharded.ShardedEngine (the
same host code the GPU run uses; the C oracle stands in for the device library).  Because draws are keyed by global
agent indices the sharded run must be bit-identical to the unsharded oracle."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_COMM, DAYS = 4, 30


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dysturbd_b200.population import EpiParams
        from dysturbd_b200.sharded import ShardedEngine
        from dysturbd_b200.synth import COMMUNITY, synth_population
        from oracle import port as oracle_port
        n = COMMUNITY * N_COMM
        cpr = N_COMM // world
        pop = synth_population(n, seed=21, communities=(rank * cpr, (rank + 1) * cpr), seeds_per_22493=150, cross=0.05)
        prm = EpiParams(norm_factor=6.0, seed=77)
        sh = ShardedEngine(pop, prm, backend=oracle_port.OracleSim)     # CPU stand-in -> mode "nccl" over gloo
        ref = oracle_port.OracleSim(synth_population(n, seed=21, seeds_per_22493=150, cross=0.05), prm) if rank == 0 else None
        ok, cross_events = True, 0
        for d in range(DAYS):
            sh.step(d)
            g = sh.stats(d)
            if rank == 0:
                ref.step(d)
                keep = [i for i in range(18) if i not in (14, 15)]
                ok &= bool(np.array_equal(g[keep], ref.stats()[keep]))
                ok &= bool(np.array_equal(g[32:53], ref.stats()[32:53]))
                st, rs = sh.state(), ref.state()
                for k in st:
                    ok &= bool(np.array_equal(st[k], rs[k][:pop.N]))
            a, s_ = sh.sim.events()
            cross_events += int(((s_ < pop.lo) | (s_ >= pop.lo + pop.N)).sum())
        tot = np.array([cross_events], dtype=np.int64)
        import torch
        t = torch.from_numpy(tot)
        dist.all_reduce(t)
        if rank == 0:
            out.put((ok, int(t.item()), int(ref.stats()[9])))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_equals_unsharded():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    ok, cross, ever = out.get(timeout=10)
    assert ok, "sharded run diverged from the unsharded oracle"
    assert ever > 2000, ever            # a real epidemic
    assert cross > 0, "no infection crossed the shard boundary: the exchange was not exercised"
