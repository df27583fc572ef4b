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


def _network_worker(rank, world, port, out):
    """rank r builds the rows [lo, hi) of interaction_prob from the whole graph (rows are independent: no exchange on the data
    path); the pieces are gathered over gloo only to be compared with the reference's matrix"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dysturbd_b200 import network
        from oracle import port as oracle_port
        from helpers import GOLDEN
        z = np.load(os.path.join(GOLDEN, "network_city300.npz"))
        n = len(z["agents"])
        order = np.lexsort((z["edge_col"], z["edge_row"]))
        indptr = np.concatenate([[0], np.cumsum(np.bincount(z["edge_row"], minlength=n))]).astype(np.int64)
        idx, w = z["edge_col"][order].astype(np.int32), z["edge_w"][order]

        def paths_csr(indptr, idx, w, src, capacity=None, skip_self=True, device=0, out=None, max_dist=np.inf, col_keep=None):
            d = oracle_port.shortest_paths(indptr, idx, w, src)                    # (the C oracle stands in for the device call)
            d[np.arange(len(src)), src] = np.inf
            r, c = np.nonzero(np.isfinite(d))
            col, dist_, off = out
            col[off:off + len(r)], dist_[off:off + len(r)] = c, d[r, c]
            return np.concatenate([[0], np.cumsum(np.bincount(r, minlength=len(src)))]), col[off:off + len(r)], dist_[off:off + len(r)]
        network.shortest_paths_csr = paths_csr
        lo, hi = rank * n // world, (rank + 1) * n // world
        part = network.interaction_prob(indptr, idx, w, device=rank, row_range=(lo, hi), rows_per_call=150)
        parts = [None] * world
        dist.all_gather_object(parts, (part.indptr, part.indices, part.data))
        if rank == 0:
            ptr = np.concatenate([[0]] + [p[0][1:] + off for p, off in zip(parts, np.cumsum([0] + [p[0][-1] for p in parts[:-1]]))])
            ok = np.array_equal(ptr, z["ip_indptr"]) and np.array_equal(np.concatenate([p[1] for p in parts]), z["ip_indices"])
            data = np.concatenate([p[2] for p in parts])
            ok &= bool(np.all(np.abs(data - z["ip_data"]) <= np.spacing(z["ip_data"])))
            out.put((bool(ok), int(ptr[-1])))
    finally:
        dist.destroy_process_group()


def test_two_rank_interaction_prob_by_row_ranges():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_network_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    ok, nnz = out.get(timeout=10)
    assert ok, "the row ranges of two ranks do not add up to the reference's interaction_prob"
    assert nnz == 9308
