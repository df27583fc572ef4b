"""The N>1 path on GPUs: two ranks, one B200 each, NCCL all-gather of the day-start infectious bytes in place in the
library's buffer (dysturbd_b200.sharded.ShardedEngine).  The sharded run must equal the single-GPU run bit for bit
(draws are keyed by global agent indices).  Skipped on a one-GPU box; the CPU twin is tests/test_sharded_cpu.py."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_COMM, DAYS = 6, 35


def _worker(rank, world, port, out, mode):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dysturbd_b200 import engine
        from dysturbd_b200.population import EpiParams
        from dysturbd_b200.sharded import ShardedEngine
        from dysturbd_b200.synth import COMMUNITY, synth_population
        n = COMMUNITY * N_COMM
        cpr = N_COMM // world
        pop = synth_population(n, seed=33, communities=(rank * cpr, (rank + 1) * cpr), seeds_per_22493=150, cross=0.05)
        prm = EpiParams(norm_factor=6.0, seed=7)
        sh = ShardedEngine(pop, prm, device=rank, mode=mode.split("-")[0])
        ref = engine.Engine(synth_population(n, seed=33, seeds_per_22493=150, cross=0.05), prm, device=rank) if rank == 0 else None
        ok, cross, why = True, 0, []
        windows = mode.endswith("-windows")      # dys_step_many: k_days with the exchange fused into the day
        sizes = [5, 7, 2, 6, 7, 3, 5]
        starts = {sum(sizes[:i]): n for i, n in enumerate(sizes)}
        for d in range(DAYS):
            if windows:
                if d in starts:
                    sh.step_many(d, starts[d])
            else:
                sh.step(d)
            g = sh.stats(d)
            st = sh.state()
            if rank == 0:
                ref.step(d)
                r = ref.stats(d)
                keep = [i for i in range(28) if i not in (14, 15) and not (windows and i in (16, 17))]
                if not (np.array_equal(g[keep], r[keep]) and np.array_equal(g[32:53], r[32:53])):
                    ok = False
                    why.append("day %d stats %s vs %s" % (d, g[:28].tolist(), r[:28].tolist()))
                rs = ref.state()
                if not windows or d + 1 in starts or d == DAYS - 1:      # (inside a window the state is already days ahead)
                    for k in st:
                        if not np.array_equal(st[k], rs[k][:pop.N]):
                            ok = False
                            why.append("day %d state %s differs for %d agents" % (d, k, int((st[k] != rs[k][:pop.N]).sum())))
                # per-area / per-building outputs only cover the owned ranges (the daily transfer must not grow with N)
                o = sh.outputs(d)
                if not np.array_equal(o["sa_hist"], ref.sa_hist(d)[pop.sa_range[0]:pop.sa_range[1]]):
                    ok = False
                    why.append("day %d sa_hist" % d)
                if not np.array_equal(o["building_counts"], ref.building_counts(d)[pop.bld_range[0]:pop.bld_range[1]]):
                    ok = False
                    why.append("day %d building_counts" % d)
            a, s_ = sh.sim.events(d)
            cross += int(((s_ < pop.lo) | (s_ >= pop.lo + pop.N)).sum())
        t = torch.tensor([cross], device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            out.put((ok, int(t.item()), int(ref.stats(DAYS - 1)[9]), why[:6]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl", "peer-windows"])
def test_two_gpu_sharded_equals_single_gpu(mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29700 + os.getpid() % 2000 + {"peer": 0, "nccl": 7, "peer-windows": 13}[mode]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, mode)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    ok, cross, ever, why = out.get(timeout=10)
    assert ok, "sharded run diverged from the single-GPU run: %s" % why
    assert ever > 3000 and cross > 0
