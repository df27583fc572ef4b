#!/usr/bin/env python
"""The workload of the checked-build runs (tests/test_gpu_checked.py; compute-sanitizer where a pool allows it): a 30 k-agent
city through dys_step_many windows -- single run, two replicas in one launch, the device lockdown policy -- and, under
torchrun with two ranks, the peer exchange fused into the window.
  compute-sanitizer --tool racecheck|synccheck|memcheck python tools/sanitize_run.py
  compute-sanitizer --tool memcheck --target-processes all python -m torch.distributed.run --nproc-per-node 2 tools/sanitize_run.py --peer"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DYS_SPIN_TIMEOUT_MS", "600000")      # kernels run 10-100x slower under the sanitizer


def main():
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import COMMUNITY, synth_population
    if "--version" in sys.argv:
        print(engine.load_library().dys_version().decode())
    if "--peer" in sys.argv:
        import torch
        import torch.distributed as dist
        from dysturbd_b200.sharded import ShardedEngine
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
        cpr = 2
        pop = synth_population(COMMUNITY * cpr * world, seed=33, communities=(rank * cpr, (rank + 1) * cpr), seeds_per_22493=300, cross=0.05)
        sh = ShardedEngine(pop, EpiParams(norm_factor=6.0, seed=7), device=rank, mode="peer", flags=0)
        for d in range(0, 14, 7):
            sh.step_many(d, 7)
        s = sh.stats(13)
        if rank == 0:
            print("peer windows ok: mode %s, ever infected %d" % (sh.mode, int(s[9])))
        sh.close()
        dist.destroy_process_group()
        return
    pop = synth_population(30000, seed=7, seeds_per_22493=300)
    prm = EpiParams(norm_factor=6.0, seed=2024)
    g = engine.Engine(pop, prm)
    for d in range(0, 21, 7):
        g.step_many(d, 7)
    ever = int(g.stats(20)[9])
    g.close()
    g = engine.Engine(pop, prm, replicas=2, flags=0)
    g.set_replica_params([5.0, 7.0], [1.0, 0.5], [1, 2])
    g.set_lockdown(engine.SC_GRADUAL | engine.SC_ALL, np.where(pop.bld_lu >= 3, engine.CLASS_NONRES, 0).astype(np.uint8))
    for d in range(0, 14, 7):
        g.step_many(d, 7)
    g.sync()
    g.close()
    print("single-GPU windows ok: ever infected %d" % ever)


if __name__ == "__main__":
    main()
