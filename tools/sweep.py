#!/usr/bin/env python
"""Monte-Carlo parameter sweep (BASELINE.json configs[4]): replicas of one loaded city over a grid of
norm_factor x quarantine compliance, batched into ONE launch per window (dys_params.n_replicas); replicas are
independent, so several GPUs each run their share with no communication.
usage:  python tools/sweep.py [--communities 89] [--grid 4] [--days 120]
        torchrun --nproc-per-node N tools/sweep.py ...        (the grid is dealt round-robin to the ranks)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--communities", type=int, default=89)
    ap.add_argument("--grid", type=int, default=4, help="grid x grid replicas")
    ap.add_argument("--days", type=int, default=120)
    a = ap.parse_args()
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import COMMUNITY, synth_population
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    pop = synth_population(COMMUNITY * a.communities, seed=1)
    norms = np.linspace(3.0, 9.0, a.grid)
    comps = np.linspace(0.25, 1.0, a.grid)
    jobs = [(i, j) for i in range(a.grid) for j in range(a.grid)][rank::world]
    eng = engine.Engine(pop, EpiParams(), device=local, flags=0, replicas=len(jobs))
    eng.set_replica_params([norms[i] for i, _ in jobs], [comps[j] for _, j in jobs], [1000 + i * a.grid + j for i, j in jobs])
    t0 = time.perf_counter()
    d = 0
    while d < a.days:
        n = min(7, a.days - d)
        eng.step_many(d, n)
        d += n
    eng.sync()
    dt = time.perf_counter() - t0
    out = []
    for r, (i, j) in enumerate(jobs):
        eng.select_replica(r)
        s = eng.stats(a.days - 1)
        out.append({"norm_factor": float(norms[i]), "compliance": float(comps[j]), "ever_infected": int(s[9]), "dead": int(s[7])})
    print(json.dumps({"rank": rank, "replicas": len(jobs), "seconds": dt, "agent_days_per_s": pop.N * a.days * len(jobs) / dt,
                      "results": out}))


if __name__ == "__main__":
    main()
