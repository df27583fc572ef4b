#!/usr/bin/env python
"""Developer tool: resident time per simulated day of dys_step_many windows (CUDA events on the library's stream).
usage: python tools/quick_time.py [--communities 89] [--replicas 1] [--days 140] [--warm 28] [--norm 6.0] [--window 7]"""
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
    ap.add_argument("--replicas", type=int, default=1)
    ap.add_argument("--days", type=int, default=140)
    ap.add_argument("--warm", type=int, default=28)
    ap.add_argument("--norm", type=float, default=6.0)
    ap.add_argument("--window", type=int, default=7)
    ap.add_argument("--flags", type=int, default=7)
    ap.add_argument("--seeds", type=int, default=20)
    a = ap.parse_args()
    import torch
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import COMMUNITY, synth_population
    t0 = time.perf_counter()
    import pickle
    cache = "/tmp/dys_pop_%d_%d.pkl" % (a.communities, a.seeds)
    if os.path.exists(cache):
        pop = pickle.load(open(cache, "rb"))
    else:
        pop = synth_population(COMMUNITY * a.communities, seed=1, seeds_per_22493=a.seeds)
        try:
            pickle.dump(pop, open(cache, "wb"), protocol=4)
        except Exception:
            pass
    t1 = time.perf_counter()
    eng = engine.Engine(pop, EpiParams(norm_factor=a.norm, seed=20201019), flags=a.flags, replicas=a.replicas)
    if a.replicas > 1:
        eng.set_replica_params(None, None, [20201019 + r for r in range(a.replicas)])
    eng.sync()
    t2 = time.perf_counter()
    stream = torch.cuda.ExternalStream(eng.cuda_stream())
    d = 0
    while d < a.warm:
        n = min(a.window, a.warm - d)
        eng.step_many(d, n)
        d += n
    eng.sync()
    seg = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(stream)
    marks = []
    while d < a.warm + a.days:
        n = min(a.window, a.warm + a.days - d)
        e = torch.cuda.Event(enable_timing=True)
        eng.step_many(d, n)
        e.record(stream)
        marks.append((d, n, e))
        d += n
    ev[1].record(stream)
    eng.sync()
    ms = ev[0].elapsed_time(ev[1])
    prev = ev[0]
    per = []
    for d0, n, e in marks:
        per.append((d0, prev.elapsed_time(e) / n * 1e3))
        prev = e
    st = eng.stats(a.warm + a.days - 1)
    N = pop.N * a.replicas
    print(json.dumps({"agents": pop.N, "replicas": a.replicas, "days": a.days, "us_per_day": ms / a.days * 1e3,
                      "agent_days_per_s": N * a.days / (ms * 1e-3), "gen_s": t1 - t0, "load_s": t2 - t1,
                      "ever_infected_r0": int(st[9]), "active_r0": int(st[0]),
                      "us_per_day_by_window": [round(x[1], 1) for x in per[:: max(1, len(per) // 20)]]}))
    eng.close()


if __name__ == "__main__":
    main()
