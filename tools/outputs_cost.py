#!/usr/bin/env python
"""Developer tool: resident days per second with and without the per-day output block (per-area histograms, per-building
counts, events): separates the kernels from the device->host copy stream."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dysturbd_b200 import engine
from dysturbd_b200.population import EpiParams
from dysturbd_b200.synth import COMMUNITY, synth_population

pop = synth_population(COMMUNITY * 89, seed=1, seeds_per_22493=20)
prm = EpiParams(norm_factor=6.0, seed=20201019)
init = pop.copy_state()
for name, flags in (("all outputs", engine.WANT_SA_HIST | engine.WANT_BUILDING_COUNTS | engine.WANT_EVENTS), ("stats only", 0)):
    e = engine.Engine(pop, prm, flags=flags)
    for rep in range(2):
        e.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
        e.sync()
        t0 = time.perf_counter()
        d = 0
        while d < 368:
            n = min(7, 368 - d)
            e.step_many(d, n)
            d += n
        t1 = time.perf_counter()
        e.sync()
        t2 = time.perf_counter()
    print("%-12s %.2f us/day (host queueing %.2f us/day)" % (name, (t2 - t0) / 368 * 1e6, (t1 - t0) / 368 * 1e6))
    e.close()
