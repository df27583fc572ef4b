"""Reference side of the distribution test (BASELINE.json: "epidemic curves across independent seeds must agree in
distribution, KS test at p > 0.05").  Runs the UNMODIFIED contagious3.py loop 64 times on the population of the
city300 golden fixture, each time with its own MT19937 stream (the reference's own generator), and stores three
epidemic-curve summaries per run.  Needs /root/reference:   python tests/golden/make_ks_reference.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import literal  # noqa: E402
from helpers import Fixture  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RUNS, DAYS = 64, 80

if __name__ == "__main__":
    fx = Fixture("city300_dense_nolockdown")
    setup = dict(agents=fx.agents, build=fx.build, bld_visits_by_agents=fx.visits, interaction_prob=fx.ip.toarray())
    n = len(fx.agents)
    out = np.zeros((RUNS, 3), dtype=np.int64)       # final size, peak active, peak day
    for k in range(RUNS):
        r = literal.reference_days(setup, ["noLockdown"], DAYS, literal.DenseDraws(n, 1000 + k),
                                   overrides={"norm_factor": fx.meta["norm_factor"]})
        st = r["outputs"]["Results"]["Stats"]
        active = np.array([st[d]["Active infected"] for d in sorted(st)])
        out[k] = (st[max(st)]["Total infected"], active.max(), int(active.argmax()))
        print(k, out[k], flush=True)
    np.savez_compressed(os.path.join(HERE, "ks_city300_reference.npz"), summaries=out, days=DAYS, runs=RUNS,
                        norm_factor=fx.meta["norm_factor"])
