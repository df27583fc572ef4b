"""Distribution-level parity (BASELINE.json: KS test, p > 0.05): 64 epidemics of the UNMODIFIED reference loop, each on
its own MT19937 stream (tests/golden/make_ks_reference.py -> ks_city300_reference.npz), against 64 epidemics on keyed
Philox streams -- from the C oracle (CPU) and from the CUDA engine (GPU) -- on the same population and parameters.
Summaries: final size, peak of active infections, day of the peak."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, Fixture

REF = os.path.join(GOLDEN, "ks_city300_reference.npz")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="ks_city300_reference.npz not generated")
P_MIN = 0.05


def _summaries(make_sim, fx, runs, days):
    out = np.zeros((runs, 3), dtype=np.int64)
    for k in range(runs):
        sim = make_sim(5000 + k)
        active, total = [], 0
        for d in range(days):
            sim.step(d)
            s = sim.stats(d)
            active.append(int(s[0]))
            total = int(s[9])
            if active[-1] == 0:            # the reference's loop condition (contagious3.py:176)
                break
        if hasattr(sim, "close"):
            sim.close()
        active = np.array(active)
        out[k] = (total, active.max(), int(active.argmax()))
    return out


def _check(ours, ref):
    from scipy.stats import ks_2samp
    for j, name in enumerate(("final size", "peak active", "peak day")):
        p = ks_2samp(ours[:, j], ref[:, j]).pvalue
        assert p > P_MIN, "%s: KS p = %.4f (ours mean %.1f, reference mean %.1f)" % (name, p, ours[:, j].mean(), ref[:, j].mean())


@needs_ref
def test_ks_oracle_vs_reference(oracle_lib):
    from dysturbd_b200.population import EpiParams
    z = np.load(REF)
    fx = Fixture("city300_dense_nolockdown")
    mk = lambda seed: oracle_lib.OracleSim(fx.pop, EpiParams(norm_factor=float(z["norm_factor"]), seed=seed))
    _check(_summaries(mk, fx, int(z["runs"]), int(z["days"])), z["summaries"])


@needs_ref
@pytest.mark.gpu
def test_ks_engine_vs_reference():
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    z = np.load(REF)
    fx = Fixture("city300_dense_nolockdown")
    mk = lambda seed: engine.Engine(fx.pop, EpiParams(norm_factor=float(z["norm_factor"]), seed=seed))
    _check(_summaries(mk, fx, int(z["runs"]), int(z["days"])), z["summaries"])
