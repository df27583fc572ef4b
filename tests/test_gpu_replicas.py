"""Replicas batched in one launch (BASELINE.json configs[4]), agents loaded as recovered, and the abort
path of the spinning loops -- all through the C ABI, against the CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pop(n=30000, seed=5):
    from dysturbd_b200.synth import synth_population
    return synth_population(n, seed=seed, seeds_per_22493=200)


def test_replicas_in_one_launch_equal_fresh_oracle_runs(oracle_lib):
    """R = 6 replicas of one city advance in ONE k_days launch per window; every replica must equal a fresh oracle run
    with that replica's (norm_factor, compliance, seed): stats and histograms every day, per-area / per-building outputs,
    the set of infection events, and the state after every window."""
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    pop = _pop()
    R = 6
    norms = [3.0, 4.0, 5.0, 6.0, 7.0, 8.0]
    comps = [1.0, 0.9, 0.7, 0.5, 0.25, 1.0]
    seeds = [101, 102, 103, 104, 105, 101]
    g = engine.Engine(pop, EpiParams(), replicas=R)
    g.set_replica_params(norms, comps, seeds)
    oracles = [oracle_lib.OracleSim(pop, EpiParams(norm_factor=norms[r], compliance=comps[r], seed=seeds[r])) for r in range(R)]
    day = 0
    for n in (7, 1, 7, 3, 7, 7, 2, 6):
        g.step_many(day, n)
        for r in range(R):
            g.select_replica(r)
            o = oracles[r]
            for d in range(day, day + n):
                o.step(d)
                out = g.outputs(d)
                so = o.stats()
                keep = [i for i in range(14)]
                assert np.array_equal(out["stats"][keep], so[keep]), "replica %d day %d: %s vs %s" % (r, d, out["stats"][:14], so[:14])
                assert np.array_equal(out["stats"][32:53], so[32:53]), "histogram replica %d day %d" % (r, d)
                assert np.array_equal(out["sa_hist"], o.sa_hist()), "sa_hist replica %d day %d" % (r, d)
                assert np.array_equal(out["building_counts"], o.building_counts()), "bcount replica %d day %d" % (r, d)
                ea, es = g.events(d)
                oa, os_ = o.events()
                assert np.array_equal(ea, oa) and np.array_equal(es, os_), "events replica %d day %d" % (r, d)
            sg, so_ = g.state(), o.state()
            for k in sg:
                assert np.array_equal(sg[k], so_[k]), "%s replica %d after day %d" % (k, r, day + n - 1)
        day += n
    finals = [int(o.stats()[9]) for o in oracles]
    assert max(finals) > 3000 and len(set(finals)) > 3      # real, different epidemics
    g.close()


def test_replica_set_state_and_rerun(oracle_lib):
    """set_state on all replicas restarts the batch; a single selected replica can be reloaded on its own"""
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    pop = _pop(12000, seed=9)
    g = engine.Engine(pop, EpiParams(norm_factor=6.0), replicas=3, flags=0)
    g.set_replica_params([6.0, 6.0, 6.0], None, [1, 2, 3])
    g.step_many(0, 7)
    g.step_many(7, 7)
    first = []
    for r in range(3):
        g.select_replica(r)
        first.append(g.stats(13).copy())
    g.select_replica(-1)
    init = pop.copy_state()
    g.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
    g.step_many(0, 7)
    g.step_many(7, 7)
    for r in range(3):
        g.select_replica(r)
        assert np.array_equal(g.stats(13), first[r])
    o = oracle_lib.OracleSim(pop, EpiParams(norm_factor=6.0, seed=2))
    for d in range(14):
        o.step(d)
    assert np.array_equal(first[1][:14], o.stats()[:14])
    g.close()


@pytest.mark.parametrize("windows", [False, True])
def test_agents_loaded_as_recovered_stay_in_the_histograms(oracle_lib, windows):
    """a population may be LOADED with recovered agents whose infection is recent (e.g. seeded immunity).  The
    reference counts them in the days-since-infection histograms (compute_R, contagious3.py:29-42) for 21 days; the sweep
    may only skip a recovered agent once it has left the histogram."""
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    from helpers import compare_sims
    pop = _pop(20000, seed=3)
    rng = np.random.default_rng(0)
    rec = rng.choice(pop.N, 1500, replace=False)
    rec = rec[pop.status[rec] == 2]
    pop.status[rec] = 12
    pop.t_inf[rec] = rng.integers(-15, 1, len(rec))      # infected up to 15 days before the run starts
    pop.src[rec] = -1
    prm = EpiParams(norm_factor=6.0, seed=17)
    g = engine.Engine(pop, prm)
    o = oracle_lib.OracleSim(pop, prm)
    if not windows:
        compare_sims(o, g, 26)
    else:
        day = 0
        for n in (7, 7, 7, 5):
            g.step_many(day, n)
            for d in range(day, day + n):
                o.step(d)
                out = g.outputs(d)
                assert np.array_equal(out["stats"][:14], o.stats()[:14]), d
                assert np.array_equal(out["stats"][32:53], o.stats()[32:53]), "histogram day %d" % d
                assert np.array_equal(out["sa_hist"], o.sa_hist()), "sa_hist day %d" % d
            day += n
    g.close()


def test_abort_and_barrier_timeout():
    """dys_abort ends the handle's launches with DYS_ERR_ABORTED until the state is reloaded; a CTA that never reaches
    the grid barrier makes the others time out (DYS_SPIN_TIMEOUT_MS) instead of hanging the GPU."""
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    pop = _pop(12000, seed=2)
    init = pop.copy_state()
    g = engine.Engine(pop, EpiParams(norm_factor=6.0, seed=1), flags=0)
    g.step_many(0, 7)
    ref = g.stats(6).copy()
    g.abort()
    with pytest.raises(engine.DysError) as e:
        g.step_many(7, 7)
    assert e.value.code == -7
    g.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
    g.step_many(0, 7)
    assert np.array_equal(g.stats(6), ref)
    g.close()
    os.environ["DYS_SPIN_TIMEOUT_MS"] = "300"
    os.environ["DYS_TEST_HANG_BLOCK"] = "3"
    try:
        g = engine.Engine(pop, EpiParams(norm_factor=6.0, seed=1), flags=0)
        g.step_many(0, 7)
        with pytest.raises(engine.DysError) as e:
            g.sync()
        assert e.value.code == -7 and "timed out" in str(e.value)
    finally:
        del os.environ["DYS_SPIN_TIMEOUT_MS"], os.environ["DYS_TEST_HANG_BLOCK"]
    g.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
    g.step_many(0, 7)
    assert np.array_equal(g.stats(6), ref)
    g.close()
