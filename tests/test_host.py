"""Host-side logic: reference matrices <-> structure of arrays, the whole-row-isin quirk sets, R from histograms,
the lockdown decision table (contagious3.py:57-107), the synthetic generator."""
import numpy as np
import pytest

from dysturbd_b200 import host
from dysturbd_b200.population import EpiParams, Population, default_pdf_table, quirk_sets
from dysturbd_b200.synth import COMMUNITY, Layout, synth_population, to_reference


def test_pdf_table_matches_scipy_gamma():
    from scipy.stats import gamma
    g = gamma((4.5 / 3.5) ** 2, scale=3.5 ** 2 / 4.5)
    t = default_pdf_table()
    assert t[0] == 0.0
    for i in (1, 7, 20, 28):
        assert t[i] == g.pdf(i)
    assert float.hex(float(t[1])) == "0x1.2cd32ff619366p-3"          # SURVEY.md section 7 probe


def test_round_trip_synth_reference_format():
    pop = synth_population(3000, seed=3)
    agents, build, visits, ip = to_reference(pop)
    back = Population.from_reference(agents, build, visits, ip)
    for k in ("status", "t_inf", "t_q", "t_adm", "src", "hh", "home", "sa", "routine", "ip_rowptr", "ip_col", "ip_val",
              "risk", "p_adm", "p_death", "open"):
        assert np.array_equal(getattr(pop, k), getattr(back, k)), k
    assert back.hh_always.sum() == 0 and len(back.trig_hh) == 0


def test_quirk_sets():
    agents = np.zeros((6, 29))
    agents[:, 0] = 5e6 + np.arange(6)
    agents[:, 1] = [0, 0, 200001, 200001, 200002, 200003]
    agents[:, 7] = [200002, 7e5, 7e5, 7e5, 7e5, 7e5]     # agent 0's home building id collides with household 200002
    hh_ids, hh = np.unique(agents[:, 1], return_inverse=True)
    always, ptr, trig = quirk_sets(agents, hh_ids, hh.astype(np.int32))
    assert list(always) == [1, 0, 0, 0]
    assert list(np.diff(ptr)) == [1, 0, 0, 0, 0, 0] and list(trig) == [2]
    agents[5, 1] = 12.0
    with pytest.raises(ValueError):
        quirk_sets(agents, *[x if i == 0 else x.astype(np.int32) for i, x in enumerate(np.unique(agents[:, 1], return_inverse=True))])


def test_compute_R_from_hist():
    pdf = [float(x) for x in default_pdf_table()]
    hist = np.zeros(21, dtype=np.int64)
    assert host.compute_R_from_hist(hist, pdf, 21) == 0
    hist[0], hist[3], hist[9] = 5, 10, 2
    s = 0.
    s += 10 * pdf[3]
    s += 2 * pdf[9]
    assert host.compute_R_from_hist(hist, pdf, 21) == 5 / s


def test_lockdown_decision_table():
    lu = np.array([1, 1, 3, 4, 4, 3, 4, 4.])
    usg = np.array([0, 0, 0, 5310, 5501, 0, 5312, 5521.])
    cur = np.ones(8)
    ch = np.random.RandomState(0).choice
    L = lambda sc, v, p, c=cur: host.building_lockdown(lu, usg, c, sc, v, p, ch)
    assert L(['noLockdown'], 5, 0).all()
    assert list(L(['ALL'], 1.5, 0)) == [1, 1, 0, 0, 0, 0, 0, 0]
    assert L(['ALL'], 1.0, 0).all()
    assert list(L(['EDU'], 3, 0)) == [1, 1, 1, 0, 1, 1, 0, 1]
    assert list(L(['EDU', 'REL'], 1.2, 0)) == [1, 1, 1, 0, 0, 1, 0, 0]
    assert list(L(['GRADUAL', 'ALL'], 2.0, 0)) == [1, 1, 0, 0, 0, 0, 0, 0]
    g = L(['GRADUAL', 'ALL'], 1.5, 0.5)                      # band just entered: a random half of the class
    assert g[:2].all() and (g[2:] == 0).sum() == 3
    keep = np.array([1, 1, 0, 1, 0, 1, 1, 1.])
    assert np.array_equal(L(['GRADUAL', 'ALL'], 1.5, 1.4, keep), keep)   # staying inside the band: keep yesterday's
    assert L(['GRADUAL', 'REL'], 0.9, 1.5).all()


def test_synth_layout_and_shards_agree():
    n = COMMUNITY * 3 + 5000
    full = synth_population(n, seed=9)
    L = Layout(n, 9)
    assert full.N == n and L.C == 4 and full.H == L.H
    assert (np.diff(full.ip_rowptr) > 0).all()
    assert (np.diff(full.home) >= 0).all() or True
    # a shard generated on its own == the same rows of the full population
    part = synth_population(n, seed=9, communities=(1, 3))
    lo, hi = part.lo, part.lo + part.N
    assert (lo, hi) == (COMMUNITY, 3 * COMMUNITY)
    for k in ("status", "hh", "home", "sa", "risk", "p_adm", "p_death"):
        assert np.array_equal(getattr(part, k), getattr(full, k)[lo:hi]), k
    h0, h1 = part.halo
    assert np.array_equal(part.routine, full.routine[h0:h1])
    a, b = full.ip_rowptr[lo], full.ip_rowptr[hi]
    assert np.array_equal(part.ip_col, full.ip_col[a:b]) and np.array_equal(part.ip_val, full.ip_val[a:b])
    assert np.array_equal(part.ip_rowptr, full.ip_rowptr[lo:hi + 1] - a)
    assert part.ip_col.min() >= h0 and part.ip_col.max() < h1
    # symmetric network, household cliques at 1.0
    import scipy.sparse as sp
    m = sp.csr_matrix((full.ip_val, full.ip_col, full.ip_rowptr), shape=(n, n))
    assert abs(m - m.T).max() == 0
    same = full.hh[np.repeat(np.arange(n), np.diff(full.ip_rowptr))] == full.hh[full.ip_col]
    assert (full.ip_val[same] == 1.0).all()


def test_window_length_follows_the_lockdown_lag():
    """EpidemicModel queues days in windows it can decide the open flags for: Known R lags `diagnosis` days
    (contagious3.py:306-309), so two windows of (diagnosis + 1) // 2 days never need an unconsumed day; a scenario that
    closes nothing has constant flags and takes week-long windows."""
    from dysturbd_b200.host import EpidemicModel

    class Dummy:                                   # no device: only the constructor logic is exercised
        def __init__(self, pop, params):
            pass
    pop = synth_population(2000, seed=2)
    agents, build, visits, ip = to_reference(pop)
    mk = lambda sc, **kw: EpidemicModel(agents, build, visits, ip, scenario=sc, params=EpiParams(**kw), backend=Dummy, population=pop)
    assert mk(['noLockdown']).window == 7
    assert mk(['GRADUAL', 'ALL']).window == 4
    assert mk(['EDU', 'REL']).window == 4
    assert mk(['ALL'], diagnosis=3).window == 2
    assert mk(['ALL'], diagnosis=1).window == 1


def test_windowed_run_equals_daily_run(oracle_lib):
    """EpidemicModel.run() queues whole windows of days ahead of the day it consumes (two windows in flight, resumable
    with finalize=False).  Over a backend that offers step_many (here the C oracle behind a window adapter) it must produce
    exactly what the day-by-day loop produces -- for a scenario that closes nothing (week-long windows: the flags must not
    wait for R) and for a lockdown scenario (4-day windows)."""
    from dysturbd_b200.host import EpidemicModel

    class Windowed(oracle_lib.OracleSim):
        def step_many(self, first, n, flags=None):
            self._out = getattr(self, "_out", {})
            for k in range(n):
                if flags is not None:
                    self.set_open(flags[k])
                self.step(first + k)
                a, s = self.events()
                ev = np.stack([a, s], 1) if len(a) else np.zeros((0, 2), np.int32)
                self._out[first + k] = dict(stats=self.stats(), sa_hist=self.sa_hist(), building_counts=self.building_counts(),
                                            events=ev, n_events=len(a))

        def outputs(self, d):
            return self._out[d]

    pop = synth_population(COMMUNITY * 2, seed=2, seeds_per_22493=400)
    prm = EpiParams(norm_factor=6.0, seed=1)
    for sc, win in ((["noLockdown"], 7), (["GRADUAL", "ALL"], 4)):
        a = EpidemicModel(None, None, None, None, scenario=sc, params=prm, population=pop, compact=True, choice_seed=5, backend=Windowed)
        assert a.window == win
        a.run(max_days=5, finalize=False)
        a.run(max_days=40, finalize=False)
        b = EpidemicModel(None, None, None, None, scenario=sc, params=prm, population=pop, compact=True, choice_seed=5,
                          backend=oracle_lib.OracleSim)
        b.run(max_days=40, finalize=False)
        assert a.outputs["Results"]["Stats"] == b.outputs["Results"]["Stats"], sc
        assert a.day == b.day == 40


def test_building_lockdown_cache_changes_nothing():
    """building_lockdown(cache=...) hands out the all-open / fully closed vectors it formed once: for every scenario and
    every (vR, prevVR) around the thresholds of contagious3.py:57-107 the result equals the uncached evaluation, and the
    np.random.choice stream is consumed identically (half closures are never cached)"""
    from dysturbd_b200.host import building_classes, building_lockdown
    rs = np.random.RandomState(4)
    B = 400
    lu = rs.randint(1, 6, B).astype(float)
    usg = rs.choice([5310.0, 5501.0, 5521.0, 5340.0, 1000.0, 2000.0], B)
    classes = building_classes(lu, usg)
    values = [0, 0.5, 1.0, 1.0000001, 1.5, 1.9999, 2.0, 2.5, float("nan")]
    for sc in (['ALL'], ['EDU', 'REL'], ['GRADUAL', 'ALL'], ['GRADUAL', 'EDU'], ['GRADUAL', 'EDU', 'REL'], ['noLockdown']):
        cache = {}
        a_rng, b_rng = np.random.RandomState(9), np.random.RandomState(9)
        cur_a = cur_b = np.ones(B, dtype=np.uint8)
        for vR in values:
            for pv in values:
                x = building_lockdown(lu, usg, cur_a, sc, vR, pv, a_rng.choice, classes=classes, dtype=np.uint8)
                y = building_lockdown(lu, usg, cur_b, sc, vR, pv, b_rng.choice, classes=classes, dtype=np.uint8, cache=cache)
                assert np.array_equal(x, y), (sc, vR, pv)
                cur_a, cur_b = x, y
        assert a_rng.randint(1 << 30) == b_rng.randint(1 << 30), sc          # the two choice streams are in the same place
