"""GPU parity tests proper: the CUDA path, called through the C ABI, against (a) the golden vectors of the UNMODIFIED
reference and (b) the CPU oracle on the same seeded inputs.  Integer state, infection events and per-day counts are
bit-exact; float64 infection probabilities are compared with == (tolerance stated by BASELINE.json: <= 1e-12 relative;
achieved: 0 ulp)."""
import copy

import numpy as np
import pytest

from helpers import FIXTURES, Fixture, compare_outputs, compare_sims, run_against_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from dysturbd_b200 import engine
    engine.load_library()
    return engine


def test_device_philox_equals_numpy(eng):
    from oracle import philox_np
    rs = np.random.RandomState(1)
    a = rs.randint(0, 2 ** 27, 5000).astype(np.uint32)
    b = rs.randint(0, 2 ** 27, 5000).astype(np.uint32)
    for seed, day, stream in ((0, 0, 0), (42, 17, 2), (2 ** 45 + 3, 364, 1), (7, 9, 3)):
        dev = eng.keyed_uniform_device(seed, day, stream, a, b)
        assert np.array_equal(dev, philox_np.keyed_uniform(seed, day, stream, a, b))


@pytest.mark.parametrize("name", FIXTURES)
def test_engine_matches_reference_golden(name, eng):
    fx = Fixture(name)
    sim = eng.Engine(fx.pop, fx.params)
    assert run_against_fixture(sim, fx) == fx.days
    sim.close()


def test_injected_probabilities_bit_exact(eng, oracle_lib):
    """P = (((ip * risk) * pdf) * exposure) * norm_factor for every evaluated pair, CUDA vs oracle, with ==."""
    fx = Fixture("city300_dense_nolockdown")
    g = eng.Engine(fx.pop, fx.params)
    o = oracle_lib.OracleSim(fx.pop, fx.params, capture_prob=True)
    nz = 0
    for d in range(25):
        ua, ud, up = fx.dense_draws(d)
        g.step(d, ua, ud, up)
        o.step(d, ua, ud, up)
        P = g.pair_prob()
        assert np.array_equal(P, o.pair_prob), "day %d" % d
        nz += int((P > 0).sum())
    assert nz > 500
    g.close()


@pytest.mark.parametrize("name", [f for f in FIXTURES if "dense" not in f])
def test_epidemic_model_outputs_match_reference(name, eng):
    """the reference-facing host loop on the GPU == the reference's outputs dict (Stats, per-area R, Buildings, IO_mat)."""
    from dysturbd_b200.host import EpidemicModel
    fx = Fixture(name)
    m = EpidemicModel(fx.agents, fx.build, fx.visits, fx.ip, scenario=fx.scenario, params=fx.params,
                      choice_seed=fx.meta["choice_seed"], **fx.model_kw)
    out = m.run(max_days=fx.days)
    compare_outputs(out, fx.outputs, fx.days)
    chain = np.array(out["Results"]["Infections chain"])
    assert np.array_equal(chain[:, 1], fx.exp["t_inf"][fx.days - 1]) and np.array_equal(chain[:, 2], fx.exp["src_id"][fx.days - 1])


def test_checkpoint_resume_and_json_on_gpu(eng, tmp_path):
    """stop after 9 days, continue from checkpoint() in a new model on the GPU (dys_set_state), finish with run() (fused
    windows): the reference's outputs, and the JSON file create_figures.py would read"""
    import json
    from dysturbd_b200.host import EpidemicModel
    fx = Fixture("city300_philox_gradual_all")
    mk = lambda: EpidemicModel(fx.agents, fx.build, fx.visits, fx.ip, scenario=fx.scenario, params=fx.params,
                               choice_seed=fx.meta["choice_seed"])
    a = mk()
    for _ in range(9):
        a.step()
    ck = a.checkpoint()
    b = mk()
    b.restore(ck)
    out = b.run(max_days=fx.days)
    compare_outputs(out, fx.outputs, fx.days)
    path = tmp_path / "sim0.json"
    b.write_json(path)
    got = json.load(open(path))
    assert got["Results"]["Stats"] == fx.outputs["Results"]["Stats"] and got["Results"]["IO_mat"] == fx.outputs["Results"]["IO_mat"]


def _closures(pop, seed):
    rs = np.random.RandomState(seed)

    def f(day):
        o = np.ones(pop.B, np.uint8)
        if 8 <= day < 20:                         # a lockdown window: non-residential buildings, a random half
            nr = np.nonzero(pop.bld_lu >= 3)[0]
            o[rs.choice(nr, size=len(nr) // 2, replace=False)] = 0
        if day in (12, 13):                       # ... and a few residential ones (home slots get masked too)
            o[rs.choice(pop.B, size=max(1, pop.B // 50), replace=False)] = 0
        return o
    return f


@pytest.mark.parametrize("n,degree,norm,days", [(5000, 14, 6.0, 45), (30000, 14, 6.0, 40), (12000, 60, 1.5, 30),
                                                (6000, 200, 0.5, 25), (1024, 14, 8.0, 30), (1500, 14, 8.0, 30)])
def test_engine_vs_oracle_synthetic(n, degree, norm, days, eng, oracle_lib):
    """all three row-scan widths (4 / 8 / 32 lanes per row), closures (exact exposure re-check), several communities"""
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(n, seed=n % 97, degree=degree, seeds_per_22493=150)
    prm = EpiParams(norm_factor=norm, seed=1234 + n)
    g = eng.Engine(pop, prm)
    o = oracle_lib.OracleSim(pop, prm)
    compare_sims(o, g, days, open_fn=_closures(pop, n))
    assert o.stats()[9] > n // 25          # a real epidemic (ever infected), not a fizzle
    g.close()


def test_engine_vs_oracle_compliance_and_quirk(eng, oracle_lib):
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(8000, seed=3, seeds_per_22493=150)
    pop.hh_always = pop.hh_always.copy()
    pop.hh_always[[5, 77, 300]] = 1
    # agent-triggered quirk households
    trig_agents = np.arange(0, pop.N, 50)
    ptr = np.zeros(pop.N + 1, np.int64)
    ptr[trig_agents + 1] = 2
    pop.trig_ptr = np.cumsum(ptr)
    pop.trig_hh = np.random.RandomState(0).randint(0, pop.H, 2 * len(trig_agents)).astype(np.int32)
    prm = EpiParams(norm_factor=6.0, seed=99, compliance=0.6)
    g = eng.Engine(pop, prm)
    o = oracle_lib.OracleSim(pop, prm)
    compare_sims(o, g, 40)
    g.close()


def test_set_state_and_set_run_replay(eng, oracle_lib):
    """Monte-Carlo replicas reuse a loaded population: dys_set_state + dys_set_run must reproduce a fresh handle."""
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(6000, seed=11, seeds_per_22493=150)
    g = eng.Engine(pop, EpiParams(norm_factor=6.0, seed=1))
    for d in range(10):
        g.step(d)
    g.set_state(pop.status, pop.t_inf, pop.t_q, pop.t_adm, pop.src)
    g.set_run(norm_factor=4.0, seed=2)
    o = oracle_lib.OracleSim(pop, EpiParams(norm_factor=4.0, seed=2))
    compare_sims(o, g, 25)
    g.close()


def test_io_mat_on_device(eng):
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(12000, seed=2, seeds_per_22493=150)
    g = eng.Engine(pop, EpiParams(norm_factor=6.0, seed=5),
                   flags=eng.WANT_SA_HIST | eng.WANT_BUILDING_COUNTS | eng.WANT_EVENTS | eng.WANT_IO_MAT)
    tot = 0
    for d in range(20):
        g.step(d)
        a, s = g.events(d)
        m = np.zeros((pop.S, pop.S), np.int32)
        np.add.at(m, (pop.sa[a], pop.sa[s]), 1)
        assert np.array_equal(g.io_mat(d), m)
        tot += len(a)
    assert tot > 100
    g.close()


def test_error_paths(eng):
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(2048, seed=0)
    bad = copy.copy(pop)
    bad.home = pop.home.copy()
    bad.home[3] = pop.B + 5
    with pytest.raises(eng.DysError) as e:
        eng.Engine(bad, EpiParams())
    assert e.value.code == -5
    bad = copy.copy(pop)
    bad.ip_col = pop.ip_col.copy()
    bad.ip_col[:2] = bad.ip_col[:2][::-1]          # columns not ascending
    with pytest.raises(eng.DysError):
        eng.Engine(bad, EpiParams())
    g = eng.Engine(pop, EpiParams())
    g.step(0)
    with pytest.raises(eng.DysError):
        g.stats(5)                                  # not in the ring
    with pytest.raises(eng.DysError):
        g.io_mat(0)                                 # not requested
    g.close()


def test_step_many_fused_windows_vs_oracle(eng, oracle_lib):
    """dys_step_many: windows of days in one cooperative launch (k_days), open flags changing inside a window"""
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(30000, seed=8, seeds_per_22493=150)
    prm = EpiParams(norm_factor=6.0, seed=4321)
    g = eng.Engine(pop, prm)
    o = oracle_lib.OracleSim(pop, prm)
    flags = _closures(pop, 17)
    day = 0
    for n in (5, 7, 2, 6, 7, 3, 7):
        f = np.stack([flags(day + k) for k in range(n)])
        g.step_many(day, n, f)
        for k in range(n):
            o.set_open(f[k])
            o.step(day + k)
            out = g.outputs(day + k)
            # 14..17 (entries scanned / candidates / draws / successes of the push) are implementation specific inside
            # a fused window: the push of day d+1 runs before other blocks finished day d and does not read u's status
            keep = list(range(14))
            assert np.array_equal(out["stats"][keep], o.stats()[keep]), "day %d" % (day + k)
            assert np.array_equal(out["stats"][32:53], o.stats()[32:53])
            assert np.array_equal(out["sa_hist"], o.sa_hist()) and np.array_equal(out["building_counts"], o.building_counts())
            ev = out["events"]
            order = np.argsort(ev[:, 0], kind="stable")
            ea, es = o.events()
            assert np.array_equal(ev[order, 0], ea) and np.array_equal(ev[order, 1], es)
        day += n
        sg, so = g.state(), o.state()
        for key in sg:
            assert np.array_equal(sg[key], so[key]), (key, day)
    assert o.stats()[9] > 3000
    g.close()


def test_replica_sweep_is_reproducible(eng, oracle_lib):
    """BASELINE.json configs[4]: replicas of one loaded city (dys_set_state + dys_set_run) over norm_factor x compliance;
    every replica equals a fresh oracle run with the same parameters and seed."""
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(9000, seed=12, seeds_per_22493=150)
    g = eng.Engine(pop, EpiParams(), flags=0)
    sizes = {}
    for k, (norm, comp) in enumerate([(4.0, 1.0), (8.0, 1.0), (8.0, 0.4)]):
        g.set_state(pop.status, pop.t_inf, pop.t_q, pop.t_adm, pop.src)
        g.set_run(norm_factor=norm, compliance=comp, seed=100 + k)
        g.step_many(0, 7)
        g.step_many(7, 7)
        g.step_many(14, 6)
        o = oracle_lib.OracleSim(pop, EpiParams(norm_factor=norm, compliance=comp, seed=100 + k))
        for d in range(20):
            o.step(d)
        sg, so = g.state(), o.state()
        for key in sg:
            assert np.array_equal(sg[key], so[key]), (key, norm, comp)
        sizes[(norm, comp)] = int(o.stats()[9])
    assert sizes[(8.0, 1.0)] > sizes[(4.0, 1.0)]
    g.close()


def test_population_loaded_with_recovered_agents(eng, oracle_lib):
    """agents that are ALREADY recovered when the run starts (immune at day 0, or infected shortly before it) stay in the
    days-since-infection histograms, per-area histograms and R for 21 days exactly as in the reference's compute_R
    (contagious3.py:29-42) -- the sweep may only skip a recovered agent once day - t_inf >= 21 (dys_load.cuh, r_idle_day).
    Day by day through dys_step, and in week-long dys_step_many windows."""
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    import copy
    pop = copy.copy(synth_population(9000, seed=8, seeds_per_22493=150))
    rs = np.random.RandomState(2)
    sus = np.nonzero(pop.status == 2)[0]
    pick = rs.choice(sus, 600, replace=False)
    pop.status, pop.t_inf = pop.status.copy(), pop.t_inf.copy()
    pop.status[pick] = 12
    pop.t_inf[pick] = rs.choice([0, -4, -15, 6], size=len(pick))          # (6: the reference would count it as infected on day 6)
    prm = EpiParams(norm_factor=6.0, seed=77)
    days = 35
    g, o = eng.Engine(pop, prm), oracle_lib.OracleSim(pop, prm)
    ref_stats, ref_hist = [], []
    for d in range(days):
        g.step(d)
        o.step(d)
        xa, xb = o.stats(d), g.stats(d)
        assert np.array_equal(xa[:14], xb[:14]) and np.array_equal(xa[32:53], xb[32:53]), d
        assert np.array_equal(o.sa_hist(d), g.sa_hist(d)), d
        ref_stats.append(np.array(xa))
        ref_hist.append(np.array(o.sa_hist(d)))
    assert sum(int(s[32 + 1:32 + 21].sum()) for s in ref_stats[:5]) > 1000      # the loaded agents really sit in the bins
    sa, sb = o.state(), g.state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    g.close()
    w = eng.Engine(pop, prm)
    for d0 in range(0, days, 7):
        w.step_many(d0, 7)
        for d in range(d0, d0 + 7):
            x = w.stats(d)
            assert np.array_equal(x[:14], ref_stats[d][:14]) and np.array_equal(x[32:53], ref_stats[d][32:53]), d
            assert np.array_equal(w.sa_hist(d), ref_hist[d]), d
    w.close()
