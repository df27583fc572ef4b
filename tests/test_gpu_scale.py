"""BASELINE.json's full single-GPU size (1 M agents) through size-independent properties, plus an oracle comparison at
a size the oracle still finishes in seconds per day."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_200k_agents_vs_oracle(oracle_lib):
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    from helpers import compare_sims
    pop = synth_population(200_000, seed=4, seeds_per_22493=60)
    prm = EpiParams(norm_factor=4.0, seed=77)
    g = engine.Engine(pop, prm)
    o = oracle_lib.OracleSim(pop, prm)
    compare_sims(o, g, 30)
    g.close()


def test_one_million_agents_properties():
    from dysturbd_b200 import engine
    from dysturbd_b200.engine import STAT
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    pop = synth_population(1_000_000, seed=1, seeds_per_22493=60)
    g = engine.Engine(pop, EpiParams(norm_factor=4.0, seed=5))
    ever_prev, dead_prev, rec_prev = 0, 0, 0
    events_total = int((pop.status == 4).sum())
    for d in range(60):
        g.step(d)
        s = g.stats(d)
        st = g.state()
        cnt = np.bincount(st["status"], minlength=16)
        # compartments partition the population
        assert cnt[[2, 4, 6, 7, 8, 10, 12, 14]].sum() == pop.N
        assert s[STAT["active"]] == cnt[4] + cnt[7] + cnt[8] + cnt[10]
        assert s[STAT["recovered"]] == cnt[12] and s[STAT["dead"]] == cnt[14] and s[STAT["hosp"]] == cnt[10]
        assert s[STAT["quarantined"]] == cnt[6] + cnt[7] + cnt[8]
        ever = int(s[STAT["ever_infected"]])
        events_total += int(s[STAT["n_events"]])
        assert ever == events_total                        # every infection is an event, nobody is infected twice
        assert ever >= ever_prev and s[STAT["dead"]] >= dead_prev and s[STAT["recovered"]] >= rec_prev
        assert s[STAT["dead"]] == dead_prev + s[STAT["daily_deaths"]]
        ever_prev, dead_prev, rec_prev = ever, int(s[STAT["dead"]]), int(s[STAT["recovered"]])
        # histogram: bin k counts agents infected k days ago
        h = s[32:53]
        tinf = st["t_inf"][(st["status"] != 2) & (st["status"] != 6)]
        assert np.array_equal(h, np.bincount(d - tinf, minlength=21)[:21])
        assert g.sa_hist(d).sum(axis=0).tolist() == h.tolist()
        bc = g.building_counts(d)
        assert bc.sum() == cnt[4] + cnt[7] + cnt[8]
        a, src = g.events(d)
        assert len(a) == s[STAT["n_events"]] and (st["t_inf"][a] == d).all() and np.array_equal(st["src"][a], src)
        # an infector was infectious this morning: infected at least one day ago
        assert (st["t_inf"][src] < d).all() if len(a) else True
    assert ever_prev > 20_000
    g.close()


def test_fused_windows_equal_single_days_at_scale():
    """The product path (dys_step_many: k_days, one grid barrier per day, blocks pushing their own frontier, the daily
    equal-work cut over more than one staging chunk of pieces) against the two-launch path (dys_step, itself checked
    against the oracle above) on a population larger than BASELINE.json's: state, counts, histograms, per-area and
    per-building outputs and the set of infection events must be identical on every day."""
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import COMMUNITY, synth_population
    pop = synth_population(COMMUNITY * 107, seed=3, seeds_per_22493=120)      # 1.2 M agents: 9 416 pieces of 128
    prm = EpiParams(norm_factor=8.0, seed=11)
    a = engine.Engine(pop, prm)      # windows
    b = engine.Engine(pop, prm)      # single days
    day = 0
    for n in (7, 7, 3, 7, 5, 7, 7, 2, 7, 7):
        a.step_many(day, n)
        for k in range(n):
            d = day + k
            b.step(d)
            oa, sb = a.outputs(d), b.stats(d)
            assert np.array_equal(oa["stats"][:14], sb[:14]), "day %d: %s vs %s" % (d, oa["stats"][:14], sb[:14])
            assert np.array_equal(oa["stats"][18:28], sb[18:28]) and np.array_equal(oa["stats"][32:53], sb[32:53]), "day %d" % d
            assert np.array_equal(oa["sa_hist"], b.sa_hist(d)) and np.array_equal(oa["building_counts"], b.building_counts(d)), "day %d" % d
            ea, eb = a.events(d), b.events(d)
            assert np.array_equal(ea[0], eb[0]) and np.array_equal(ea[1], eb[1]), "events day %d" % d
        sa, sb = a.state(), b.state()
        for key in sa:
            assert np.array_equal(sa[key], sb[key]), "%s after day %d" % (key, day + n - 1)
        day += n
    assert int(b.stats(day - 1)[9]) > 50_000      # a real epidemic: blocks settle and push unevenly
    a.close()
    b.close()


def test_one_million_agents_365_days_vs_oracle(oracle_lib):
    """BASELINE.json configs[1] at full size on the PRODUCT path: the 1 M-agent city, 365 days in week-long dys_step_many
    windows (k_days), against the C oracle -- the integer stats and histograms of every day, the whole state after every
    window."""
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import COMMUNITY, synth_population
    pop = synth_population(COMMUNITY * 89, seed=1, seeds_per_22493=20)
    prm = EpiParams(norm_factor=6.0, seed=20201019)
    g = engine.Engine(pop, prm)
    o = oracle_lib.OracleSim(pop, prm)
    day = 0
    while day < 365:
        n = min(7, 365 - day)
        g.step_many(day, n)
        for d in range(day, day + n):
            so = o.step(d)
            out = g.outputs(d)
            assert np.array_equal(out["stats"][:14], so[:14]), "day %d: %s vs %s" % (d, out["stats"][:14], so[:14])
            assert np.array_equal(out["stats"][32:53], so[32:53]), "histogram day %d" % d
            assert np.array_equal(out["sa_hist"], o.sa_hist()) and np.array_equal(out["building_counts"], o.building_counts()), d
        day += n
        sg, so_ = g.state(), o.state()
        for k in sg:
            assert np.array_equal(sg[k], so_[k]), "%s after day %d" % (k, day - 1)
    assert int(g.stats(364)[9]) > 300_000
    g.close()


def test_ten_million_agents_lockdown_and_compliance_vs_oracle(oracle_lib):
    """BASELINE.json configs[2] at full size: the 10 M-agent metro under ['GRADUAL', 'ALL'] with quarantine compliance 0.7,
    30 days through the documented host loop (EpidemicModel: 4-day windows, flags decided on the host every day) -- the CUDA
    engine against the same loop over the C oracle: every day's Stats and the final state."""
    from dysturbd_b200.host import EpidemicModel
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import COMMUNITY, synth_population
    pop = synth_population(COMMUNITY * 888, seed=2, seeds_per_22493=400)      # many seeds: R crosses 1 and 2 within a month
    prm = EpiParams(norm_factor=9.0, compliance=0.7, seed=4242)
    sc = ["GRADUAL", "ALL"]
    a = EpidemicModel(None, None, None, None, scenario=sc, params=prm, population=pop, compact=True, choice_seed=5)
    b = EpidemicModel(None, None, None, None, scenario=sc, params=prm, population=pop, compact=True, choice_seed=5,
                      backend=oracle_lib.OracleSim)
    a.run(max_days=30, finalize=False)
    b.run(max_days=30, finalize=False)
    sa, sb = a.outputs["Results"]["Stats"], b.outputs["Results"]["Stats"]
    assert len(sa) == len(sb) == 30
    for d in range(30):
        assert sa[d] == sb[d], (d, sa[d], sb[d])
        assert np.array_equal(a.outputs["Results"]["SAs"][d]["R"], b.outputs["Results"]["SAs"][d]["R"]), d
        assert np.array_equal(a.outputs["Results"]["Buildings"][d], b.outputs["Results"]["Buildings"][d]), d
    assert max(s["Closed buildings"] for s in sa.values()) > 0 and sa[29]["Quarantined"] > 0
    xa, xb = a.sim.state(), b.sim.state()
    for k in xa:
        assert np.array_equal(xa[k], xb[k]), k
    a.sim.close()
