"""phase H (building_lockdown, contagious3.py:57-107, 368-385) decided inside the day kernel.
Deterministic scenarios are held to the REFERENCE's golden outputs; the GRADUAL half closures (keyed draws instead of
np.random.choice) to the host twin of the same rule."""
import numpy as np
import pytest

from helpers import Fixture, compare_outputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["synth2000_philox_all", "city800_philox_edu_rel", "city300_philox_diff_all_keyfix"])
def test_device_lockdown_reproduces_the_reference(name):
    """ALL and EDU+REL close whole classes: no random choice is involved, so the device policy must reproduce the literal
    contagious3.py run -- every Stats entry (Closed buildings, R, Known R included), SAs, Buildings, IO_mat, every day.
    The DIFF fixture (per-area decisions on per-area R, dys_set_lockdown_areas) is the reference's text with its 'Vis_R' key
    repaired (oracle/literal.py)."""
    from dysturbd_b200.host import EpidemicModel
    fx = Fixture(name)
    m = EpidemicModel(fx.agents, fx.build, fx.visits, fx.ip, scenario=fx.scenario, params=fx.params, lockdown='device', **fx.model_kw)
    assert m.lockdown == 'device' and m.window == 7
    m.run(max_days=fx.days)
    compare_outputs(m.outputs, fx.outputs, fx.days)
    assert max(s['Closed buildings'] for s in m.outputs['Results']['Stats'].values()) > 0
    st = m.sim.state()
    assert np.array_equal(st["status"], fx.exp["status_x2"][fx.days - 1])
    assert np.array_equal(st["t_inf"], fx.exp["t_inf"][fx.days - 1])
    assert m.sim.h2d_bytes() == 0                 # no flag vector ever travelled
    m.sim.close()


def _city(n=40000, seed=21):
    from dysturbd_b200.synth import synth_population
    return synth_population(n, seed=seed, seeds_per_22493=120)


@pytest.mark.parametrize("scenario", [["GRADUAL", "ALL"], ["GRADUAL", "EDU", "REL"], ["ALL"], ["DIFF", "GRADUAL", "ALL"], ["DIFF", "EDU", "REL"]])
def test_device_lockdown_equals_its_host_twin(scenario):
    """the keyed rule on the device (week-long windows, no uploads) against the same rule decided on the host day by day
    (4-day windows, flags uploaded): identical Stats every day and identical final state"""
    from dysturbd_b200.host import EpidemicModel
    from dysturbd_b200.population import EpiParams
    pop = _city()
    prm = EpiParams(norm_factor=7.0, seed=99)
    kw = {"diff_key_fix": True} if "DIFF" in scenario else {}
    a = EpidemicModel(None, None, None, None, scenario=scenario, params=prm, population=pop, lockdown='device', compact=True, **kw)
    b = EpidemicModel(None, None, None, None, scenario=scenario, params=prm, population=pop, lockdown='keyed', compact=True, **kw)
    assert a.lockdown == 'device'
    a.run(max_days=120)
    b.run(max_days=120)
    sa, sb = a.outputs['Results']['Stats'], b.outputs['Results']['Stats']
    assert len(sa) == len(sb) and len(sa) > 60
    for d in sa:
        assert sa[d] == sb[d], (d, sa[d], sb[d])
    closed = [sa[d]['Closed buildings'] for d in sa]
    assert max(closed) > 0
    if 'GRADUAL' in scenario:
        assert len(set(closed)) > 2               # open, a half closure and a full closure all occurred
    xa, xb = a.sim.state(), b.sim.state()
    for k in xa:
        assert np.array_equal(xa[k], xb[k]), k
    assert a.sim.h2d_bytes() == 0 and b.sim.h2d_bytes() > 0
    a.sim.close()
    b.sim.close()


def test_device_lockdown_with_replicas():
    """every replica decides its own flags from its own R history"""
    from dysturbd_b200 import engine
    from dysturbd_b200.population import EpiParams
    pop = _city(24000, seed=4)
    R, norms, seeds = 3, [5.0, 7.0, 9.0], [1, 2, 3]
    bits = engine.SC_GRADUAL | engine.SC_ALL
    cls = np.where(pop.bld_lu >= 3, engine.CLASS_NONRES, 0).astype(np.uint8)
    g = engine.Engine(pop, EpiParams(), replicas=R, flags=0)
    g.set_replica_params(norms, None, seeds)
    g.set_lockdown(bits, cls)
    days = 70
    for d in range(0, days, 7):
        g.step_many(d, 7)
    for r in range(R):
        s = engine.Engine(pop, EpiParams(norm_factor=norms[r], seed=seeds[r]), flags=0)
        s.set_lockdown(bits, cls)
        g.select_replica(r)
        for d in range(0, days, 7):
            s.step_many(d, 7)
            for dd in range(d, d + 7):
                x, y = g.stats(dd), s.stats(dd)
                assert np.array_equal(x[:14], y[:14]) and np.array_equal(x[28:31], y[28:31]), (r, dd)
        s.close()
    g.close()
