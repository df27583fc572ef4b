"""BASELINE.json configs[0]: the repo's shipped city (22 493 agents, 1 015 buildings, interaction_prob with 162.5 M
non-zeros) against the UNMODIFIED contagious3.py loop run for 100 days under keyed Philox draws
(oracle/make_config1_ref.py -> tests/golden/config1/, committed).  Bit-exact agent state and Stats."""
import numpy as np
import pytest

from helpers import Config1, config1_available, config1_ip_cached



def test_config1_fixture_is_committed():
    """the only full-size pin of the reference: its absence is a failure, not a skip"""
    assert config1_available(), "tests/golden/config1/config1_*.npz missing (oracle/make_config1_ref.py generates them)"


@pytest.mark.skipif(not config1_ip_cached(), reason="interaction_prob cache absent (rebuilding it takes minutes; the GPU test does)")
def test_oracle_matches_reference_on_shipped_city(oracle_lib):
    c1 = Config1(allow_recompute=False)
    sim = oracle_lib.OracleSim(c1.pop, c1.params)
    for d in range(12):
        sim.step(d)
        c1.check_day(sim, d)


@pytest.mark.gpu
def test_engine_matches_reference_on_shipped_city_100_days():
    from dysturbd_b200 import engine
    from dysturbd_b200.host import EpidemicModel
    c1 = Config1()
    m = EpidemicModel(None, None, None, None, scenario=["noLockdown"], params=c1.params, population=c1.pop)
    for d in range(c1.days):
        m.step()
        c1.check_day(m.sim, d)
        got, exp = m.outputs["Results"]["Stats"][d], c1.outputs["Results"]["Stats"][str(d)]
        for k in ("R", "Known R", "Total infected", "Susceptible", "Closed buildings"):
            assert got[k] == exp[k], (d, k, got[k], exp[k])
        assert np.array_equal(m.open.astype(np.uint8), c1.exp["open_next"][d])
    assert m.outputs["Results"]["Stats"][c1.days - 1]["Total infected"] > 5000
    m.sim.close()
