"""The C restatement (oracle/dys_oracle.c) against the committed golden vectors produced by the UNMODIFIED
reference loop (tests/golden/make_golden.py): bit-exact agent state and Stats every day; and the host-side output
builder (dysturbd_b200/host.py) driven by the oracle reproduces the reference's whole outputs dict, floats included."""
import numpy as np
import pytest

from helpers import FIXTURES, Fixture, compare_outputs, run_against_fixture


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_golden(name, oracle_lib):
    fx = Fixture(name)
    sim = oracle_lib.OracleSim(fx.pop, fx.params)
    assert run_against_fixture(sim, fx) == fx.days
    assert fx.outputs["Results"]["Stats"][str(fx.days - 1)]["Total infected"] > 100   # a real epidemic, not a fizzle


@pytest.mark.parametrize("name", [f for f in FIXTURES if "dense" not in f])
def test_host_outputs_match_reference(name, oracle_lib):
    """EpidemicModel with the oracle as backend (tests only) == the reference's outputs dict, including the lockdown
    decisions it takes itself (same np.random.choice stream as oracle/literal.py)."""
    from dysturbd_b200.host import EpidemicModel
    fx = Fixture(name)
    m = EpidemicModel(fx.agents, fx.build, fx.visits, fx.ip, scenario=fx.scenario, params=fx.params,
                      choice_seed=fx.meta["choice_seed"], backend=oracle_lib.OracleSim)
    out = m.run(max_days=fx.days)
    assert m.day == fx.days
    compare_outputs(out, fx.outputs, fx.days)
    # lockdown flags the host derived == the reference's
    assert np.array_equal(m.open.astype(np.uint8), fx.exp["open_next"][fx.days - 1])


def test_quirk_household_zero_fires(oracle_lib):
    """contagious3.py:267: household id 0 matches the zero columns of every diagnosed row -> its status-2 members are
    quarantined whenever ANYONE is diagnosed.  The synthetic fixture renames a 5-person household to id 0; switching the quirk off must change
    the trajectory, i.e. the golden vectors really pin it."""
    fx = Fixture("synth2000_philox_all")
    assert fx.pop.hh_always.sum() == 1
    import copy
    pop = copy.copy(fx.pop)
    pop.hh_always = np.zeros_like(fx.pop.hh_always)
    sim = oracle_lib.OracleSim(pop, fx.params)
    with pytest.raises(AssertionError):
        run_against_fixture(sim, fx)
