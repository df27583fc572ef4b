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
                      choice_seed=fx.meta["choice_seed"], backend=oracle_lib.OracleSim, **fx.model_kw)
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


def test_written_json_is_the_reference_file(oracle_lib, tmp_path):
    """EpidemicModel.write_json == what contagious3.py:389-392 dumps: the golden `outputs` were stored after a JSON round
    trip of the reference's dict, so the file this library writes must load to exactly that (every key, every float)."""
    import json
    from dysturbd_b200.host import EpidemicModel
    fx = Fixture("city300_philox_gradual_all")
    m = EpidemicModel(fx.agents, fx.build, fx.visits, fx.ip, scenario=fx.scenario, params=fx.params,
                      choice_seed=fx.meta["choice_seed"], backend=oracle_lib.OracleSim, **fx.model_kw)
    m.run(max_days=fx.days)
    path = tmp_path / "sim0.json"
    m.write_json(path, total_time=1.5)
    got = json.load(open(path))
    assert got["Total_time"] == 1.5
    for key in ("Stats", "SAs", "Buildings", "IO_mat"):
        assert got["Results"][key] == {d: fx.outputs["Results"][key][d] for d in got["Results"][key]}, key
        assert len(got["Results"][key]) == fx.days
    assert len(got["Results"]["Infections chain"]) == fx.pop.N


def test_checkpoint_and_resume(oracle_lib):
    """a run stopped after 9 days and continued from checkpoint() in a NEW model equals the uninterrupted run: agent
    state, the host's R history (the lockdown decision lags 7 days), the flags and the np.random.choice stream all travel"""
    from dysturbd_b200.host import EpidemicModel
    fx = Fixture("city300_philox_gradual_all")
    mk = lambda: EpidemicModel(fx.agents, fx.build, fx.visits, fx.ip, scenario=fx.scenario, params=fx.params,
                               choice_seed=fx.meta["choice_seed"], backend=oracle_lib.OracleSim, **fx.model_kw)
    a = mk()
    for _ in range(9):
        a.step()
    ck = a.checkpoint()
    b = mk()
    b.restore(ck)
    while b.day < fx.days and b.active():
        b.step()
    compare_outputs(b.outputs, fx.outputs, fx.days)
    assert np.array_equal(b.open.astype(np.uint8), fx.exp["open_next"][fx.days - 1])
    st = b.sim.state()
    assert np.array_equal(st["status"], fx.exp["status_x2"][fx.days - 1]) and np.array_equal(st["t_q"], fx.exp["t_q"][fx.days - 1])


def test_compact_outputs_hold_the_same_numbers(oracle_lib):
    """compact=True (arrays instead of the reference's nested dicts, for populations with hundreds of areas) carries
    exactly the reference's values: per-area R and vis_R (floats compared with ==), building counts and fractions, and
    the origin matrix rebuilt from the day's (infected area, infector area) pairs."""
    from dysturbd_b200.host import EpidemicModel
    fx = Fixture("city800_philox_edu_rel")
    m = EpidemicModel(fx.agents, fx.build, fx.visits, fx.ip, scenario=fx.scenario, params=fx.params,
                      choice_seed=fx.meta["choice_seed"], backend=oracle_lib.OracleSim, compact=True)
    out = m.run(max_days=fx.days)["Results"]
    sas = [float(x) for x in m.pop.sa_ids]
    for d in range(fx.days):
        exp = fx.outputs["Results"]
        assert out["Stats"][d] == exp["Stats"][str(d)]
        for k, sa in enumerate(sas):
            assert out["SAs"][d]["R"][k] == exp["SAs"][str(d)]["R"][str(sa)], (d, sa)
            assert out["SAs"][d]["vis_R"][k] == exp["SAs"][str(d)]["vis_R"][str(sa)], (d, sa)
        eb = exp["Buildings"][str(d)]
        tab = m.buildings_table(d)
        assert len(eb) == len(tab)
        for row, (bid, cnt, frac) in zip(tab, eb):
            assert row[0] == cnt and row[1] == frac
        mat = {sa: {sa2: 0 for sa2 in sas} for sa in sas}
        for a, s in m.io_pairs(d):
            mat[sas[a]][sas[s]] += 1
        assert {str(a): {str(b): v for b, v in r.items()} for a, r in mat.items()} == exp["IO_mat"][str(d)]
