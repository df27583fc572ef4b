"""Pins the oracle to the reference itself, live: runs the literal contagious3.py loop (oracle/literal.py) on a fresh
household sub-sample under dense MT19937 draws and checks state AND the float64 infection probabilities (0 ulp).
Only runs where /root/reference exists (the authoring container); the golden fixtures cover the GPU box."""
import numpy as np
import pytest

from oracle import literal

pytestmark = pytest.mark.skipif(not literal.available(), reason="/root/reference not present")


def test_literal_loop_vs_oracle_state_and_probabilities(oracle_lib, tmp_path):
    from dysturbd_b200.population import EpiParams, Population
    ind, bld = literal.subsample_csv(str(tmp_path), 200, 3, force_households=(0,))
    s = literal.reference_setup(ind, bld, seed=5)
    n = len(s["agents"])
    draws = literal.DenseDraws(n, 8)
    r = literal.reference_days(s, ["ALL"], 25, draws, overrides={"norm_factor": 4.0}, capture_prob=True)
    pop = Population.from_reference(s["agents"], s["build"], s["bld_visits_by_agents"], s["interaction_prob"])
    o = oracle_lib.OracleSim(pop, EpiParams(norm_factor=4.0), capture_prob=True)
    idx = {v: i for i, v in enumerate(pop.agent_ids)}
    nz = 0
    for d, snap in enumerate(r["snapshots"]):
        o.step(d, *draws.for_day(d))
        st = o.state()
        assert np.array_equal(st["status"], np.round(snap["status"] * 2).astype(np.uint8))
        for k in ("t_inf", "t_q", "t_adm"):
            assert np.array_equal(st[k], snap[k])
        assert np.array_equal(st["src"], [idx[v] if v != 0 else -1 for v in snap["src_id"]])
        u, i, P = r["P"][d]
        if len(u) and len(i):
            assert np.array_equal(o.pair_prob[np.ix_(u, i)], P)      # float64, bit for bit
            nz += int((P > 0).sum())
        o.set_open(snap["open_next"])
    assert nz > 100
