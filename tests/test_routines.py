"""routine generation (contagious3.py:139-170).  The golden fixture is the UNMODIFIED reference's own
bld_visits_by_agents on a 120-household sub-sample of the shipped city with keyed draws injected into its routine loop
(tests/golden/make_golden_routines.py); the C oracle and the CUDA kernel must both reproduce it exactly."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "routines_city120.npz")


def _fixture():
    z = np.load(GOLDEN)
    nb = dict(bld_ids=z["bld_ids"], home=z["home"], anchors=z["anchors"], out=(z["out_ptr"], z["out_idx"]),
              inn=(z["in_ptr"], z["in_idx"]), near=(z["near_ptr"], z["near_idx"]))
    return nb, int(z["seed"]), int(z["k"]), z["visits"]


def _same(vis, exp):
    return np.array_equal(np.nan_to_num(vis, nan=-1.0), np.nan_to_num(exp, nan=-1.0))


def test_oracle_routines_match_the_reference(oracle_lib):
    from dysturbd_b200 import routines
    nb, seed, k, exp = _fixture()
    r = oracle_lib.routines(nb["home"], nb["anchors"], nb["out"], nb["inn"], nb["near"], seed, V=exp.shape[1], k=k)
    assert _same(routines.to_visits(r, nb["bld_ids"]), exp)
    lengths = (r >= 0).sum(axis=1)
    assert lengths.min() >= 2 and lengths.max() == 10          # home + up to 3 x (2 stops + anchor)


@pytest.mark.gpu
def test_device_routines_match_the_reference():
    from dysturbd_b200 import routines
    nb, seed, k, exp = _fixture()
    r = routines.generate(nb, seed, V=exp.shape[1], k=k)
    assert _same(routines.to_visits(r, nb["bld_ids"]), exp)


@pytest.mark.gpu
def test_device_routines_match_the_oracle_on_a_large_city(oracle_lib):
    """200 k agents, 10 k buildings, random sorted neighbourhood lists of uneven length (empty intersections included)"""
    from dysturbd_b200 import engine, routines
    rng = np.random.default_rng(3)
    N, B = 200_000, 10_000

    def lists(rows, mean):
        ln = rng.poisson(mean, rows)
        ln[rng.random(rows) < 0.05] = 0
        ptr = np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)
        idx = np.concatenate([np.sort(rng.choice(B, n, replace=False)) for n in ln]).astype(np.int32) if ln.sum() else np.zeros(0, np.int32)
        return ptr, idx
    home = rng.integers(0, B, N).astype(np.int32)
    anchors = np.stack([np.where(rng.random(N) < 0.8, rng.integers(0, B, N), -1), np.where(rng.random(N) < 0.5, rng.integers(0, B, N), -1),
                        home], axis=1).astype(np.int32)
    near = lists(N, 3)
    near_ptr, near_idx = near
    # every agent's near list holds its home (distance 0 < a_dist), so the union is never empty
    fixed = []
    for a in range(N):
        row = near_idx[near_ptr[a]:near_ptr[a + 1]]
        fixed.append(np.union1d(row, [home[a]]))
    near = (np.concatenate([[0], np.cumsum([len(x) for x in fixed])]).astype(np.int64), np.concatenate(fixed).astype(np.int32))
    nb = dict(bld_ids=np.arange(B, dtype=np.float64) + 1, home=home, anchors=anchors, out=lists(B, 40), inn=lists(B, 40), near=near)
    g = routines.generate(nb, 99, V=10, k=2)
    o = oracle_lib.routines(home, anchors, nb["out"], nb["inn"], nb["near"], 99, V=10, k=2)
    assert np.array_equal(g, o)
    assert ((g >= 0).sum(axis=1) > 3).mean() > 0.5
    with pytest.raises(engine.DysError):
        bad = dict(nb)
        bad["out"] = (nb["out"][0], nb["out"][1][::-1].copy())
        routines.generate(bad, 1)
