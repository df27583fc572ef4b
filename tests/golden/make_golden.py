"""Generates the committed golden vectors from the UNMODIFIED reference (needs /root/reference; run in the authoring
container only):   python tests/golden/make_golden.py

Each fixture = inputs in the reference's own format (agents[N,29], build[B,15], bld_visits_by_agents[N,V],
interaction_prob as CSR) + the literal contagious3.py:172-387 loop's per-day agent columns and output dict under
controlled draws (oracle/literal.py).  The reference ships no tests or golden vectors of its own;
these are outputs of the reference itself.
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import literal  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k if not hasattr(k, "item") else k.item()): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if hasattr(o, "item"):
        return o.item()
    return o


def save(name, setup, run, meta):
    ip = sp.csr_matrix(setup["interaction_prob"])
    ip.sort_indices()
    snaps = run["snapshots"]
    outputs = run["outputs"]
    for d in outputs["Results"]["Stats"].values():
        d.pop("Time", None)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        agents=setup["agents"], build=setup["build"], visits=setup["bld_visits_by_agents"],
        ip_indptr=ip.indptr.astype(np.int64), ip_indices=ip.indices.astype(np.int32), ip_data=ip.data,
        status_x2=np.array([np.round(s["status"] * 2) for s in snaps]).astype(np.uint8),
        t_inf=np.array([s["t_inf"] for s in snaps]).astype(np.int32),
        t_q=np.array([s["t_q"] for s in snaps]).astype(np.int32),
        t_adm=np.array([s["t_adm"] for s in snaps]).astype(np.int32),
        src_id=np.array([s["src_id"] for s in snaps]).astype(np.float64),
        open_next=np.array([s["open_next"] for s in snaps]).astype(np.uint8),
        meta=json.dumps(meta), outputs=json.dumps(_jsonable(outputs)))
    print(name, "N", len(setup["agents"]), "days", len(snaps), "nnz", ip.nnz,
          "total infected", outputs["Results"]["Stats"][len(snaps) - 1]["Total infected"])


def city(name, n_households, csv_seed, setup_seed, scenario, norm, draws_kind, draws_seed, days, force_hh=(0,), diff_key_fix=False):
    ind, bld = literal.subsample_csv("/tmp/dys_golden_" + name, n_households, csv_seed, force_households=force_hh)
    setup = literal.reference_setup(ind, bld, seed=setup_seed)
    n = len(setup["agents"])
    draws = literal.DenseDraws(n, draws_seed) if draws_kind == "dense" else literal.PhiloxDraws(n, draws_seed)
    run = literal.reference_days(setup, scenario, days, draws, overrides={"norm_factor": norm}, diff_key_fix=diff_key_fix)
    save(name, setup, run, dict(kind="city_subsample", diff_key_fix=diff_key_fix, n_households=n_households, csv_seed=csv_seed, setup_seed=setup_seed,
                                scenario=scenario, norm_factor=norm, draws=draws_kind, draws_seed=draws_seed,
                                choice_seed=12345))


def synthetic(name, n, seed, scenario, norm, draws_seed, days):
    from dysturbd_b200.synth import synth_population, to_reference
    pop = synth_population(n, seed=seed, seeds_per_22493=200)
    agents, build, visits, ip = to_reference(pop)
    # give one large household the id 0: under the reference's whole-row isin (contagious3.py:267) its status-2
    # members are quarantined whenever ANYONE is diagnosed -- in this epidemic that fires many times
    sizes = np.bincount(pop.hh)
    big = int(np.nonzero(sizes >= 5)[0][3])
    agents[pop.hh == big, 1] = 0.0
    setup = dict(agents=agents, build=build, bld_visits_by_agents=visits, interaction_prob=ip)
    run = literal.reference_days(setup, scenario, days, literal.PhiloxDraws(n, draws_seed), overrides={"norm_factor": norm})
    save(name, setup, run, dict(kind="synthetic", n=n, seed=seed, scenario=scenario, norm_factor=norm, draws="philox",
                                draws_seed=draws_seed, choice_seed=12345))


if __name__ == "__main__":
    assert literal.available(), "needs /root/reference"
    if "--diff" in sys.argv:
        # the DIFF (per-area) lockdown with the reference's 'Vis_R' / 'vis_R' key mismatch repaired at exec time
        # (literal.reference_days(diff_key_fix=True)); unrepaired, every DIFF scenario raises KeyError on day 1
        city("city300_philox_diff_all_keyfix", 300, 1, 7, ["DIFF", "ALL"], 3.0, "philox", 99, 60, diff_key_fix=True)
        city("city800_philox_diff_gradual_edu_rel_keyfix", 800, 2, 11, ["DIFF", "GRADUAL", "EDU", "REL"], 3.0, "philox", 5, 70, diff_key_fix=True)
        sys.exit(0)
    # shipped city, household sub-samples (household id 0 forced in: exercises the whole-row isin of contagious3.py:267)
    city("city300_dense_nolockdown", 300, 1, 7, ["noLockdown"], 3.0, "dense", 3, 60)
    city("city300_philox_gradual_all", 300, 1, 7, ["GRADUAL", "ALL"], 3.0, "philox", 99, 60)
    city("city800_philox_edu_rel", 800, 2, 11, ["EDU", "REL"], 1.0, "philox", 5, 70)
    synthetic("synth2000_philox_all", 2000, 5, ["ALL"], 6.0, 42, 60)
