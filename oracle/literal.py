"""TEST INFRASTRUCTURE (oracle) -- runs the UNMODIFIED reference code under controlled random draws.

Only usable where /root/reference exists (the authoring container).  Nothing here is copied from
the reference: the source of /root/reference/contagious3.py is read as text at run time and its
slices are exec'd in a namespace; create_random_data.create_data and communities.create_network
are imported from /root/reference.  It is used to (a) validate the C restatement in
oracle/dys_oracle.c and (b) generate the committed golden vectors under tests/golden/
(tests/golden/make_golden.py).  It never runs on the GPU box and never inside the product path.

Slices (line numbers of /root/reference/contagious3.py):
  defs  = lines 15..108   compute_R, compute_vis_R, building_lockdown
  setup = lines 119..170  create_network -> all-pairs shortest paths -> interaction_prob -> routines
  loop  = lines 172..387  the per-day `while` (the hot path)

Compatibility shims (no arithmetic touched):
  np.round_ (removed in numpy 2) -> np.round;  nx.to_scipy_sparse_matrix (removed in networkx 3) ->
  nx.to_scipy_sparse_array;  ragged np.array(list) -> dtype=object;  data paths passed explicitly.
"""
import os
import sys
import random as _pyrandom
import time as _time

import numpy as np

REFERENCE_DIR = os.environ.get("DYS_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "contagious3.py"))


def _source_lines():
    with open(os.path.join(REFERENCE_DIR, "contagious3.py")) as f:
        return f.read().split("\n")


def _dedent(lines, n):
    out = []
    for ln in lines:
        if ln.strip() == "":
            out.append("")
        else:
            assert ln[:n] == " " * n, ln
            out.append(ln[n:])
    return "\n".join(out)


def _import_reference():
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    if not hasattr(np, "round_"):
        np.round_ = np.round                       # shim 1 (create_random_data.py:38)
    import networkx as nx
    if not hasattr(nx, "to_scipy_sparse_matrix"):
        nx.to_scipy_sparse_matrix = nx.to_scipy_sparse_array   # shim 2 (contagious3.py:125)
    import parameters            # noqa: F401  (the reference's own module)
    import create_random_data
    import communities
    return parameters, create_random_data, communities, nx


def subsample_csv(out_dir, n_households, seed, force_households=()):
    """Write a household sub-sample of the shipped population (whole households; all buildings kept)
    so that the reference's own create_data can be run on a small city."""
    import pandas as pd
    ind = pd.read_csv(os.path.join(REFERENCE_DIR, "Data", "civ_withCar_bldg_np.csv"), header=None)
    rng = np.random.RandomState(seed)
    hh = ind[11].unique()
    pick = set(rng.choice(hh, size=min(n_households, len(hh)), replace=False).tolist())
    pick.update(force_households)
    sub = ind[ind[11].isin(pick)]
    os.makedirs(out_dir, exist_ok=True)
    ind_path = os.path.join(out_dir, "ind.csv")
    sub.to_csv(ind_path, header=False, index=False)
    return ind_path, os.path.join(REFERENCE_DIR, "Data", "bldg_with_inst_orig.csv")


class _NumpyProxy:
    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, name):
        return getattr(np, name)


class _RoutineRandom:
    """Stands in for `np.random` / `random.choice` inside the exec'd SETUP slice when the routine draws are injected
    (contagious3.py:150 `np.random.randint(2)`, :155 `choice(union)`).  Keyed on WHAT the draw is for: agent row n and
    the index q of the coin among that agent's coins; the choice that follows a successful coin shares its q.
        coin   = 1 if u(seed, 0, STREAM_ROUTINE_COIN,   n, q) >= 0.5 else 0
        choice = union[int(u(seed, 0, STREAM_ROUTINE_CHOICE, n, q) * len(union))]"""
    STREAM_COIN, STREAM_CHOICE = 5, 6

    def __init__(self, ns, seed):
        from . import philox_np
        self.ns, self.seed, self.p = ns, seed, philox_np
        self.agent, self.q = None, -1

    def randint(self, high):
        assert high == 2
        n = self.ns["n"]
        if n != self.agent:
            self.agent, self.q = n, -1
        self.q += 1
        return int(self.p.keyed_uniform(self.seed, 0, self.STREAM_COIN, np.array([n]), self.q)[0] >= 0.5)

    def choice(self, seq):
        u = self.p.keyed_uniform(self.seed, 0, self.STREAM_CHOICE, np.array([self.ns["n"]]), self.q)[0]
        return seq[int(u * len(seq))]


def reference_setup(ind_csv, bld_csv, seed, verbose=False, routine_seed=None):
    """create_data + contagious3.py:119-170, literally.  Returns dict(agents, households, build,
    bld_visits_by_agents, interaction_prob, timings).  routine_seed: inject keyed draws into the routine generation
    (:139-170) and also return the shortest-path matrix `dists` with its `nodes` (what that code reads)."""
    parameters, crd, communities, nx = _import_reference()
    np.random.seed(seed)
    _pyrandom.seed(seed)
    t0 = _time.time()
    agents, households, build, jobs = crd.create_data(ind_csv, bld_csv)
    t_data = _time.time() - t0
    src = _source_lines()
    setup = _dedent(src[118:170], 8)             # lines 119..170
    setup = setup.replace("bld_visits_by_agents = np.array(bld_visits_by_agents)",
                          "bld_visits_by_agents = np.array(bld_visits_by_agents, dtype=object)")  # shim 3
    ns = {
        "np": np, "nx": nx, "time": _time.time, "choice": _pyrandom.choice,
        "create_network": communities.create_network,
        "shortest_path": __import__("scipy.sparse.csgraph", fromlist=["shortest_path"]).shortest_path,
        "agents": agents, "households": households, "build": build,
        "k": parameters.k, "a_dist": parameters.a_dist, "bld_dist": parameters.bld_dist,
        "outputs": {}, "t": _time.time(),
        "print": (print if verbose else (lambda *a, **k: None)),
    }
    if routine_seed is not None:
        rr = _RoutineRandom(ns, routine_seed)
        ns["np"] = _NumpyProxy(rr)
        ns["choice"] = rr.choice
    t0 = _time.time()
    exec(compile(setup, "contagious3.py[119:170]", "exec"), ns)
    t_setup = _time.time() - t0
    extra = {"dists": ns["dists"], "nodes": ns["nodes"]} if routine_seed is not None else {}
    return {
        **extra,
        "agents": ns["agents"], "households": households, "build": ns["build"],
        "bld_visits_by_agents": np.asarray(ns["bld_visits_by_agents"], dtype=np.float64),
        "interaction_prob": ns["interaction_prob"],
        "G": ns["G"],
        "timings": {"create_data": t_data, "network+apsp+routines": t_setup,
                    "network_time": ns["outputs"].get("network_time"),
                    "interaction_time": ns["outputs"].get("interaction_time")},
    }


class DenseDraws:
    """Injected draws, literal form: for each day three pre-drawn arrays u_adm[N], u_death[N],
    u_pair[N, N] (row = susceptible, col = infectious, ORIGINAL indices) from numpy's own
    MT19937-based RandomState -- i.e. 'the identical pre-drawn uniform random array'."""

    def __init__(self, n, seed):
        self.n = n
        self.rs = np.random.RandomState(seed)
        self.days = {}

    def for_day(self, day):
        if day not in self.days:
            self.days[day] = (self.rs.random_sample(self.n), self.rs.random_sample(self.n),
                              self.rs.random_sample((self.n, self.n)))
        return self.days[day]

    def adm(self, day):
        return self.for_day(day)[0]

    def death(self, day):
        return self.for_day(day)[1]

    def pair(self, day, uninfected, infected):
        return self.for_day(day)[2][np.ix_(uninfected, infected)]


class PhiloxDraws:
    """Injected draws, keyed form: u(seed, day, stream, a, b) of oracle/philox_np.py, materialised lazily."""

    def __init__(self, n, seed):
        from . import philox_np
        self.n = n
        self.seed = seed
        self.p = philox_np
        self._idx = np.arange(n, dtype=np.int64)

    def adm(self, day):
        return self.p.keyed_uniform(self.seed, day, self.p.STREAM_ADMISSION, self._idx, 0)

    def death(self, day):
        return self.p.keyed_uniform(self.seed, day, self.p.STREAM_DEATH, self._idx, 0)

    def pair(self, day, uninfected, infected):
        if len(uninfected) == 0 or len(infected) == 0:
            return np.zeros((len(uninfected), len(infected)))
        return self.p.keyed_uniform(self.seed, day, self.p.STREAM_INFECTION,
                                    np.asarray(uninfected)[:, None], np.asarray(infected)[None, :])


class _RandomProxy:
    """Stands in for `np.random` inside the exec'd day loop.  The day consumes exactly three
    `random(...)` calls in a fixed order (contagious3.py:212, :224, :239)."""

    def __init__(self, ns, draws, choice_seed):
        self.ns = ns
        self.draws = draws
        self.calls_today = 0
        self.day_seen = None
        self._choice_rs = np.random.RandomState(choice_seed)
        self.choice_log = []

    def random(self, shape=None):
        day = self.ns["day"]
        if day != self.day_seen:
            self.day_seen = day
            self.calls_today = 0
        k = self.calls_today
        self.calls_today += 1
        if k == 0:
            out = self.draws.adm(day)
        elif k == 1:
            out = self.draws.death(day)
        elif k == 2:
            out = self.draws.pair(day, self.ns["uninfected"], self.ns["infected"])
        else:
            raise RuntimeError("unexpected 4th np.random.random call in one day")
        out = np.asarray(out, dtype=np.float64)
        assert out.shape == tuple(np.atleast_1d(shape)), (out.shape, shape)
        return out

    def choice(self, a, replace=True, size=None):
        # GRADUAL lockdowns only (contagious3.py:64,72,76): deterministic, logged for replay.
        out = self._choice_rs.choice(a, replace=replace, size=size)
        self.choice_log.append((self.ns["day"], np.array(out)))
        return out

    def randint(self, *a, **k):
        raise RuntimeError("np.random.randint is not used by the day loop")


SNAP_COLS = {"status": 13, "t_inf": 16, "n_inf": 17, "t_q": 19, "n_q": 20, "src_id": 21, "t_adm": 24, "n_adm": 25}


def reference_days(setup, scenario, max_days, draws, choice_seed=12345, capture_prob=False, verbose=False,
                   overrides=None, diff_key_fix=False):
    """Run contagious3.py:172-387 literally for at most max_days days on a COPY of `setup`.
    diff_key_fix: the ONE departure from the unmodified text this harness can make, and only on request: the `DIFF`
    branch reads outputs['Results']['SAs'][day-1]['Vis_R'] (:371) while the loop writes that dict under 'vis_R' (:319), so
    every DIFF scenario dies with KeyError on day 1.  With diff_key_fix the read uses the key that is written (SURVEY.md
    section 8f-3: "the DIFF per-area variant after fixing the 'Vis_R' key"); fixtures made this way say so in their meta.
    Returns dict(snapshots=[per-day dict of agent columns + open flags for the NEXT day], outputs=the
    reference's own outputs dict, day_seconds=[...], P=[per-day (uninfected, infected, infection_prob)] if capture_prob)."""
    parameters, _, _, _ = _import_reference()
    src = _source_lines()
    defs = "\n".join(src[14:108])                # lines 15..108
    loop = _dedent(src[171:387], 8)              # lines 172..387
    head = "while len(agents[agents[:, 13] == 2])"
    assert loop.count(head) == 1
    i = loop.index(head)
    j = loop.index(":", loop.index("> 0", i))
    loop = loop[:j] + " and day < __max_days__" + loop[j:]
    tail = "    day += 1"
    assert loop.count(tail) == 1
    loop = loop.replace(tail, "    __snapshot__()\n" + tail)
    if diff_key_fix:
        bad = "['Vis_R'][s]"
        assert loop.count(bad) == 1 and loop.count("['vis_R'] = sas_vis_R") == 1
        loop = loop.replace(bad, "['vis_R'][s]")
    if capture_prob:
        probe = "    rand_cont = np.random.random(infection_prob.shape)"
        assert loop.count(probe) == 1
        loop = loop.replace(probe, "    __probe__(uninfected, infected, infection_prob)\n" + probe)

    agents = np.array(setup["agents"], dtype=np.float64, copy=True)
    build = np.array(setup["build"], dtype=np.float64, copy=True)
    snaps, secs, probs = [], [], []
    ns = {
        "agents": agents, "build": build,
        "bld_visits_by_agents": np.array(setup["bld_visits_by_agents"], copy=True),
        "interaction_prob": setup["interaction_prob"],
        "scenario": list(scenario), "day": 0, "__max_days__": int(max_days),
        "outputs": {"Sim": 0, "Results": {"Stats": {}, "Buildings": {}, "SAs": {}, "IO_mat": {}}},
        "time": _time.time,
        "norm_factor": parameters.norm_factor, "recover": parameters.recover,
        "contagious_risk_day": parameters.contagious_risk_day, "quarantine": parameters.quarantine,
        "diagnosis": parameters.diagnosis, "hospital_recover": parameters.hospital_recover,
        "print": (print if verbose else (lambda *a, **k: None)),
    }
    ns.update(overrides or {})      # e.g. norm_factor for parameter sweeps (BASELINE.json config 5)
    rnd = _RandomProxy(ns, draws, choice_seed)
    ns["np"] = _NumpyProxy(rnd)
    t_last = [_time.perf_counter()]

    def snapshot():
        now = _time.perf_counter()
        secs.append(now - t_last[0])
        s = {k: ns["agents"][:, c].copy() for k, c in SNAP_COLS.items()}
        s["open_next"] = ns["build"][:, 10].copy()
        snaps.append(s)
        t_last[0] = _time.perf_counter()

    def probe(u, i, p):
        probs.append((np.array(u), np.array(i), np.array(p)))

    ns["__snapshot__"] = snapshot
    ns["__probe__"] = probe
    exec(compile(defs, "contagious3.py[15:108]", "exec"), ns)
    t_last[0] = _time.perf_counter()
    exec(compile(loop, "contagious3.py[172:387]", "exec"), ns)
    return {"snapshots": snaps, "outputs": ns["outputs"], "day_seconds": secs, "P": probs,
            "choice_log": rnd.choice_log, "agents_final": ns["agents"], "build_final": ns["build"]}
