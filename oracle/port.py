"""TEST INFRASTRUCTURE (oracle) -- ctypes wrapper of oracle/dys_oracle.c (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdys_oracle.so")

N_STATS = 64
HIST0 = 32
STAT_NAMES = ["active", "daily_inf", "recovered", "quarantined", "daily_quar", "hosp", "daily_hosp", "dead",
              "daily_deaths", "ever_infected", "new_diag", "n_events", "sus0", "inf0", "edges_scanned",
              "candidates", "exposed", "hits"]


def build(force=False):
    src = os.path.join(HERE, "dys_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-ffp-contract=off", "-shared", src, "-o", LIB, "-lm"])
    return LIB


class _Ctx(C.Structure):
    _fields_ = [
        ("N", C.c_int64), ("H", C.c_int64), ("B", C.c_int64), ("S", C.c_int64),
        ("V", C.c_int32),
        ("recover", C.c_int32), ("hospital_recover", C.c_int32), ("diagnosis", C.c_int32), ("quarantine", C.c_int32),
        ("adm_min_day", C.c_int32), ("adm_max_day", C.c_int32), ("death_min_adm_day", C.c_int32),
        ("norm_factor", C.c_double), ("compliance", C.c_double), ("pdf", C.c_double * 64),
        ("seed", C.c_uint64),
        ("status", C.c_void_p), ("t_inf", C.c_void_p), ("t_q", C.c_void_p), ("t_adm", C.c_void_p), ("src", C.c_void_p),
        ("hh", C.c_void_p), ("home", C.c_void_p), ("sa", C.c_void_p),
        ("risk", C.c_void_p), ("p_adm", C.c_void_p), ("p_death", C.c_void_p),
        ("routine", C.c_void_p),
        ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
        ("open", C.c_void_p),
        ("hh_always", C.c_void_p), ("trig_ptr", C.c_void_p), ("trig_hh", C.c_void_p),
        ("u_adm", C.c_void_p), ("u_death", C.c_void_p), ("u_pair", C.c_void_p),
        ("stats", C.c_void_p), ("sa_hist", C.c_void_p), ("bld_infected", C.c_void_p),
        ("ev_agent", C.c_void_p), ("ev_src", C.c_void_p), ("pair_prob", C.c_void_p),
        ("lo", C.c_int64), ("n_glob", C.c_int64), ("rt_lo", C.c_int64), ("ib_global", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_day.argtypes = [C.POINTER(_Ctx), C.c_int32]
        _lib.orc_day.restype = C.c_int
        _lib.orc_keyed_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.orc_keyed_uniform.restype = C.c_double
        _lib.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_snapshot.argtypes = [C.POINTER(_Ctx), C.c_int32, C.c_void_p]
    return _lib


def routines(home, anchors, out_csr, in_csr, near_csr, seed, V=10, k=2):
    """oracle/dys_oracle.c:orc_routines -- routine generation (contagious3.py:139-170) on sorted neighbourhood lists"""
    L = lib()
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    home, anchors = c(home, np.int32), c(anchors, np.int32)
    ptrs = [c(x[0], np.int64) for x in (out_csr, in_csr, near_csr)]
    idxs = [c(x[1], np.int32) for x in (out_csr, in_csr, near_csr)]
    out = np.zeros((len(home), V), dtype=np.int32)
    L.orc_routines.argtypes = [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 8 + [C.c_uint64, C.c_void_p]
    rc = L.orc_routines(len(home), V, k, _p(home), _p(anchors), _p(ptrs[0]), _p(idxs[0]), _p(ptrs[1]), _p(idxs[1]),
                        _p(ptrs[2]), _p(idxs[2]), seed, _p(out))
    if rc != 0:
        raise RuntimeError("orc_routines failed")
    return out


def max_threads():
    return int(lib().orc_max_threads())


def keyed_uniform(seed, day, stream, a, b=0):
    return lib().orc_keyed_uniform(seed, day, stream, a, b)


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(ctr.ctypes.data, key.ctypes.data, out.ctypes.data)
    return out


def _p(a):
    return None if a is None else a.ctypes.data


class OracleSim:
    """One population, stepped day by day on the CPU.  `pop` is a dysturbd_b200.population.Population (its
    state arrays are COPIED); `params` a dysturbd_b200.population.EpiParams."""

    def __init__(self, pop, params, capture_prob=False, threads=None):
        self.pop, self.params = pop, params
        self.threads = threads
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        self.status = c(pop.status, np.uint8).copy()
        self.t_inf = c(pop.t_inf, np.int32).copy()
        self.t_q = c(pop.t_q, np.int32).copy()
        self.t_adm = c(pop.t_adm, np.int32).copy()
        self.src = c(pop.src, np.int32).copy()
        self.open = c(pop.open, np.uint8).copy()
        self._keep = dict(
            hh=c(pop.hh, np.int32), home=c(pop.home, np.int32), sa=c(pop.sa, np.int32),
            risk=c(pop.risk, np.float64), p_adm=c(pop.p_adm, np.float64), p_death=c(pop.p_death, np.float64),
            routine=c(pop.routine, np.int32), rowptr=c(pop.ip_rowptr, np.int64), col=c(pop.ip_col, np.int32),
            val=c(pop.ip_val, np.float64), hh_always=c(pop.hh_always, np.uint8),
            trig_ptr=c(pop.trig_ptr, np.int64), trig_hh=c(pop.trig_hh, np.int32))
        self._stats = np.zeros(N_STATS, dtype=np.int64)
        self._sa_hist = np.zeros((max(pop.S, 1), 21), dtype=np.int32)
        self.bld_infected = np.zeros(pop.B, dtype=np.int32)
        self.ev_agent = np.zeros(pop.N, dtype=np.int32)
        self.ev_src = np.zeros(pop.N, dtype=np.int32)
        self.pair_prob = np.zeros((pop.N_glob, pop.N_glob)) if capture_prob else None
        k = self._keep
        ctx = _Ctx()
        ctx.N, ctx.H, ctx.B, ctx.S, ctx.V = pop.N, pop.H, pop.B, pop.S, pop.V
        p = params
        ctx.recover, ctx.hospital_recover, ctx.diagnosis, ctx.quarantine = p.recover, p.hospital_recover, p.diagnosis, p.quarantine
        ctx.adm_min_day, ctx.adm_max_day, ctx.death_min_adm_day = p.adm_min_day, p.adm_max_day, p.death_min_adm_day
        ctx.norm_factor, ctx.compliance, ctx.seed = p.norm_factor, p.compliance, p.seed
        for i in range(64):
            ctx.pdf[i] = float(p.pdf[i])
        ctx.status, ctx.t_inf, ctx.t_q, ctx.t_adm, ctx.src = map(_p, (self.status, self.t_inf, self.t_q, self.t_adm, self.src))
        for name in ("hh", "home", "sa", "risk", "p_adm", "p_death", "routine", "rowptr", "col", "val",
                     "hh_always", "trig_ptr", "trig_hh"):
            setattr(ctx, name, _p(k[name]))
        ctx.open = _p(self.open)
        ctx.stats, ctx.sa_hist, ctx.bld_infected = _p(self._stats), _p(self._sa_hist), _p(self.bld_infected)
        ctx.ev_agent, ctx.ev_src, ctx.pair_prob = _p(self.ev_agent), _p(self.ev_src), _p(self.pair_prob)
        ctx.lo, ctx.n_glob, ctx.rt_lo = pop.lo, pop.N_glob, pop.halo[0]
        ctx.ib_global = None
        self._ib_global = None
        self.ctx = ctx

    def set_open(self, open_flags):
        self.open[:] = np.asarray(open_flags, dtype=np.uint8)

    def set_state(self, status, t_inf, t_q, t_adm, src):
        """the mutable agent columns, in place (same surface as Engine.set_state)"""
        self.status[:] = np.asarray(status, dtype=np.uint8)
        self.t_inf[:] = np.asarray(t_inf, dtype=np.int32)
        self.t_q[:] = np.asarray(t_q, dtype=np.int32)
        self.t_adm[:] = np.asarray(t_adm, dtype=np.int32)
        self.src[:] = np.asarray(src, dtype=np.int32)

    def step(self, day, u_adm=None, u_death=None, u_pair=None):
        keep = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None for x in (u_adm, u_death, u_pair)]
        self.ctx.u_adm, self.ctx.u_death, self.ctx.u_pair = map(_p, keep)
        if self.pair_prob is not None:
            self.pair_prob[:] = 0.0
        if self.threads is not None:
            lib().orc_set_threads(int(self.threads))
        rc = lib().orc_day(C.byref(self.ctx), int(day))
        if rc != 0:
            raise RuntimeError("orc_day failed")
        return self._stats.copy()

    # sharded populations: the same two-half day as dys_step_begin / dys_step_end (include/dysturbd.h)
    def step_begin(self, day):
        self._ib_local = np.zeros(self.pop.N, dtype=np.uint8)
        lib().orc_snapshot(C.byref(self.ctx), int(day), self._ib_local.ctypes.data)

    def ib_local(self):
        return self._ib_local

    def set_ib_global(self, ib):
        self._ib_global = np.ascontiguousarray(ib, dtype=np.uint8)
        assert self._ib_global.size >= self.pop.N_glob
        self.ctx.ib_global = self._ib_global.ctypes.data

    def step_end(self, day):
        return self.step(day)

    # the same read-out surface as dysturbd_b200.engine.Engine, so tests can drive either through one harness
    def stats(self, day=None):
        return self._stats.copy()

    def sa_hist(self, day=None):
        return self._sa_hist[:self.pop.S].copy()

    def building_counts(self, day=None):
        return self.bld_infected.copy()

    def state(self):
        return dict(status=self.status.copy(), t_inf=self.t_inf.copy(), t_q=self.t_q.copy(),
                    t_adm=self.t_adm.copy(), src=self.src.copy())

    def events(self, day=None):
        n = int(self._stats[STAT_NAMES.index("n_events")])
        return self.ev_agent[:n].copy(), self.ev_src[:n].copy()


def similarity_edges(age, income, x, y, hh, min_prob, w_a=1.0, w_i=1.0, w_d=1.0):
    """oracle/dys_oracle.c:orc_similarity_edges -- communities.py:22-40, 65-70 without the N x N matrices.
    Returns (indptr[n+1], col, prob, (max age diff, max income diff, max distance))."""
    L = lib()
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    age, income, x, y, hh = c(age, np.float64), c(income, np.float64), c(x, np.float64), c(y, np.float64), c(hh, np.int64)
    n = len(age)
    L.orc_similarity_edges.argtypes = [C.c_int64] + [C.c_void_p] * 5 + [C.c_double] * 4 + [C.c_int] + [C.c_void_p] * 5
    count, maxes = np.zeros(n, dtype=np.int64), np.zeros(3)
    args = (n, _p(age), _p(income), _p(x), _p(y), _p(hh), w_a, w_i, w_d, min_prob)
    assert L.orc_similarity_edges(*args, 0, _p(count), None, None, None, _p(maxes)) == 0
    indptr = np.concatenate([[0], np.cumsum(count)]).astype(np.int64)
    col, prob = np.zeros(indptr[-1], dtype=np.int32), np.zeros(indptr[-1])
    assert L.orc_similarity_edges(*args, 1, None, _p(indptr), _p(col), _p(prob), _p(maxes)) == 0
    return indptr, col, prob, tuple(maxes)


def shortest_paths(indptr, indices, w, sources):
    """oracle/dys_oracle.c:orc_shortest_paths -- scipy.sparse.csgraph.shortest_path(g, directed=True, indices=sources)"""
    L = lib()
    indptr, indices = np.ascontiguousarray(indptr, dtype=np.int64), np.ascontiguousarray(indices, dtype=np.int32)
    w, sources = np.ascontiguousarray(w, dtype=np.float64), np.ascontiguousarray(sources, dtype=np.int32)
    n = len(indptr) - 1
    out = np.empty((len(sources), n))
    L.orc_shortest_paths.argtypes = [C.c_int64] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 2
    if L.orc_shortest_paths(n, _p(indptr), _p(indices), _p(w), len(sources), _p(sources), _p(out)) != 0:
        raise RuntimeError("orc_shortest_paths: negative weight, bad index or bad source")
    return out


def building_edges(fs, x, y, min_prob):
    """oracle/dys_oracle.c:orc_building_edges -- communities.py:12-21, 53-54 (gravity model between buildings, beta = 2).
    Returns (indptr[B+1], col, prob, row_sum[B])."""
    L = lib()
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    fs, x, y = c(fs), c(x), c(y)
    B = len(fs)
    L.orc_building_edges.argtypes = [C.c_int64] + [C.c_void_p] * 3 + [C.c_double, C.c_int] + [C.c_void_p] * 5
    count, row_sum = np.zeros(B, dtype=np.int64), np.zeros(B)
    assert L.orc_building_edges(B, _p(fs), _p(x), _p(y), min_prob, 0, _p(count), None, None, None, _p(row_sum)) == 0
    indptr = np.concatenate([[0], np.cumsum(count)]).astype(np.int64)
    col, prob = np.zeros(indptr[-1], dtype=np.int32), np.zeros(indptr[-1])
    assert L.orc_building_edges(B, _p(fs), _p(x), _p(y), min_prob, 1, None, _p(indptr), _p(col), _p(prob), None) == 0
    return indptr, col, prob, row_sum
