"""ctypes binding of include/dysturbd.h (dysturbd_b200/libdysturbd.so) -- the reference-side stub a maintainer
would add.  Arrays may be numpy arrays (host) or torch tensors (CPU or CUDA on the engine's
device); only their data pointers cross the ABI.

There is deliberately NO fallback: if the shared library is missing or no B200 is visible, construction fails.
"""
import ctypes as C
import os

import numpy as np

from .population import EpiParams, Population

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DYS_LIB") or os.path.join(HERE, "libdysturbd.so")      # (DYS_LIB: developer ablation builds)

N_STATS = 64
HIST0 = 32
HIST_LEN = 21
N_KERNELS = 3
KERNEL_NAMES = ("k_snapshot", "k_push", "k_agent")
STAT = {name: i for i, name in enumerate(
    ["active", "daily_inf", "recovered", "quarantined", "daily_quar", "hosp", "daily_hosp", "dead", "daily_deaths",
     "ever_infected", "new_diag", "n_events", "sus0", "inf0", "edges_scanned", "candidates", "exposed", "hits",
     "changed"])}

WANT_SA_HIST, WANT_BUILDING_COUNTS, WANT_EVENTS, WANT_IO_MAT, TIME_KERNELS = 1, 2, 4, 8, 16
SC_GRADUAL, SC_ALL, SC_EDU, SC_REL = 1, 2, 4, 8
CLASS_NONRES, CLASS_EDU, CLASS_REL = 2, 4, 8
ST_CLOSED, ST_R_BITS, ST_KNOWN_R_BITS = 28, 29, 30
BUF_STATUS, BUF_INFECTIOUS, BUF_STATS, BUF_AGENT_REC = 0, 1, 2, 3
PEER_HANDLE_BYTES = 128


class DysError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dysturbd error %d: %s" % (code, msg))
        self.code = code


class _Params(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int64), ("n_households", C.c_int64), ("n_buildings", C.c_int64), ("n_areas", C.c_int64),
        ("n_slots", C.c_int32), ("recover", C.c_int32), ("hospital_recover", C.c_int32), ("diagnosis", C.c_int32),
        ("quarantine", C.c_int32), ("adm_min_day", C.c_int32), ("adm_max_day", C.c_int32), ("death_min_adm_day", C.c_int32),
        ("norm_factor", C.c_double), ("compliance", C.c_double), ("pdf_table", C.c_double * 64),
        ("seed", C.c_uint64), ("flags", C.c_int32), ("n_replicas", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32), ("shard_lo", C.c_int64), ("shard_hi", C.c_int64),
        ("halo_lo", C.c_int64), ("halo_hi", C.c_int64),
        ("hh_lo", C.c_int64), ("hh_hi", C.c_int64), ("bld_lo", C.c_int64), ("bld_hi", C.c_int64),
        ("area_lo", C.c_int64), ("area_hi", C.c_int64),
    ]


class _DayOutputs(C.Structure):
    _fields_ = [("day", C.c_int32), ("events_inline", C.c_int32), ("n_events", C.c_int64), ("stats", C.POINTER(C.c_int64)),
                ("sa_hist", C.POINTER(C.c_int32)), ("building_counts", C.POINTER(C.c_int32)), ("events", C.POINTER(C.c_int32)),
                ("events_capacity", C.c_int64)]


class _Injected(C.Structure):
    _fields_ = [("u_adm", C.c_void_p), ("u_death", C.c_void_p), ("u_pair", C.c_void_p)]


EXPORTS = ["dys_create", "dys_destroy", "dys_last_error", "dys_load_population", "dys_load_graph", "dys_load_quirk",
           "dys_set_open", "dys_set_run", "dys_set_state", "dys_step", "dys_step_many", "dys_step_begin", "dys_step_end", "dys_get_stats",
           "dys_get_sa_hist", "dys_get_building_counts", "dys_get_events", "dys_get_io_mat", "dys_get_state",
           "dys_get_pair_prob", "dys_get_outputs", "dys_sync", "dys_get_kernel_times", "dys_device_buffer", "dys_cuda_stream",
           "dys_device_bytes", "dys_version", "dys_keyed_uniform_device", "dys_peer_export", "dys_peer_attach", "dys_peer_detach",
           "dys_set_replica_params", "dys_select_replica", "dys_abort", "dys_h2d_bytes", "dys_compute_r_rows", "dys_set_lockdown", "dys_set_lockdown_areas", "dys_generate_routines", "dys_consume_day", "dys_similarity_edges", "dys_building_edges", "dys_shortest_paths", "dys_shortest_paths_csr"]

_lib = None


def load_library():
    """dlopen the in-tree library (built by dysturbd_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DysError(-2, "%s is missing: run `python -m dysturbd_b200._build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.dys_create.argtypes = [C.POINTER(_Params), C.c_int, C.POINTER(vp)]
    lib.dys_destroy.argtypes = [vp]
    lib.dys_last_error.argtypes = [vp]
    lib.dys_last_error.restype = C.c_char_p
    lib.dys_load_population.argtypes = [vp] + [vp] * 12
    lib.dys_load_graph.argtypes = [vp, vp, vp, vp]
    lib.dys_load_quirk.argtypes = [vp, vp, vp, vp]
    lib.dys_set_open.argtypes = [vp, vp]
    lib.dys_set_run.argtypes = [vp, C.c_double, C.c_double, C.c_uint64]
    lib.dys_set_state.argtypes = [vp] + [vp] * 5
    lib.dys_step.argtypes = [vp, i32, C.POINTER(_Injected)]
    lib.dys_step_many.argtypes = [vp, i32, i32, vp]
    lib.dys_step_begin.argtypes = [vp, i32, C.POINTER(_Injected)]
    lib.dys_step_end.argtypes = [vp, i32]
    lib.dys_get_stats.argtypes = [vp, i32, vp]
    lib.dys_get_sa_hist.argtypes = [vp, i32, vp]
    lib.dys_get_building_counts.argtypes = [vp, i32, vp]
    lib.dys_get_events.argtypes = [vp, i32, vp, vp, i64, C.POINTER(i64)]
    lib.dys_get_io_mat.argtypes = [vp, i32, vp]
    lib.dys_get_state.argtypes = [vp] + [vp] * 5
    lib.dys_get_pair_prob.argtypes = [vp, vp]
    lib.dys_get_outputs.argtypes = [vp, i32, C.POINTER(_DayOutputs)]
    lib.dys_sync.argtypes = [vp]
    lib.dys_get_kernel_times.argtypes = [vp, vp, vp]
    lib.dys_device_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    lib.dys_cuda_stream.argtypes = [vp, C.POINTER(vp)]
    lib.dys_device_bytes.argtypes = [vp]
    lib.dys_device_bytes.restype = i64
    lib.dys_version.restype = C.c_char_p
    lib.dys_keyed_uniform_device.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, i64, vp]
    lib.dys_peer_export.argtypes = [vp, vp]
    lib.dys_peer_attach.argtypes = [vp, C.c_int, vp, vp]
    lib.dys_peer_detach.argtypes = [vp]
    lib.dys_set_replica_params.argtypes = [vp, vp, vp, vp]
    lib.dys_select_replica.argtypes = [vp, i32]
    lib.dys_abort.argtypes = [vp]
    lib.dys_set_lockdown.argtypes = [vp, i32, vp]
    lib.dys_set_lockdown_areas.argtypes = [vp, i32, vp, vp]
    lib.dys_consume_day.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, i64, vp]
    lib.dys_generate_routines.argtypes = [C.c_int, i64, i64, i32, i32] + [vp] * 8 + [C.c_uint64, vp]
    lib.dys_h2d_bytes.argtypes = [vp]
    lib.dys_h2d_bytes.restype = i64
    lib.dys_compute_r_rows.argtypes = [vp, i32, i64, i64, vp, i32, vp]
    _lib = lib
    return lib


def _ptr(x, dtype=None, n=None):
    """data pointer of a numpy array or torch tensor (must be contiguous, of the expected dtype)."""
    if x is None:
        return None, None
    if isinstance(x, np.ndarray):
        if dtype is not None and x.dtype != dtype:
            raise TypeError("expected %s, got %s" % (dtype, x.dtype))
        if not x.flags.c_contiguous:
            raise ValueError("array must be C-contiguous")
        if n is not None and x.size != n:
            raise ValueError("expected %d elements, got %d" % (n, x.size))
        return x.ctypes.data, x
    if hasattr(x, "data_ptr"):      # torch tensor (CPU or CUDA); the DLPack exchange of the north star reduces to this
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if n is not None and x.numel() != n:
            raise ValueError("expected %d elements, got %d" % (n, x.numel()))
        return x.data_ptr(), x
    raise TypeError("numpy array or torch tensor expected")


def compute_R_rows(hist, pdf, recover, out=None):
    """compute_R (contagious3.py:29-42) for every row of an int32 / int64 [rows, >=21] histogram (host memory), by the
    library's C helper: the reference's order of float operations, bit for bit"""
    lib = _lib or load_library()
    if not hist.flags.c_contiguous:
        hist = np.ascontiguousarray(hist)
    rows, stride = (1, hist.shape[0]) if hist.ndim == 1 else hist.shape
    if out is None:
        out = np.empty(rows, dtype=np.float64)
    rc = lib.dys_compute_r_rows(hist.ctypes.data, hist.dtype.itemsize, rows, stride, pdf.ctypes.data, int(recover), out.ctypes.data)
    if rc != 0:
        raise DysError(rc, "dys_compute_r_rows: bad argument")
    return out


def keyed_uniform_device(seed, day, stream, a, b, device=0):
    """u(day, stream, a, b) evaluated by the CUDA Philox (tests)."""
    lib = load_library()
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(np.broadcast_to(b, a.shape), dtype=np.uint32)
    out = np.zeros(a.shape, dtype=np.float64)
    rc = lib.dys_keyed_uniform_device(device, seed, day, stream, a.ctypes.data, b.ctypes.data, a.size, out.ctypes.data)
    if rc != 0:
        raise DysError(rc, (lib.dys_last_error(None) or b"").decode())
    return out


class Engine:
    """One population on one GPU.  Mirrors the call order of run_EM() (epidemiological_model.py:224-234):
    `step(day)` is one pass of the reference's day loop; outputs are the integer arrays behind contagious3.py:326-366."""

    def __init__(self, pop: Population, params: EpiParams, device=0, flags=WANT_SA_HIST | WANT_BUILDING_COUNTS | WANT_EVENTS,
                 rank=0, world=1, replicas=1):
        self.lib = load_library()
        self.pop, self.params = pop, params
        self.N_glob = pop.N_glob
        p = _Params()
        p.n_agents, p.n_households, p.n_buildings, p.n_areas, p.n_slots = self.N_glob, pop.H, pop.B, pop.S, pop.V
        p.recover, p.hospital_recover, p.diagnosis, p.quarantine = params.recover, params.hospital_recover, params.diagnosis, params.quarantine
        p.adm_min_day, p.adm_max_day, p.death_min_adm_day = params.adm_min_day, params.adm_max_day, params.death_min_adm_day
        p.norm_factor, p.compliance, p.seed, p.flags = params.norm_factor, params.compliance, params.seed, flags
        for i in range(64):
            p.pdf_table[i] = float(params.pdf[i])
        p.rank, p.world = rank, world
        p.n_replicas = self.replicas = int(replicas)
        p.shard_lo, p.shard_hi = pop.lo, pop.lo + pop.N
        p.halo_lo, p.halo_hi = pop.halo
        self.hh_range = pop.hh_range or (0, pop.H)
        self.bld_range = pop.bld_range or (0, pop.B)
        self.sa_range = pop.sa_range or (0, pop.S)
        p.hh_lo, p.hh_hi = self.hh_range
        p.bld_lo, p.bld_hi = self.bld_range
        p.area_lo, p.area_hi = self.sa_range
        self.S_out, self.B_out = self.sa_range[1] - self.sa_range[0], self.bld_range[1] - self.bld_range[0]
        self.flags = flags
        self.h = C.c_void_p()
        rc = self.lib.dys_create(C.byref(p), device, C.byref(self.h))
        if rc != 0:
            raise DysError(rc, (self.lib.dys_last_error(None) or b"").decode())
        self.n_loc = pop.N
        self._o = _DayOutputs()
        self._o_ref = C.byref(self._o)
        self._views = {}
        self.load(pop)

    def _ck(self, rc):
        if rc != 0:
            raise DysError(rc, (self.lib.dys_last_error(self.h) or b"").decode())

    def load(self, pop):
        n = pop.N
        a = lambda x, dt, cnt=n: _ptr(np.ascontiguousarray(x, dtype=dt) if isinstance(x, np.ndarray) else x, None, cnt)
        keep = [a(pop.status, np.uint8), a(pop.t_inf, np.int32), a(pop.t_q, np.int32), a(pop.t_adm, np.int32),
                a(pop.src, np.int32), a(pop.hh, np.int32), a(pop.home, np.int32), a(pop.sa, np.int32),
                a(pop.risk, np.float64), a(pop.p_adm, np.float64), a(pop.p_death, np.float64),
                a(pop.routine, np.int32, (pop.halo[1] - pop.halo[0]) * pop.V)]
        self._ck(self.lib.dys_load_population(self.h, *[k[0] for k in keep]))
        g = [a(pop.ip_rowptr, np.int64, n + 1), a(pop.ip_col, np.int32, None), a(pop.ip_val, np.float64, None)]
        self._ck(self.lib.dys_load_graph(self.h, *[k[0] for k in g]))
        if pop.hh_always is not None and (pop.hh_always.any() or len(pop.trig_hh)):
            hl, hh_ = self.hh_range
            q = [a(pop.hh_always[hl:hh_], np.uint8, hh_ - hl), a(pop.trig_ptr, np.int64, n + 1), a(pop.trig_hh - hl, np.int32, None)]
            self._ck(self.lib.dys_load_quirk(self.h, *[k[0] for k in q]))
        self.set_open(pop.open)

    def set_open(self, open_flags):
        p, keep = _ptr(np.ascontiguousarray(open_flags, dtype=np.uint8) if isinstance(open_flags, np.ndarray) else open_flags,
                       None, self.pop.B)
        self._ck(self.lib.dys_set_open(self.h, p))

    def set_run(self, norm_factor=None, compliance=None, seed=None):
        pr = self.params
        if norm_factor is not None:
            pr.norm_factor = norm_factor
        if compliance is not None:
            pr.compliance = compliance
        if seed is not None:
            pr.seed = seed
        self._ck(self.lib.dys_set_run(self.h, pr.norm_factor, pr.compliance, pr.seed))

    def set_replica_params(self, norm_factor=None, compliance=None, seed=None):
        """per-replica run parameters (arrays of `replicas`); None keeps the current values"""
        arr = lambda x, dt: None if x is None else np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=dt), (self.replicas,)))
        nf, cp, sd = arr(norm_factor, np.float64), arr(compliance, np.float64), arr(seed, np.uint64)
        self._ck(self.lib.dys_set_replica_params(self.h, *[None if x is None else x.ctypes.data for x in (nf, cp, sd)]))

    def select_replica(self, r):
        """point the getters (outputs, stats, state, events ...) and set_state at replica r (-1: all / replica 0)"""
        self._ck(self.lib.dys_select_replica(self.h, int(r)))
        self._views = {}

    def set_lockdown(self, scenario_bits, bld_class=None, bld_area=None):
        """decide the open flags on the device from now on (dys_set_lockdown); scenario_bits = 0 switches it off.
        bld_area [B]: the per-area (DIFF) policy, each area on its own residents' R (dys_set_lockdown_areas)"""
        p = None
        if bld_class is not None:
            p, _keep = _ptr(np.ascontiguousarray(bld_class, dtype=np.uint8), None, self.pop.B)
        if bld_area is not None:
            a = np.ascontiguousarray(bld_area, dtype=np.int32)
            if a.shape != (self.pop.B,):
                raise ValueError("bld_area must have one entry per building")
            self._ck(self.lib.dys_set_lockdown_areas(self.h, int(scenario_bits), p, a.ctypes.data))
        else:
            self._ck(self.lib.dys_set_lockdown(self.h, int(scenario_bits), p))

    def abort(self):
        self._ck(self.lib.dys_abort(self.h))

    def h2d_bytes(self):
        return int(self.lib.dys_h2d_bytes(self.h))

    def set_state(self, status, t_inf, t_q, t_adm, src):
        n = self.n_loc
        ks = [_ptr(np.ascontiguousarray(status, np.uint8), None, n)] + [_ptr(np.ascontiguousarray(x, np.int32), None, n)
                                                                         for x in (t_inf, t_q, t_adm, src)]
        self._ck(self.lib.dys_set_state(self.h, *[k[0] for k in ks]))

    def _inj(self, u_adm, u_death, u_pair):
        if u_adm is None and u_death is None and u_pair is None:
            return None, None
        inj = _Injected()
        keep = []
        for name, x, cnt in (("u_adm", u_adm, self.N_glob), ("u_death", u_death, self.N_glob), ("u_pair", u_pair, self.N_glob ** 2)):
            if x is None:
                continue
            if isinstance(x, np.ndarray):
                x = np.ascontiguousarray(x, dtype=np.float64)
            p, k = _ptr(x, None, cnt)
            keep.append(k)
            setattr(inj, name, p)
        return inj, keep

    def step(self, day, u_adm=None, u_death=None, u_pair=None):
        inj, keep = self._inj(u_adm, u_death, u_pair)
        self._ck(self.lib.dys_step(self.h, int(day), C.byref(inj) if inj is not None else None))

    def step_many(self, first_day, n_days, open_flags=None):
        """queue n_days (<= 14; <= 7 to stay exact under lockdown scenarios) consecutive days in one call;
        open_flags: None or uint8 [n_days, B]"""
        p = None
        if open_flags is not None:
            p, _keep = _ptr(np.ascontiguousarray(open_flags, dtype=np.uint8) if isinstance(open_flags, np.ndarray) else open_flags,
                            None, n_days * self.pop.B)
        self._ck(self.lib.dys_step_many(self.h, int(first_day), int(n_days), p))

    def step_begin(self, day):
        self._ck(self.lib.dys_step_begin(self.h, int(day), None))

    def step_end(self, day):
        self._ck(self.lib.dys_step_end(self.h, int(day)))

    def outputs(self, day):
        """Everything day `day` produced, as zero-copy numpy views of the library's pinned host ring (valid until
        DYS_OUT_RING - 1 = 15 further days are stepped): dict(stats, sa_hist, building_counts, events[n,2], n_events)."""
        o = self._o
        rc = self.lib.dys_get_outputs(self.h, int(day), self._o_ref)
        if rc != 0:
            self._ck(rc)
        key = C.cast(o.stats, C.c_void_p).value          # the ring slot: its numpy views are built once
        views = self._views.get(key)
        if views is None:
            as_np = np.ctypeslib.as_array
            views = (as_np(o.stats, (N_STATS,)),
                     as_np(o.sa_hist, (self.S_out, HIST_LEN)) if o.sa_hist else None,
                     as_np(o.building_counts, (self.B_out,)) if o.building_counts else None,
                     as_np(o.events, (int(o.events_capacity), 2)) if o.events and o.events_capacity else None)
            self._views[key] = views
        n_ev, n_in = int(o.n_events), int(o.events_inline)
        ev = None
        if o.events:
            if n_ev > n_in:                                # rare: more infections than travel with the day's block
                a, s_ = self.events(day)
                ev = np.stack([a, s_], axis=1)
            elif views[3] is not None:
                ev = views[3][:n_in]                       # (a slice of the slot's cached view: no per-day ctypes cast)
            else:
                ev = np.zeros((0, 2), dtype=np.int32)
        return {"stats": views[0], "sa_hist": views[1], "building_counts": views[2], "events": ev, "n_events": n_ev}

    def consume_day(self, day, pdf_ptr, recover, stats_ptr, r_ptr, r_sa_ptr, bc_ptr, ev_ptr, ev_cap, n_ev_ptr):
        """dys_consume_day on raw addresses (the host loop pre-computes them: one ctypes call per simulated day)"""
        rc = self.lib.dys_consume_day(self.h, day, pdf_ptr, recover, stats_ptr, r_ptr, r_sa_ptr, bc_ptr, ev_ptr, ev_cap, n_ev_ptr)
        if rc != 0:
            self._ck(rc)

    def stats(self, day):
        out = np.zeros(N_STATS, dtype=np.int64)
        self._ck(self.lib.dys_get_stats(self.h, int(day), out.ctypes.data))
        return out

    def sa_hist(self, day):
        out = np.zeros((self.S_out, HIST_LEN), dtype=np.int32)
        self._ck(self.lib.dys_get_sa_hist(self.h, int(day), out.ctypes.data))
        return out

    def building_counts(self, day):
        out = np.zeros(self.B_out, dtype=np.int32)
        self._ck(self.lib.dys_get_building_counts(self.h, int(day), out.ctypes.data))
        return out

    def events(self, day):
        """today's (infected, infector) row indices, sorted by infected index (the device order is arbitrary)."""
        n = C.c_int64(0)
        self._ck(self.lib.dys_get_events(self.h, int(day), None, None, 0, C.byref(n)))
        a = np.zeros(n.value, dtype=np.int32)
        s = np.zeros(n.value, dtype=np.int32)
        if n.value:
            self._ck(self.lib.dys_get_events(self.h, int(day), a.ctypes.data, s.ctypes.data, n.value, C.byref(n)))
            o = np.argsort(a, kind="stable")
            a, s = a[o], s[o]
        return a, s

    def io_mat(self, day):
        out = np.zeros((self.pop.S, self.pop.S), dtype=np.int32)
        self._ck(self.lib.dys_get_io_mat(self.h, int(day), out.ctypes.data))
        return out

    def state(self):
        n = self.n_loc
        out = dict(status=np.zeros(n, np.uint8), t_inf=np.zeros(n, np.int32), t_q=np.zeros(n, np.int32),
                   t_adm=np.zeros(n, np.int32), src=np.zeros(n, np.int32))
        self._ck(self.lib.dys_get_state(self.h, *[out[k].ctypes.data for k in ("status", "t_inf", "t_q", "t_adm", "src")]))
        return out

    def pair_prob(self):
        out = np.zeros((self.N_glob, self.N_glob), dtype=np.float64)
        self._ck(self.lib.dys_get_pair_prob(self.h, out.ctypes.data))
        return out

    def sync(self):
        self._ck(self.lib.dys_sync(self.h))

    def kernel_times(self):
        ms = np.zeros(N_KERNELS, dtype=np.float64)
        cnt = np.zeros(N_KERNELS, dtype=np.int64)
        self._ck(self.lib.dys_get_kernel_times(self.h, ms.ctypes.data, cnt.ctypes.data))
        return ms, cnt

    def device_buffer(self, which):
        p, b = C.c_void_p(), C.c_int64()
        self._ck(self.lib.dys_device_buffer(self.h, which, C.byref(p), C.byref(b)))
        return p.value, b.value

    def peer_export(self):
        """CUDA IPC handles (bytes) of this rank's snapshot buffer and flag array"""
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._ck(self.lib.dys_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, handles, halos):
        """handles: every rank's peer_export() bytes, halos: every rank's (shard_lo, shard_hi, halo_lo, halo_hi)"""
        blob = b"".join(handles)
        hal = np.ascontiguousarray(halos, dtype=np.int64).reshape(len(handles), 4)
        self._ck(self.lib.dys_peer_attach(self.h, len(handles), blob, hal.ctypes.data))

    def peer_detach(self):
        self._ck(self.lib.dys_peer_detach(self.h))

    def cuda_stream(self):
        p = C.c_void_p()
        self._ck(self.lib.dys_cuda_stream(self.h, C.byref(p)))
        return p.value

    def device_bytes(self):
        return int(self.lib.dys_device_bytes(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.dys_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
