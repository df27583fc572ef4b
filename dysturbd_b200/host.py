"""Host side of the day loop: the reference's call surface over the device library.

`EpidemicModel` takes exactly what contagious3.py has in hand when its `while` starts (agents, build,
bld_visits_by_agents, interaction_prob, scenario -- contagious3.py:114-174), runs each day on the GPU through the
C ABI (`Engine.step` == one pass of contagious3.py:176-366) and rebuilds the reference's own `outputs` dict
(contagious3.py:111, 326-366, 389) from the integer arrays the device emits.  Every float in that dict (R, Known R,
building fractions) is formed here, on the host, with the reference's order of operations (compute_R,
contagious3.py:29-42), so results compare with `==`.  Phase H (building_lockdown, contagious3.py:57-107, 368-385)
is O(B) bookkeeping and stays on the host.
"""
import numpy as np

from . import engine as _engine
from .engine import HIST0, HIST_LEN, STAT, Engine
from .population import EpiParams, Population

EDU_CODES = (5310, 5312, 5338, 5523, 5525, 5305, 5300, 5340)     # contagious3.py:67-70
REL_CODES = (5501, 5521)                                         # contagious3.py:74


def compute_R_from_hist(hist, pdf, recover):
    """compute_R (contagious3.py:29-42) on a days-since-infection histogram: hist[0] = today's infections,
    hist[i] = #agents ever infected with day - t_inf == i.  Same accumulation order, python floats."""
    new_infections = int(hist[0])
    sum_I = 0.
    for i in range(1, recover):
        sum_I += int(hist[i]) * pdf[i]
    return new_infections / sum_I if sum_I > 0 else 0


def compute_R_rows(sa_hist, pdf, recover):
    """compute_R_from_hist for every row of a [S, 21] histogram at once, bit for bit: np.cumsum accumulates sequentially
    (np.sum would not: it sums pairwise), which is the reference's `sum_I += n * pdf(i)` order; 0.0 + x is exact."""
    prod = sa_hist[:, 1:recover].astype(np.float64) * np.asarray(pdf[1:recover], dtype=np.float64)
    sum_I = np.cumsum(prod, axis=1)[:, -1] if prod.shape[1] else np.zeros(len(sa_hist))
    new = sa_hist[:, 0].astype(np.float64)
    return np.where(sum_I > 0, new / np.where(sum_I > 0, sum_I, 1.0), 0.0)


def building_classes(lu, usg):
    """the three index sets building_lockdown works on (contagious3.py:63, 67-70, 74): EDU codes, REL codes, land use >= 3"""
    return np.where(np.isin(usg, EDU_CODES))[0], np.where(np.isin(usg, REL_CODES))[0], np.where(lu >= 3)[0]


def building_lockdown(lu, usg, current, sc, vR, prevVR, choice, keyed=None, classes=None, dtype=float, ids=None, cache=None):
    """building_lockdown(b, sc, vR, prevVR) of contagious3.py:57-107 on the two building columns it reads
    (b[:,1] land use, b[:,9] USG code) and b[:,10] (`current`).  `choice` stands for np.random.choice.
    keyed = (seed, day): the half closures of GRADUAL close each building of the class with probability 1/2 on the keyed
    draw u(day, STREAM_LOCKDOWN, building) instead -- the host twin of the device policy (dys_set_lockdown).
    ids: the buildings' indices in the whole table when lu / usg / current are the rows of ONE area (the DIFF scenarios call
    this once per area, contagious3.py:368-376); the keyed draws are keyed on those.
    cache (a dict the caller keeps for ONE building table and scenario): the all-open vector and the vector of a full
    closure never change, so they are formed once and handed out again -- treat the result as read-only."""
    if cache is not None:
        # which branch of :57-107 applies is decided by the two R values alone
        if 'GRADUAL' in sc:
            what = ('half' if (prevVR <= 1 or prevVR >= 2) else 'keep') if 1 < vR < 2 else ('full' if vR >= 2 else 'open')
        else:
            what = 'full' if 1 < vR else 'open'
        if what == 'keep':
            return current if isinstance(current, np.ndarray) and current.dtype == dtype else np.array(current, dtype=dtype)
        if what in ('open', 'full'):
            if what not in cache:
                cache[what] = building_lockdown(lu, usg, current, sc, 0 if what == 'open' else 3, 0, choice, classes=classes, dtype=dtype)
                cache[what].setflags(write=False)
            return cache[what]
    lockdown = np.ones(len(lu), dtype=dtype)      # (the reference's float64 vector; the model keeps its flags as bytes)
    edu, rel, non_res = classes if classes is not None else building_classes(lu, usg)

    def close(idx, half):
        if half and keyed is not None:
            from .philox_host import STREAM_LOCKDOWN, keyed_uniform
            lockdown[idx[keyed_uniform(keyed[0], keyed[1], STREAM_LOCKDOWN, idx if ids is None else ids[idx]) < 0.5]] = 0
        elif half:
            lockdown[choice(idx, replace=False, size=int(idx.size * 0.5))] = 0
        else:
            lockdown[idx] = 0

    def apply(half):
        if 'ALL' in sc:
            close(non_res, half)
        else:
            if 'EDU' in sc:
                close(edu, half)
            if 'REL' in sc:
                close(rel, half)

    if 'GRADUAL' in sc:
        if 1 < vR < 2:
            if prevVR <= 1 or prevVR >= 2:
                apply(True)
            else:
                lockdown = np.array(current, dtype=dtype)
        elif vR >= 2:
            apply(False)
    elif 1 < vR:
        apply(False)
    return lockdown


class EpidemicModel:
    """contagious3.py:172-389 with the day on the GPU.

    run() queues days in windows (one cooperative launch each) and keeps two windows in flight.  The open flags of day
    d+1 depend on `Known R` of days d and d-1, i.e. on R of days d-`diagnosis` and d-`diagnosis`-1 (contagious3.py:306-309,
    379-384): with c days consumed, the flags of every day <= c + diagnosis are already determined.  Two windows of
    (diagnosis + 1) // 2 days therefore never need a day that has not been consumed -- exact for every scenario, while the
    host's dictionary building overlaps the device's work.  A scenario that closes nothing (no ALL / EDU / REL keyword)
    has constant flags and uses week-long windows."""
    MAX_WINDOW = 7

    def __init__(self, agents, build, bld_visits_by_agents, interaction_prob, scenario=('noLockdown',), params=None,
                 device=0, choice_seed=None, backend=None, population=None, compact=False, reduce_stats=None,
                 upload_flags=True, keep_raw=False, lockdown='host', diff_key_fix=False):
        self.params = params or EpiParams()
        self.pop = population or Population.from_reference(agents, build, bld_visits_by_agents, interaction_prob,
                                                                   diagnosis=self.params.diagnosis)
        self.scenario = list(scenario)
        self.sim = backend(self.pop, self.params) if backend is not None else Engine(self.pop, self.params, device=device)
        # compact=True: 'SAs', 'Buildings' and 'IO_mat' hold numpy arrays instead of the reference's nested dicts and
        # lists (same numbers: per-area R [S], vis_R [S]; per inhabited building count and fraction [n, 2]; the day's
        # (infected, infector) area indices [n_events, 2]).  The reference's schema costs O(S^2 + B) Python objects per
        # day -- fine for its 19 areas, 50 ms per day for a million agents.  'Stats' is always the reference's dict.
        self.compact = bool(compact)
        # sharded populations (one EpidemicModel per rank over a ShardedEngine): `reduce_stats` sums a [days, 64] block of
        # stats over the ranks; R, Known R and the lockdown decision then use the GLOBAL counts while the per-area and
        # per-building outputs stay local to the rank's areas / buildings
        self._reduce = reduce_stats
        self._gstats = {}
        self._upload = bool(upload_flags)
        # keep_raw: every day's raw 64-integer stats block (work counters included) and the flag vector it ran under
        self.raw_stats = {} if keep_raw else None
        self.flags_log = {} if keep_raw else None
        self.day = 0                 # next day to consume
        self.outputs = {'Results': {'Stats': {}, 'Buildings': {}, 'SAs': {}, 'IO_mat': {}}}
        self._choice = np.random.RandomState(choice_seed).choice
        p = self.pop
        self._bld_pop = np.bincount(p.home, minlength=p.B)
        self._inhabited = np.nonzero(self._bld_pop > 0)[0]
        self._b0 = p.bld_range[0] if p.bld_range else 0          # outputs cover the owned buildings / areas only
        self._s0, self._s1 = p.sa_range if p.sa_range else (0, p.S)
        self._pdf = [float(x) for x in self.params.pdf]
        self._pdf_arr = np.ascontiguousarray(self.params.pdf, dtype=np.float64)
        self._n_closed_cache = {}
        self._ck, self._ck_base = None, 0
        self._ev_arena, self._ev_used = None, 0
        self._ev_room = max(1024, (self.pop.N + 1023) // 1024 * 1024)      # a day cannot infect more agents than there are
        try:                                     # the library's host-side helper (exact); numpy otherwise (also exact)
            _engine.load_library()
            self._R_rows = _engine.compute_R_rows
        except Exception:
            self._R_rows = lambda h, pdf, rec: compute_R_rows(np.atleast_2d(np.asarray(h)), pdf, rec)
        self._active = int(np.isin(p.status, (4, 7, 8, 10)).sum())
        if self._reduce is not None:
            self._active = int(self._reduce(np.array([[self._active] + [0] * 63], dtype=np.int64))[0, 0])
        self._R = {}                 # day -> R of that day (contagious3.py:281)
        # DIFF (per-area lockdown, contagious3.py:368-376).  The reference reads SAs[day-1]['Vis_R'] there but writes that
        # dict under 'vis_R' (:319): every DIFF run ends in KeyError on day 1, and so does this class by default.
        # diff_key_fix=True reads the key that is written -- the per-area policy the code evidently meant (pinned by the
        # '*_keyfix' golden fixtures, made from the reference's own text with that one key repaired at exec time).
        self._diff = 'DIFF' in self.scenario
        self._diff_fix = bool(diff_key_fix)
        self._Rsa = {}               # day -> per-area R of that day [S] (kept for the DIFF decision only)
        if self._diff and self._diff_fix:
            if self._reduce is not None:
                raise NotImplementedError("DIFF decides per area from the rank's own areas: not wired for sharded populations")
            area = np.maximum(np.asarray(p.bld_sa, dtype=np.int64), 0)
            order = np.argsort(area, kind="stable")
            cuts = np.searchsorted(area[order], np.arange(p.S + 1))
            self._area_blds = [order[cuts[k]:cuts[k + 1]] for k in range(p.S)]          # building rows of each area, ascending
            self._bld_area = area
        self._flags = {0: self.pop.open.astype(np.uint8)}   # day -> build[:,10] in force on that day (0 / 1 bytes)
        self._queued = 0             # next day to queue
        closes = self._closes = any(k in self.scenario for k in ('ALL', 'EDU', 'REL', 'DIFF'))
        self._ones = np.ones(self.pop.B, dtype=np.uint8)
        self._open_window = None
        self._classes = building_classes(self.pop.bld_lu, self.pop.bld_usg)      # (static: not recomputed every day)
        self._ld_cache = {}          # the all-open and the fully closed vector of this scenario (building_lockdown, cache=)
        self.window = max(1, min(self.MAX_WINDOW, (self.params.diagnosis + 1) // 2)) if closes else self.MAX_WINDOW
        # lockdown = 'host'   phase H here, with the reference's np.random.choice (exact for every scenario)
        #            'keyed'  phase H here, GRADUAL half closures on keyed draws (the host twin of the device policy)
        #            'device' phase H inside the day kernel (dys_set_lockdown): no flags travel, week-long windows; exact
        #                     for the deterministic scenarios, keyed half closures for GRADUAL
        self.lockdown = lockdown
        if lockdown == 'device':
            bits = sum(b for k, b in (('GRADUAL', _engine.SC_GRADUAL), ('ALL', _engine.SC_ALL), ('EDU', _engine.SC_EDU),
                                      ('REL', _engine.SC_REL)) if k in self.scenario)
            cls = (np.where(p.bld_lu >= 3, _engine.CLASS_NONRES, 0) | np.where(np.isin(p.bld_usg, EDU_CODES), _engine.CLASS_EDU, 0) |
                   np.where(np.isin(p.bld_usg, REL_CODES), _engine.CLASS_REL, 0)).astype(np.uint8)
            if (bits & ~_engine.SC_GRADUAL) and (not self._diff or self._diff_fix):   # (DIFF as shipped: the host path mirrors the reference's KeyError)
                self.sim.set_lockdown(bits, cls, self._bld_area.astype(np.int32) if self._diff else None)
                self.window = self.MAX_WINDOW
            else:
                self.lockdown = 'host'           # a scenario that closes nothing

    @property
    def open(self):
        """build[:, 10] as the reference leaves it after the last consumed day (= the flags of the next day)"""
        return self._flags_for(self.day)

    def active(self):
        """the `while` condition of contagious3.py:176"""
        return self._active > 0

    def _known_R(self, day):
        d = self.params.diagnosis
        return self._R[day - d] if day > d else 0                                               # :306-309

    def _flags_for(self, day):
        """phase H (contagious3.py:368-385): the flags of `day` were decided at the end of day-1 from Known R"""
        if day not in self._flags:
            prev = day - 1
            if not self._closes:
                # no ALL / EDU / REL keyword: building_lockdown closes nothing whatever R is (contagious3.py:57-107), so the
                # flags of every day after day 0 are all ones -- known without waiting for R (week-long windows)
                self._flags[day] = self._ones
            elif self._diff and not self._diff_fix:
                if prev != 0:
                    raise KeyError('Vis_R')      # the reference reads a key it never writes (:371 vs :319)
                self._flags[day] = self._ones
            elif self._diff:
                self._flags[day] = self._diff_flags(prev)
            else:
                cur = self._flags_for(prev)
                prevVR = self._known_R(prev - 1) if prev != 0 else 0
                self._flags[day] = building_lockdown(self.pop.bld_lu, self.pop.bld_usg, cur, self.scenario,
                                                     self._known_R(prev), prevVR, self._choice,
                                                     keyed=(self.params.seed, prev) if self.lockdown == 'keyed' else None,
                                                     classes=self._classes, dtype=np.uint8, cache=self._ld_cache)
            self._flags.pop(day - 20, None)
        return self._flags[day]

    def _sa_known_R(self, day):
        d = self.params.diagnosis
        return self._Rsa[day - d] if day > d else np.zeros(self.pop.S)                          # :311-318

    def _diff_flags(self, prev):
        """contagious3.py:368-376 (with the key repaired): building_lockdown once per area, in ascending area order, on the
        area's own rows with the area's visible R of `prev` and of the day before.  Areas whose visible R is <= 1 are all
        open and draw nothing, so only the others are visited (same np.random.choice sequence as the reference)."""
        p, sc = self.pop, self.scenario
        vR = self._sa_known_R(prev)
        pv = self._sa_known_R(prev - 1) if prev != 0 else np.zeros(p.S)
        cur = self._flags_for(prev)
        out = np.ones(p.B, dtype=np.uint8)
        keyed = (self.params.seed, prev) if self.lockdown == 'keyed' else None
        for s in np.nonzero(vR > 1)[0]:
            m = self._area_blds[s]
            out[m] = building_lockdown(p.bld_lu[m], p.bld_usg[m], cur[m], sc, float(vR[s]), float(pv[s]), self._choice,
                                       keyed=keyed, dtype=np.uint8, ids=m)
        return out

    def _queue(self, n):
        first = self._queued
        if self.lockdown == 'device':            # the flags are decided inside the day kernel
            self.sim.step_many(first, n, None)
            self._queued += n
            return
        fl = [self._flags_for(first + k) for k in range(n)]
        if all(f is self._ones for f in fl):     # (a scenario that closes nothing: the same all-open window every time)
            if self._open_window is None or len(self._open_window) < n:
                self._open_window = np.ones((self.MAX_WINDOW, self.pop.B), dtype=np.uint8)
            flags = self._open_window[:n]
        else:
            flags = np.stack(fl)
        if self.flags_log is not None:
            for k in range(n):
                self.flags_log[first + k] = flags[k]
        if hasattr(self.sim, "step_many"):
            # the day's building flags are this path's only per-day host input (contagious3.py:384 re-assigns build[:,10]
            # every day); upload_flags=False offers them only when they changed
            changed = self._upload or any(not np.array_equal(flags[k], self._last_flags) for k in range(n)) \
                if hasattr(self, "_last_flags") else True
            self.sim.step_many(first, n, flags if changed else None)
            self._last_flags = flags[-1]
        else:                                    # tests drive this loop with the CPU oracle as backend
            assert n == 1
            self.sim.set_open(flags[0])
            self.sim.step(first)
        self._queued += n

    def _consume(self):
        if self.compact:
            return self._consume_compact()
        p, sim, day, prm = self.pop, self.sim, self.day, self.params
        res = self.outputs['Results']
        st, sa_hist, bc, ev_a, ev_s = self._day_arrays(day, sort_events=True)
        closed_today = int(st[_engine.ST_CLOSED]) if self.lockdown == 'device' else (self.pop.B - int(np.count_nonzero(self._flags_for(day))))
        hist = st[HIST0:HIST0 + HIST_LEN]
        R = compute_R_from_hist(hist, self._pdf, prm.recover)
        self._R[day] = R
        vis_R = self._known_R(day)
        sas = [float(x) for x in p.sa_ids]
        res['SAs'][day] = {'R': {sa: compute_R_from_hist(sa_hist[k], self._pdf, prm.recover) for k, sa in enumerate(sas)}}
        res['SAs'][day]['vis_R'] = {sa: (res['SAs'][day - prm.diagnosis]['R'][sa] if day > prm.diagnosis else 0) for sa in sas}
        if self._diff:
            self._Rsa[day] = np.array([res['SAs'][day]['R'][sa] for sa in sas], dtype=np.float64)
            self._Rsa.pop(day - 2 * prm.diagnosis - 4, None)
        S = self._stats_dict(st, R, vis_R, closed_today)
        res['Stats'][day] = S
        res['Buildings'][day] = [[float(p.bld_ids[b]), int(bc[b]), int(bc[b]) / int(self._bld_pop[b])] for b in self._inhabited]
        mat = {sa: {sa2: 0 for sa2 in sas} for sa in sas}                                       # :356-366
        if len(ev_a):
            for a, s in zip(p.sa[ev_a], p.sa[ev_s]):
                mat[sas[a]][sas[s]] += 1
        res['IO_mat'][day] = mat
        self._active = S['Active infected']
        self.day += 1
        return S

    def _day_arrays(self, day, sort_events):
        """the day's integer outputs: (stats[64], sa_hist[S,21], building_counts[B], infected agents, their infectors)"""
        sim = self.sim
        if hasattr(sim, "outputs"):              # the CUDA engine: one pinned block per day, no further transfers
            o = sim.outputs(day)
            st, sa_hist, bc = o["stats"], o["sa_hist"], o["building_counts"]
            if self._reduce is not None:         # sharded: the counts behind R / Stats are sums over the ranks, one
                if day not in self._gstats:      # all-reduce per window of queued days
                    days = list(range(day, min(self._queued, day + self.window)))
                    g = self._reduce(np.stack([sim.outputs(d)["stats"] for d in days]))
                    self._gstats = dict(zip(days, g))
                st = self._gstats.pop(day)
            ev = self._ev_pairs = o["events"]
            if sort_events:
                order = np.argsort(ev[:, 0], kind="stable")
                return st, sa_hist, bc, ev[order, 0], ev[order, 1]
            return st, sa_hist, bc, ev[:, 0], ev[:, 1]
        ev_a, ev_s = sim.events(day)
        self._ev_pairs = None
        return sim.stats(day), sim.sa_hist(day), sim.building_counts(day), ev_a, ev_s

    def _stats_dict(self, st, R, vis_R, closed_today):
        day, res = self.day, self.outputs['Results']
        if self.raw_stats is not None:
            self.raw_stats[day] = np.array(st)
        v = st[:9].tolist()
        new_infections = v[STAT['daily_inf']]
        S = {'Active infected': v[0], 'Daily infections': new_infections, 'Recovered': v[2], 'Quarantined': v[3],
             'Daily quarantined': v[4], 'Hospitalized': v[5], 'Daily hospitalizations': v[6], 'Total Dead': v[7],
             'Daily deaths': v[8], 'R': R, 'Known R': vis_R, 'Closed buildings': closed_today}
        if day == 0:
            S['Total infected'] = S['Active infected']                                          # :347-348
        else:
            S['Total infected'] = res['Stats'][day - 1]['Total infected'] + new_infections
        S['Susceptible'] = self.pop.N_glob - S['Total infected']
        return S

    CHUNK = 64                   # days per block of preallocated compact outputs

    def _new_chunk(self):
        n, S, B = self.CHUNK, self._s1 - self._s0, self.sim.B_out
        ck = dict(st=np.empty((n, 64), np.int64), R=np.empty(n, np.float64), Rsa=np.empty((n, S), np.float64),
                  bc=np.empty((n, B), np.int32), nev=np.zeros(1, np.int64))
        ck["ptr"] = {k: v.ctypes.data for k, v in ck.items()}
        self._ck_base, self._ck = self.day, ck

    def _consume_compact(self):
        """compact=True: the same numbers as arrays -- 'SAs'[day] = {'R': float64[S], 'vis_R': float64[S]} (owned areas),
        'Buildings'[day] = int32[B] infected residents per owned building (buildings_table() gives the reference's
        [count, fraction] rows), 'IO_mat'[day] = int32[n, 2] (infected, infector) agent rows (io_pairs() gives the area
        pairs the reference counts).  On the CUDA engine a day is ONE library call (dys_consume_day: copies out of the pinned
        ring + compute_R in the reference's order of float operations) into preallocated blocks of 64 days."""
        p, day, prm, sim = self.pop, self.day, self.params, self.sim
        res = self.outputs['Results']
        fast = self._reduce is None and hasattr(sim, "consume_day")
        if fast:
            i = day - self._ck_base
            if self._ck is None or i >= self.CHUNK:
                self._new_chunk()
                i = 0
            ck, ptr = self._ck, self._ck["ptr"]
            if self._ev_arena is None or self._ev_used + self._ev_room > len(self._ev_arena):
                self._ev_arena, self._ev_used = np.empty((max(1 << 20, 4 * self._ev_room), 2), np.int32), 0
            S = ck["Rsa"].shape[1]
            sim.consume_day(day, self._pdf_arr.ctypes.data, prm.recover, ptr["st"] + i * 512, ptr["R"] + i * 8, ptr["Rsa"] + i * 8 * S,
                            ptr["bc"] + i * 4 * ck["bc"].shape[1], self._ev_arena.ctypes.data + self._ev_used * 8, self._ev_room, ptr["nev"])
            st, R, R_sa, bc = ck["st"][i], float(ck["R"][i]), ck["Rsa"][i], ck["bc"][i]
            n_ev = int(ck["nev"][0])
            ev_pairs = self._ev_arena[self._ev_used:self._ev_used + n_ev]
            self._ev_used += n_ev
        else:
            st, sa_hist, bc, ev_a, ev_s = self._day_arrays(day, sort_events=False)
            ev = self._ev_pairs
            R = float(self._R_rows(st[HIST0:HIST0 + HIST_LEN], self._pdf_arr, prm.recover)[0])
            R_sa = self._R_rows(np.asarray(sa_hist)[:self._s1 - self._s0], self._pdf_arr, prm.recover)
            bc = np.array(bc, dtype=np.int32)
            ev_pairs = np.array(ev, dtype=np.int32) if ev is not None else np.stack([ev_a, ev_s], axis=1).astype(np.int32)
        if self.lockdown == 'device':
            closed_today = int(st[_engine.ST_CLOSED])
        else:
            flags = self._flags_for(day)
            closed_today = self._n_closed_cache.get(id(flags))
            if closed_today is None:
                closed_today = len(flags) - int(np.count_nonzero(flags))
                self._n_closed_cache = {id(flags): closed_today}
        self._R[day] = R
        dz = prm.diagnosis
        res['SAs'][day] = {'R': R_sa, 'vis_R': res['SAs'][day - dz]['R'] if day > dz else np.zeros(len(R_sa))}
        if self._diff:
            self._Rsa[day] = R_sa
            self._Rsa.pop(day - 2 * dz - 4, None)
        S = self._stats_dict(st, R, self._known_R(day), closed_today)
        res['Stats'][day] = S
        res['Buildings'][day] = bc
        res['IO_mat'][day] = ev_pairs
        self._active = S['Active infected']
        self.day += 1
        return S

    def buildings_table(self, day):
        """compact mode: the reference's 'Buildings'[day] rows (contagious3.py:339-346) as [n_inhabited, 2] = infected
        residents, fraction of the building's residents, in build[] order"""
        cnt = self.outputs['Results']['Buildings'][day][self._inhabited - self._b0].astype(np.float64)
        return np.stack([cnt, cnt / self._bld_pop[self._inhabited]], axis=1)

    def io_pairs(self, day):
        """compact mode: (area of the infected agent, area of its infector) of the day's infections -- what the reference's
        IO_mat counts (contagious3.py:356-366); single-GPU populations only (an infector may live on another rank)"""
        ev = self.outputs['Results']['IO_mat'][day]
        return np.stack([self.pop.sa[ev[:, 0]], self.pop.sa[ev[:, 1]]], axis=1) if len(ev) else np.zeros((0, 2), dtype=self.pop.sa.dtype)

    # ---- the reference's output file and a resumable snapshot ----
    def write_json(self, path, total_time=None):
        """outputs as contagious3.py:389-392 writes them (`json.dump(outputs)`; the day and area keys become strings
        exactly as they do there), so create_figures.py reads a run of this library unchanged"""
        import json
        if self.compact:
            raise ValueError("write_json writes the reference's schema: run the model with compact=False")
        out = dict(self.outputs)
        if total_time is not None:
            out['Total_time'] = float(total_time)                                                # :390
        with open(path, 'w') as f:
            json.dump(out, f)

    def checkpoint(self):
        """everything needed to continue this run in another process: the agent columns the day mutates, the host's
        R history (the lockdown decision lags `diagnosis` days), the flags in force, the np.random.choice stream and the
        outputs so far.  Only valid between days (nothing queued ahead of the consumed day)."""
        import copy
        if self._queued != self.day:
            raise RuntimeError("checkpoint() with days in flight: call it from step(), not inside run()")
        if hasattr(self.sim, "sync"):
            self.sim.sync()
        return {'day': self.day, 'state': {k: np.array(v) for k, v in self.sim.state().items()}, 'R': dict(self._R),
                'R_areas': {d: np.array(v) for d, v in self._Rsa.items()}, 'flags': {d: np.array(f) for d, f in self._flags.items()}, 'active': self._active,
                'choice_state': self._choice.__self__.get_state(), 'outputs': copy.deepcopy(self.outputs),
                'scenario': list(self.scenario)}

    def restore(self, ck):
        """continue from checkpoint(): the population (static columns, network) must be the one the run started with"""
        import copy
        if list(ck['scenario']) != self.scenario:
            raise ValueError("checkpoint of scenario %s restored into %s" % (ck['scenario'], self.scenario))
        st = ck['state']
        self.sim.set_state(st['status'], st['t_inf'], st['t_q'], st['t_adm'], st['src'])
        self.day = self._queued = int(ck['day'])
        self._R = dict(ck['R'])
        self._Rsa = {d: np.array(v) for d, v in ck.get('R_areas', {}).items()}
        self._flags = {d: np.array(f) for d, f in ck['flags'].items()}
        self._active = int(ck['active'])
        self._choice.__self__.set_state(ck['choice_state'])
        self.outputs = copy.deepcopy(ck['outputs'])
        if hasattr(self.sim, "set_open"):
            self.sim.set_open(self._flags_for(self.day).astype(np.uint8))

    def step(self):
        """one day, queued and consumed (contagious3.py:176-387)"""
        if self._queued == self.day:
            self._queue(1)
        S = self._consume()
        if self.lockdown != 'device':
            self._flags_for(self.day)    # phase H for tomorrow, in the reference's order of np.random.choice calls
        return S

    def run(self, max_days=None, finalize=True):
        """the reference's `while` (contagious3.py:176); finalize=False skips the once-per-run 'Infections chain' read-back
        (contagious3.py:389) so that a caller can continue the run"""
        batched = hasattr(self.sim, "step_many")
        w = self.window if batched else 1
        limit = (lambda d: max_days is None or d < max_days)
        while self.active() and limit(self.day):
            while batched and self._queued < self.day + 2 * w and limit(self._queued):      # keep two windows in flight
                n = w if max_days is None else min(w, max_days - self._queued)
                self._queue(n)
            if not batched:
                self._queue(1)
            for _ in range(self._queued - self.day if not batched else min(w, self._queued - self.day)):
                if not (self.active() and limit(self.day)):
                    break
                self._consume()
        # days queued beyond the end of the epidemic changed nothing the reference records (no infectious agent is
        # left, so columns 16 and 21 are final), and their outputs are dropped
        if hasattr(self.sim, "sync"):
            self.sim.sync()
        if not finalize:
            return self.outputs
        st = self.sim.state()
        src_id = np.where(st['src'] >= 0, self.pop.agent_ids[np.maximum(st['src'], 0)], 0.0)
        self.outputs['Results']['Infections chain'] = np.stack(
            [self.pop.agent_ids, st['t_inf'].astype(float), src_id], axis=1).tolist()          # :389
        return self.outputs
