"""Host-side data model: the reference's float64 matrices  <->  the structure-of-arrays the device library holds.

The reference keeps everything in `agents[N,29]` float64 (create_random_data.py:112-142; column map in
contagious3.py:115-118), `build[B,15]` (create_random_data.py:28-30), the padded routine table
`bld_visits_by_agents[N,V]` (contagious3.py:139-170) and the dense `interaction_prob[N,N]` (contagious3.py:129).
`Population.from_reference` converts those into dense indices + compact integer state + CSR, which is what
`include/dysturbd.h` takes; `write_back` puts the state columns back into an agents matrix.
"""
from dataclasses import dataclass, field

import numpy as np

# reference status code (agents[:,13]) * 2  -> u8
ST_S, ST_I, ST_QH, ST_QI, ST_D, ST_H, ST_R, ST_X = 2, 4, 6, 7, 8, 10, 12, 14

# agents[:, c] columns that the day loop mutates (contagious3.py:192-270); every other column is static
DYNAMIC_COLS = (13, 16, 17, 19, 20, 21, 24, 25)


@dataclass
class EpiParams:
    """parameters.py:41-46 plus the literals of contagious3.py:205 (admission window on the absolute infection
    day) and :219 (minimum admission day for mortality)."""
    norm_factor: float = 0.08
    recover: int = 21
    hospital_recover: int = 28
    diagnosis: int = 7
    quarantine: int = 7
    adm_min_day: int = 4
    adm_max_day: int = 14
    death_min_adm_day: int = 3
    compliance: float = 1.0          # extension (BASELINE.json config 5); 1.0 == reference behaviour
    seed: int = 0
    pdf: np.ndarray = field(default=None)

    def __post_init__(self):
        if self.pdf is None:
            self.pdf = default_pdf_table()
        self.pdf = np.ascontiguousarray(self.pdf, dtype=np.float64)
        assert self.pdf.shape == (64,)


def default_pdf_table():
    """contagious_risk_day.pdf(0..63) with contagious_risk_day = gamma((4.5/3.5)**2, scale=3.5**2/4.5)
    (parameters.py:44), computed on the host with scipy exactly as the reference does; the device only indexes it."""
    from scipy.stats import gamma
    return gamma((4.5 / 3.5) ** 2, scale=3.5 ** 2 / 4.5).pdf(np.arange(64.0))


def _index_of(sorted_keys, order, values, what):
    pos = np.searchsorted(sorted_keys, values)
    pos = np.clip(pos, 0, len(sorted_keys) - 1)
    ok = sorted_keys[pos] == values
    if not ok.all():
        raise ValueError("%s: %d values not found" % (what, int((~ok).sum())))
    return order[pos].astype(np.int32)


@dataclass
class Population:
    N: int
    H: int
    B: int
    S: int
    V: int
    # mutable state
    status: np.ndarray      # u8  [N]   reference code * 2
    t_inf: np.ndarray       # i32 [N]   agents[:,16]
    t_q: np.ndarray         # i32 [N]   agents[:,19]
    t_adm: np.ndarray       # i32 [N]   agents[:,24]
    src: np.ndarray         # i32 [N]   row index of agents[:,21]'s id, -1 = none
    # static per agent
    hh: np.ndarray          # i32 [N]
    home: np.ndarray        # i32 [N]
    sa: np.ndarray          # i32 [N]   (-1: home area not in build[:,3])
    risk: np.ndarray        # f64 [N]   agents[:,14]
    p_adm: np.ndarray       # f64 [N]   agents[:,23]
    p_death: np.ndarray     # f64 [N]   agents[:,26]
    routine: np.ndarray     # i32 [N,V] building index, -1 = empty slot
    # interaction_prob as CSR (columns ascending within a row)
    ip_rowptr: np.ndarray   # i64 [N+1]
    ip_col: np.ndarray      # i32 [nnz]
    ip_val: np.ndarray      # f64 [nnz]
    open: np.ndarray        # u8  [B]   build[:,10]
    # household-isin quirk (contagious3.py:267)
    hh_always: np.ndarray   # u8  [H]
    trig_ptr: np.ndarray    # i64 [N+1]
    trig_hh: np.ndarray     # i32 [ntrig]
    # ids for the reference-facing outputs
    agent_ids: np.ndarray = None
    hh_ids: np.ndarray = None
    bld_ids: np.ndarray = None
    sa_ids: np.ndarray = None
    bld_lu: np.ndarray = None      # build[:,1]
    bld_usg: np.ndarray = None     # build[:,9]
    bld_sa: np.ndarray = None      # i32 [B] index into sa_ids
    # sharded populations (synth.py): this object holds agents [lo, lo+N) of N_glob; `routine` covers `halo`
    N_glob: int = None
    lo: int = 0
    halo: tuple = None
    hh_range: tuple = None         # owned households / buildings / areas (index ranges); None = all
    bld_range: tuple = None
    sa_range: tuple = None
    age: np.ndarray = None

    def __post_init__(self):
        if self.N_glob is None:
            self.N_glob = self.N
        if self.halo is None:
            self.halo = (0, self.N_glob)

    @property
    def nnz(self):
        return int(self.ip_rowptr[-1])

    def copy_state(self):
        return {k: getattr(self, k).copy() for k in ("status", "t_inf", "t_q", "t_adm", "src")}

    @staticmethod
    def from_reference(agents, build, bld_visits_by_agents, interaction_prob, diagnosis=7):
        agents = np.asarray(agents, dtype=np.float64)
        build = np.asarray(build, dtype=np.float64)
        N, B = agents.shape[0], build.shape[0]
        agent_ids = agents[:, 0].copy()
        bld_ids = build[:, 0].copy()
        if len(np.unique(bld_ids)) != B or len(np.unique(agent_ids)) != N:
            raise ValueError("agent ids and building ids must be unique")
        bo = np.argsort(bld_ids, kind="stable")
        ao = np.argsort(agent_ids, kind="stable")
        hh_ids, hh = np.unique(agents[:, 1], return_inverse=True)
        home = _index_of(bld_ids[bo], bo, agents[:, 7], "home building (agents[:,7])")
        sa_ids = np.unique(build[:, 3])
        pos = np.clip(np.searchsorted(sa_ids, agents[:, 22]), 0, len(sa_ids) - 1)
        sa = np.where(sa_ids[pos] == agents[:, 22], pos, -1).astype(np.int32)
        bld_sa = np.searchsorted(sa_ids, build[:, 3]).astype(np.int32)

        vis = np.asarray(bld_visits_by_agents, dtype=np.float64)
        V = vis.shape[1]
        routine = np.full((N, V), -1, dtype=np.int32)
        m = ~np.isnan(vis) & (vis != 0)
        routine[m] = _index_of(bld_ids[bo], bo, vis[m], "routine building (bld_visits_by_agents)")

        status = np.round(agents[:, 13] * 2).astype(np.uint8)
        if not np.isin(status, (ST_S, ST_I, ST_QH, ST_QI, ST_D, ST_H, ST_R, ST_X)).all():
            raise ValueError("unknown status code in agents[:,13]")
        src_id = agents[:, 21]
        src = np.full(N, -1, dtype=np.int32)
        has = src_id != 0
        if has.any():
            src[has] = _index_of(agent_ids[ao], ao, src_id[has], "infector id (agents[:,21])")

        import scipy.sparse as sp
        ip = interaction_prob if sp.issparse(interaction_prob) else sp.csr_matrix(np.asarray(interaction_prob))
        ip = sp.csr_matrix(ip)
        ip.sort_indices()
        if ip.shape != (N, N):
            raise ValueError("interaction_prob must be [N,N]")

        hh_always, trig_ptr, trig_hh = quirk_sets(agents, hh_ids, hh.astype(np.int32), diagnosis)
        return Population(
            N=N, H=len(hh_ids), B=B, S=len(sa_ids), V=V,
            status=status, t_inf=agents[:, 16].astype(np.int32), t_q=agents[:, 19].astype(np.int32),
            t_adm=agents[:, 24].astype(np.int32), src=src,
            hh=hh.astype(np.int32), home=home, sa=sa,
            risk=agents[:, 14].copy(), p_adm=agents[:, 23].copy(), p_death=agents[:, 26].copy(),
            routine=routine,
            ip_rowptr=ip.indptr.astype(np.int64), ip_col=ip.indices.astype(np.int32),
            ip_val=np.ascontiguousarray(ip.data, dtype=np.float64),
            open=(build[:, 10] != 0).astype(np.uint8),
            hh_always=hh_always, trig_ptr=trig_ptr, trig_hh=trig_hh,
            agent_ids=agent_ids, hh_ids=hh_ids, bld_ids=bld_ids, sa_ids=sa_ids,
            bld_lu=build[:, 1].copy(), bld_usg=build[:, 9].copy(), bld_sa=bld_sa)

    def write_back(self, agents, day):
        """Put the state columns back into a reference-layout agents matrix (contagious3.py column map).
        Columns 17/20/25 (day counters) are rebuilt the way the reference leaves them only for agents whose
        status refreshes them (contagious3.py:192-194); they are derived quantities on the device."""
        agents[:, 13] = self.status / 2.0
        agents[:, 16] = self.t_inf
        agents[:, 19] = self.t_q
        agents[:, 24] = self.t_adm
        agents[:, 21] = np.where(self.src >= 0, self.agent_ids[np.maximum(self.src, 0)], 0.0)
        return agents


def quirk_sets(agents, hh_ids, hh, diagnosis=7):
    """Reproduce `np.isin(agents[:, 1], new_diagnosed_agents)` (contagious3.py:267): the household id of a
    status-2 agent is tested against EVERY value in the newly diagnosed agents' full rows, not just their
    household column.  Returns
      hh_always[H]  households whose id equals a value present in every newly-diagnosed row (0 from the unused
                    zero columns / admission columns, 4 = the status code, `diagnosis` = column 17), and
      (trig_ptr, trig_hh)  CSR agent -> households whose id equals one of that agent's STATIC column values.
    Household ids that could collide with a DYNAMIC column value (a day number or an infector's agent id)
    are rejected: the SoA state does not carry enough to reproduce those collisions."""
    N = agents.shape[0]
    H = len(hh_ids)
    static_cols = [c for c in range(agents.shape[1]) if c not in DYNAMIC_COLS and c != 1]
    always_vals = np.array([0.0, 4.0, float(diagnosis)])
    hh_always = np.isin(hh_ids, always_vals).astype(np.uint8)
    small = (hh_ids == np.round(hh_ids)) & (hh_ids >= 1) & (hh_ids < 100000) & (hh_always == 0)
    if small.any():
        raise ValueError("household ids in [1,100000) can collide with day counters under the reference's "
                         "whole-row isin (contagious3.py:267); renumber households >= 100000")
    if np.isin(hh_ids, agents[:, 0]).any():
        raise ValueError("a household id equals an agent id: collides with agents[:,21] under contagious3.py:267")
    vals = agents[:, static_cols]
    pos = np.clip(np.searchsorted(hh_ids, vals), 0, H - 1)
    match = (hh_ids[pos] == vals) & (hh_always[pos] == 0) & (pos != hh[:, None])
    rows, cols = np.nonzero(match)
    pairs = np.unique(np.stack([rows, pos[rows, cols]], axis=1), axis=0) if len(rows) else np.zeros((0, 2), np.int64)
    trig_ptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(trig_ptr, pairs[:, 0] + 1, 1)
    trig_ptr = np.cumsum(trig_ptr)
    return hh_always, trig_ptr, pairs[:, 1].astype(np.int32)
