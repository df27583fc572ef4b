"""Synthetic populations of the BASELINE.json sizes (1 M / 10 M / 100 M agents), generated community by community.

Mirrors the measured marginals of the shipped city: household sizes with mean ~2.6 (max 10);
age groups 38.3 / 53.3 / 8.4 % mapped to integer ages like create_random_data.py:32-34; risk / admission / mortality
~ N(mu, sigma) clipped at 0 by age band (parameters.py:8-23); 5 % as many buildings as agents, 70.6 % residential;
~1 200 agents per statistical area; routines of 2..10 slots built like contagious3.py:139-170 (home, then for each
anchor up to k=2 flexible stops taken with probability 1/2, then the anchor; the last anchor is home again);
20 seeds per 22 493 agents (create_random_data.py:109); social network = household cliques (interaction 1.0) plus
~14 similarity contacts per agent inside a community of 11 264 agents, a few to the next community, with
interaction_prob ~ U(0.95, 1) (communities.py:33 keeps links above 0.95).

Everything is a pure function of (n_agents, seed, community index), so a rank can generate exactly its own shard of a
100 M-agent population; agents are laid out by (community, statistical area, home building, household), which is also
the locality order the device kernels like.  numpy.random.Generator(PCG64([seed, community, purpose])).
"""
import numpy as np

from .population import Population, ST_I, ST_S

COMMUNITY = 11 * 1024            # agents per community (multiple of the kernels' 1024-agent granule)
HH_SIZE_P = np.array([0.27, 0.27, 0.16, 0.14, 0.09, 0.04, 0.015, 0.008, 0.004, 0.003])
AGE_GROUP_P = np.array([0.383, 0.533, 0.084])
# parameters.py:8-23 (lo, hi, mu, sigma) per age band
INFECTION = [(0, 18, 0.0742, 0.02), (18, 65, 0.0742 * 2, 0.04), (65, 85, 0.0742, 0.02), (85, 300, 0.0742 * 2, 0.04)]
ADMISSION = [(0, 30, 0.073 / 6, 0.02), (30, 50, 0.073 / 3, 0.02), (50, 60, 0.073 / 1.5, 0.02), (60, 70, 0.073, 0.02),
             (70, 80, 0.073 * 1.5, 0.02), (80, 300, 0.073 * 2.5, 0.02)]
MORTALITY = [(0, 40, 0.00002, 0), (40, 50, 0.002602, 0.001), (50, 60, 0.008822, 0.005), (60, 70, 0.026025, 0.015),
             (70, 80, 0.06402, 0.03), (80, 300, 0.174104, 0.1)]
V = 10
HH_ID0, AGENT_ID0, BLD_ID0 = 1.0e9, 5.0e9, 1.0


def _rng(seed, c, purpose):
    return np.random.Generator(np.random.PCG64([int(seed), int(c), int(purpose)]))


class Layout:
    """Global index layout: community sizes and the per-community household/building/area offsets."""

    def __init__(self, n_agents, seed):
        self.N, self.seed = int(n_agents), int(seed)
        self.C = max(1, -(-self.N // COMMUNITY))
        self.size = np.full(self.C, COMMUNITY, dtype=np.int64)
        self.size[-1] = self.N - COMMUNITY * (self.C - 1)
        self.a0 = np.concatenate([[0], np.cumsum(self.size)])
        self.nb = np.maximum(8, np.round(self.size * 0.05).astype(np.int64))        # buildings per community
        self.nres = np.maximum(2, np.round(self.nb * 0.706).astype(np.int64))       # residential ones come first
        self.ns = np.maximum(1, np.round(self.size / 1200.0).astype(np.int64))      # statistical areas
        self.b0 = np.concatenate([[0], np.cumsum(self.nb)])
        self.s0 = np.concatenate([[0], np.cumsum(self.ns)])
        self.hh_sizes = [self._hh_sizes(c) for c in range(self.C)]
        self.nh = np.array([len(x) for x in self.hh_sizes], dtype=np.int64)
        self.h0 = np.concatenate([[0], np.cumsum(self.nh)])
        self.B, self.S, self.H = int(self.b0[-1]), int(self.s0[-1]), int(self.h0[-1])

    def _hh_sizes(self, c):
        n = int(self.size[c])
        r = _rng(self.seed, c, 0)
        sizes = r.choice(np.arange(1, 11), size=n // 2 + 16, p=HH_SIZE_P / HH_SIZE_P.sum())
        cs = np.cumsum(sizes)
        k = int(np.searchsorted(cs, n))
        sizes = sizes[:k + 1].copy()
        sizes[k] -= cs[k] - n
        return sizes[sizes > 0]


def _banded_normal(r, age, table):
    out = np.zeros(len(age))
    for lo, hi, mu, sd in table:
        m = (age >= lo) & (age < hi)
        out[m] = r.normal(mu, sd, int(m.sum())) if sd > 0 else mu
    return np.maximum(out, 0.0)


def _anchors(L, c):
    """Anchor buildings (GLOBAL indices, -1 = none) of community c's agents before commuting is applied: non-residential
    buildings with heavy-tailed popularity (the largest gets ~10 % of the community)."""
    n, nb, nres, b0 = int(L.size[c]), int(L.nb[c]), int(L.nres[c]), int(L.b0[c])
    r = _rng(L.seed, c, 4)
    nnr = nb - nres
    w = 1.0 / (np.arange(nnr) + 3.0)
    w /= w.sum()
    a1 = np.where(r.random(n) < 0.845, b0 + nres + r.choice(nnr, size=n, p=w), -1)
    a2 = np.where(r.random(n) < 0.5, b0 + nres + r.choice(nnr, size=n, p=w), -1)
    return a1, a2


def _community_agents(L, c, degree=14, cross=0.02):
    """Static per-agent columns and routines of community c (local arrays, global indices)."""
    n, nb, nres, ns = int(L.size[c]), int(L.nb[c]), int(L.nres[c]), int(L.ns[c])
    a0, b0, s0, h0 = int(L.a0[c]), int(L.b0[c]), int(L.s0[c]), int(L.h0[c])
    r = _rng(L.seed, c, 1)
    sizes = L.hh_sizes[c]
    hh_local = np.repeat(np.arange(len(sizes)), sizes)
    first = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    # households fill residential buildings in order -> agents sorted by home building, buildings by area
    home_local = np.minimum((first * nres) // n, nres - 1)[hh_local]
    bld_area = np.empty(nb, dtype=np.int64)
    bld_area[:nres] = (np.arange(nres) * ns) // nres
    bld_area[nres:] = r.integers(0, ns, nb - nres)
    grp = r.choice(3, size=n, p=AGE_GROUP_P)
    age = np.where(grp == 0, r.integers(1, 18, n), np.where(grp == 1, r.integers(19, 65, n), r.integers(66, 90, n)))
    risk = _banded_normal(r, age, INFECTION)
    p_adm = _banded_normal(r, age, ADMISSION)
    p_death = _banded_normal(r, age, MORTALITY)
    anchor1, anchor2 = _anchors(L, c)
    if c + 1 < L.C and cross > 0:
        # commuters: an agent with a contact in the next community works where that contact works, so cross-community
        # (and, when sharded, cross-GPU) contacts really meet in a building
        u, v, _ = _community_contacts(L, c, degree, cross)
        far = v >= a0 + n
        nxt_a1, _ = _anchors(L, c + 1)
        uu, vv = u[far] - a0, v[far] - int(L.a0[c + 1])
        ok = nxt_a1[vv] >= 0
        anchor1 = anchor1.copy()
        anchor1[uu[ok]] = nxt_a1[vv[ok]]
    anchors = np.stack([anchor1, anchor2, home_local + b0], axis=1)              # agents_reg[:, 2:] of contagious3.py:131
    # flexible stops: k = 2 per anchor, each taken with probability 1/2, uniform over nearby buildings; a few
    # stops land in the next community so shards are genuinely coupled
    routine = np.full((n, V), -1, dtype=np.int64)
    routine[:, 0] = home_local + b0
    fill = np.ones(n, dtype=np.int64)
    nxt_b0, nxt_nb = (int(L.b0[c + 1]), int(L.nb[c + 1])) if c + 1 < L.C else (b0, nb)
    for j in range(3):
        has = anchors[:, j] >= 0
        for _ in range(2):
            take = has & (r.random(n) < 0.5)
            local = r.integers(0, nb, n) + b0
            far = r.integers(0, nxt_nb, n) + nxt_b0
            stop = np.where(r.random(n) < 0.03, far, local)
            idx = np.nonzero(take)[0]
            routine[idx, fill[idx]] = stop[idx]
            fill[idx] += 1
        idx = np.nonzero(has)[0]
        routine[idx, fill[idx]] = anchors[idx, j]
        fill[idx] += 1
    return dict(n=n, hh=hh_local + h0, home=home_local + b0, sa=bld_area[home_local] + s0, risk=risk, p_adm=p_adm,
                p_death=p_death, routine=routine.astype(np.int32), bld_area=bld_area + s0, age=age)


def _community_contacts(L, c, degree, cross):
    """Undirected similarity contacts generated BY community c: (u, v, ip) with u in c and v in c or c+1 (global)."""
    n, a0 = int(L.size[c]), int(L.a0[c])
    r = _rng(L.seed, c, 2)
    k = max(1, degree // 2)
    u = np.repeat(np.arange(n), k)
    # similarity is local: most partners live nearby (same area), some anywhere in the community
    near = np.round(r.normal(0.0, 300.0, u.size)).astype(np.int64)
    v = np.where(r.random(u.size) < 0.7, u + near, r.integers(0, n, u.size))
    v = np.mod(v, n)
    if c + 1 < L.C and cross > 0:
        far = r.random(u.size) < cross
        v = np.where(far, n + r.integers(0, int(L.size[c + 1]), u.size), v)
    keep = u != v
    u, v = u[keep], v[keep]
    ip = 0.95 + 0.05 * r.random(u.size)
    return u + a0, v + a0, ip


def synth_population(n_agents, seed=0, degree=14, cross=0.02, communities=None, seeds_per_22493=20):
    """Population for agents of communities [c_lo, c_hi) (default: all).  Arrays are local to that range except
    `routine`, which covers the halo [halo_lo, halo_hi) of agents any local row can reference (one community on
    each side), see Population.halo."""
    L = Layout(n_agents, seed)
    c_lo, c_hi = (0, L.C) if communities is None else communities
    h_lo, h_hi = max(0, c_lo - 1), min(L.C, c_hi + 1)
    parts = {c: _community_agents(L, c, degree, cross) for c in range(h_lo, h_hi)}
    own = [parts[c] for c in range(c_lo, c_hi)]
    cat = lambda k, dt: np.concatenate([p[k] for p in own]).astype(dt)
    lo, hi = int(L.a0[c_lo]), int(L.a0[c_hi])
    n = hi - lo
    routine = np.concatenate([parts[c]["routine"] for c in range(h_lo, h_hi)])
    halo = (int(L.a0[h_lo]), int(L.a0[h_hi]))
    hh = cat("hh", np.int32)

    # contacts: everything generated by communities c_lo-1 .. c_hi-1 that touches an owned row, both directions
    us, vs, ws = [], [], []
    for c in range(max(0, c_lo - 1), c_hi):
        u, v, w = _community_contacts(L, c, degree, cross)
        us += [u, v]; vs += [v, u]; ws += [w, w]
    # household cliques (same household -> probability 1, communities.py:39)
    sizes = np.concatenate([L.hh_sizes[c] for c in range(c_lo, c_hi)])
    first = lo + np.concatenate([[0], np.cumsum(sizes)[:-1]])
    for s in range(2, 11):
        f = first[sizes == s]
        if len(f) == 0:
            continue
        ii, jj = np.nonzero(~np.eye(s, dtype=bool))
        us.append((f[:, None] + ii[None, :]).ravel()); vs.append((f[:, None] + jj[None, :]).ravel())
        ws.append(np.ones(len(f) * len(ii)))
    u = np.concatenate(us); v = np.concatenate(vs); w = np.concatenate(ws)
    m = (u >= lo) & (u < hi)
    u, v, w = u[m], v[m], w[m]
    # sort by (row, col); on duplicates keep the largest probability (clique entries are 1.0)
    o = np.lexsort((-w, v, u))
    u, v, w = u[o], v[o], w[o]
    first_of = np.ones(len(u), dtype=bool)
    first_of[1:] = (u[1:] != u[:-1]) | (v[1:] != v[:-1])
    u, v, w = u[first_of], v[first_of], w[first_of]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, u - lo + 1, 1)
    rowptr = np.cumsum(rowptr)

    # seeds: status 2 for a fixed fraction, chosen per community
    status = np.full(n, ST_S, dtype=np.uint8)
    for c in range(c_lo, c_hi):
        r = _rng(L.seed, c, 3)
        nc = int(L.size[c])
        k = max(1, int(round(nc * seeds_per_22493 / 22493.0)))
        status[int(L.a0[c]) - lo + r.choice(nc, size=k, replace=False)] = ST_I

    b_lo, b_hi = int(L.b0[c_lo]), int(L.b0[c_hi])
    bld_sa = np.full(L.B, -1, dtype=np.int32)
    for c in range(h_lo, h_hi):
        bld_sa[int(L.b0[c]):int(L.b0[c + 1])] = parts[c]["bld_area"]
    bld_lu = np.ones(L.B)
    for c in range(L.C):
        nres, nb, b0 = int(L.nres[c]), int(L.nb[c]), int(L.b0[c])
        ncom = int(round((nb - nres) * 0.4))
        bld_lu[b0 + nres:b0 + nres + ncom] = 3
        bld_lu[b0 + nres + ncom:b0 + nb] = 4
    # a tenth of the public buildings are schools / places of worship (lockdown classes EDU / REL)
    bld_usg = np.zeros(L.B)
    pub = np.nonzero(bld_lu == 4)[0]
    bld_usg[pub[::10]] = 5310
    bld_usg[pub[5::10]] = 5501
    pop = Population(
        N=n, H=L.H, B=L.B, S=L.S, V=V,
        status=status, t_inf=np.zeros(n, np.int32), t_q=np.zeros(n, np.int32), t_adm=np.zeros(n, np.int32),
        src=np.full(n, -1, np.int32),
        hh=hh, home=cat("home", np.int32), sa=cat("sa", np.int32),
        risk=cat("risk", np.float64), p_adm=cat("p_adm", np.float64), p_death=cat("p_death", np.float64),
        routine=routine,
        ip_rowptr=rowptr, ip_col=v.astype(np.int32), ip_val=np.ascontiguousarray(w, dtype=np.float64),
        open=np.ones(L.B, np.uint8),
        hh_always=np.zeros(L.H, np.uint8), trig_ptr=np.zeros(n + 1, np.int64), trig_hh=np.zeros(0, np.int32),
        agent_ids=AGENT_ID0 + np.arange(lo, hi, dtype=np.float64), hh_ids=HH_ID0 + np.arange(L.H, dtype=np.float64),
        bld_ids=BLD_ID0 + np.arange(L.B, dtype=np.float64), sa_ids=np.arange(L.S, dtype=np.float64),
        bld_lu=bld_lu, bld_usg=bld_usg, bld_sa=bld_sa,
        N_glob=L.N, lo=lo, halo=halo, age=cat("age", np.int32),
        hh_range=(int(L.h0[c_lo]), int(L.h0[c_hi])), bld_range=(int(L.b0[c_lo]), int(L.b0[c_hi])),
        sa_range=(int(L.s0[c_lo]), int(L.s0[c_hi])))
    return pop


def to_reference(pop):
    """Reference-format matrices of a SMALL, unsharded synthetic population: agents[N,29], build[B,15],
    bld_visits_by_agents[N,V] (NaN padded), dense interaction_prob[N,N] -- what contagious3.py:176 consumes."""
    N, B = pop.N, pop.B
    assert getattr(pop, "lo", 0) == 0 and N <= 20000
    agents = np.zeros((N, 29))
    agents[:, 0] = pop.agent_ids
    agents[:, 1] = pop.hh_ids[pop.hh]
    agents[:, 4] = getattr(pop, "age", np.zeros(N))
    agents[:, 7] = pop.bld_ids[pop.home]
    agents[:, 13] = pop.status / 2.0
    agents[:, 14] = pop.risk
    agents[:, 16] = pop.t_inf
    agents[:, 19] = pop.t_q
    agents[:, 21] = np.where(pop.src >= 0, pop.agent_ids[np.maximum(pop.src, 0)], 0.0)
    agents[:, 22] = pop.sa_ids[pop.sa]
    agents[:, 23] = pop.p_adm
    agents[:, 24] = pop.t_adm
    agents[:, 26] = pop.p_death
    build = np.zeros((B, 15))
    build[:, 0] = pop.bld_ids
    build[:, 1] = pop.bld_lu
    build[:, 3] = pop.sa_ids[np.maximum(pop.bld_sa, 0)]
    build[:, 9] = pop.bld_usg
    build[:, 10] = pop.open
    visits = np.where(pop.routine >= 0, pop.bld_ids[np.maximum(pop.routine, 0)], np.nan)
    ip = np.zeros((N, N))
    rows = np.repeat(np.arange(N), np.diff(pop.ip_rowptr))
    ip[rows, pop.ip_col] = pop.ip_val
    return agents, build, visits, ip
