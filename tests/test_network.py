"""SURVEY.md 8f-2, CPU side: the oracle restatement of the network construction (oracle/dys_oracle.c: orc_similarity_edges,
orc_shortest_paths) against the REFERENCE's own graph -- the agent-agent edges communities.create_network put into G and the
interaction_prob contagious3.py:124-130 derived from it (tests/golden/network_city300.npz, tests/golden/config1/) -- and
against scipy's shortest_path, the very call the reference makes.  The host adapter (dysturbd_b200/network.py) is driven
with the oracle standing in for the two device calls."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.csgraph import shortest_path

from helpers import GOLDEN, REF_DIR


def _ulp_close(a, b):
    return np.all(np.abs(a - b) <= np.spacing(np.maximum(np.abs(a), np.abs(b))))


def _fixture():
    return np.load(os.path.join(GOLDEN, "network_city300.npz"))


def _csr_from_edges(n, row, col, w):
    order = np.lexsort((col, row))
    indptr = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=n))]).astype(np.int64)
    return indptr, col[order].astype(np.int32), w[order]


def check_edges_against_reference(indptr, col, prob, z):
    """(indptr, col, prob) vs the agent-agent edges of the reference's G: same pairs in np.where order, weights -log(p)"""
    n = len(z["agents"])
    assert indptr[-1] == len(z["edge_row"])
    assert np.array_equal(np.repeat(np.arange(n), np.diff(indptr)), z["edge_row"])
    assert np.array_equal(col, z["edge_col"])
    w = -np.log(prob)
    assert _ulp_close(w, z["edge_w"])          # (equal bit for bit where numpy's log is the one that made the fixture)
    return w


def test_similarity_edges_oracle_matches_reference_graph(oracle_lib):
    from dysturbd_b200 import network
    z = _fixture()
    cols = network.agent_columns(z["agents"], z["households"], z["build"])
    indptr, col, prob, maxes = oracle_lib.similarity_edges(cols["age"], cols["income"], cols["x"], cols["y"], cols["household"], 0.95)
    check_edges_against_reference(indptr, col, prob, z)
    assert (prob == 1.0).sum() > 100 and (prob < 1.0).sum() > 100       # household edges and similarity edges both occur


def test_similarity_edges_oracle_full_city(oracle_lib):
    """the full shipped city: 22 493 agents, 311 512 edges of the reference's own graph"""
    from dysturbd_b200 import network
    z = np.load(os.path.join(REF_DIR, "config1_inputs.npz"))
    cols = network.agent_columns(z["agents"], z["households"], z["build"])
    indptr, col, prob, _ = oracle_lib.similarity_edges(cols["age"], cols["income"], cols["x"], cols["y"], cols["household"], 0.95)
    check_edges_against_reference(indptr, col, prob, z)


def test_agent_columns_follow_argmax(oracle_lib):
    """communities.py:23, 34 look ids up with (table[:, None] == keys).argmax(axis=0): first match, row 0 when there is none"""
    from dysturbd_b200 import network
    z = _fixture()
    agents, hhs, build = z["agents"].copy(), z["households"], z["build"]
    agents[5, 1] = -12345.0                                   # a household id that does not exist
    cols = network.agent_columns(agents, hhs, build)
    agent_hh = (hhs[:, 0][:, None] == agents[:, 1]).argmax(axis=0)
    homes = (build[:, 0][:, None] == hhs[agent_hh, 1]).argmax(axis=0)
    assert np.array_equal(cols["household"], agent_hh) and agent_hh[5] == 0
    assert np.array_equal(cols["x"], build[homes, 6]) and np.array_equal(cols["income"], hhs[agent_hh, 2])


def test_shortest_paths_oracle_equals_scipy(oracle_lib):
    z = _fixture()
    n = len(z["agents"])
    indptr, idx, w = _csr_from_edges(n, z["edge_row"], z["edge_col"], z["edge_w"])
    g = sp.csr_matrix((w, idx, indptr), shape=(n, n))
    ref = shortest_path(g, directed=True, return_predecessors=False)
    got = oracle_lib.shortest_paths(indptr, idx, w, np.arange(n))
    assert np.array_equal(got, ref)
    assert np.isfinite(ref).mean() < 0.5 and (ref == 0).sum() > n                 # several components; zero-weight household edges


def test_shortest_paths_oracle_full_city_sample(oracle_lib):
    z = np.load(os.path.join(REF_DIR, "config1_inputs.npz"))
    n = len(z["agents"])
    indptr, idx, w = _csr_from_edges(n, z["edge_row"], z["edge_col"], z["edge_w"])
    g = sp.csr_matrix((w, idx, indptr), shape=(n, n))
    src = np.arange(0, n, n // 96)[:96].astype(np.int32)
    assert np.array_equal(oracle_lib.shortest_paths(indptr, idx, w, src), shortest_path(g, directed=True, indices=src))


def test_interaction_prob_adapter_reproduces_the_reference(oracle_lib, monkeypatch):
    """dysturbd_b200/network.py end to end (agent columns -> edges -> -log p -> shortest paths by component batches ->
    exp(-d) -> CSR in agent order) with the oracle in place of the two device calls == the reference's interaction_prob"""
    from dysturbd_b200 import network
    z = _fixture()

    def edges(cols, min_prob=0.95, w_a=1.0, w_i=1.0, w_d=1.0, device=0):
        return oracle_lib.similarity_edges(cols["age"], cols["income"], cols["x"], cols["y"], cols["household"], min_prob, w_a, w_i, w_d)[:3]
    monkeypatch.setattr(network, "similarity_edges", edges)

    def paths_csr(indptr, idx, w, src, capacity=None, skip_self=True, device=0, out=None):
        d = oracle_lib.shortest_paths(indptr, idx, w, src)
        if skip_self:
            d[np.arange(len(src)), src] = np.inf
        r, c = np.nonzero(np.isfinite(d))
        col, dist, off = out
        assert off + len(r) <= len(col)
        col[off:off + len(r)], dist[off:off + len(r)] = c, d[r, c]
        return np.concatenate([[0], np.cumsum(np.bincount(r, minlength=len(src)))]), col[off:off + len(r)], dist[off:off + len(r)]
    monkeypatch.setattr(network, "shortest_paths_csr", paths_csr)
    ip, (indptr, col, w) = network.create_interaction(z["agents"], z["households"], z["build"])
    n = len(z["agents"])
    ip2 = network.interaction_prob(indptr, col, w, rows_per_call=100)              # several calls, components split across them
    for m in (ip, ip2):
        assert np.array_equal(m.indptr, z["ip_indptr"]) and np.array_equal(m.indices, z["ip_indices"])
        assert _ulp_close(m.data, z["ip_data"])
    assert ip.shape == (n, n)
    # rows are independent: two "ranks" take disjoint row ranges, stacking their pieces gives the whole matrix
    cut = n // 3
    parts = [network.interaction_prob(indptr, col, w, row_range=r) for r in ((0, cut), (cut, n))]
    whole = sp.vstack(parts).tocsr()
    assert np.array_equal(whole.indptr, ip.indptr) and np.array_equal(whole.indices, ip.indices) and np.array_equal(whole.data, ip.data)


def _oracle_standins(monkeypatch, oracle_lib):
    from dysturbd_b200 import network

    def edges(cols, min_prob=0.95, w_a=1.0, w_i=1.0, w_d=1.0, device=0):
        return oracle_lib.similarity_edges(cols["age"], cols["income"], cols["x"], cols["y"], cols["household"], min_prob, w_a, w_i, w_d)[:3]
    monkeypatch.setattr(network, "similarity_edges", edges)
    monkeypatch.setattr(network, "building_edges", lambda build, min_prob=0.00161, device=0:
                        oracle_lib.building_edges(build[:, 4], build[:, 6], build[:, 7], min_prob)[:3])
    monkeypatch.setattr(network, "shortest_paths", lambda indptr, idx, w, src, device=0: oracle_lib.shortest_paths(indptr, idx, w, src))

    def paths_csr(indptr, idx, w, src, capacity=None, skip_self=True, device=0, out=None, max_dist=np.inf, col_keep=None):
        d = oracle_lib.shortest_paths(indptr, idx, w, src)
        if skip_self:
            d[np.arange(len(src)), src] = np.inf
        keep = d < max_dist
        if col_keep is not None:
            keep &= np.asarray(col_keep, dtype=bool)[None, :]
        r, c = np.nonzero(keep)
        assert capacity is None or len(r) <= capacity
        return np.concatenate([[0], np.cumsum(np.bincount(r, minlength=len(src)))]).astype(np.int64), c.astype(np.int32), d[r, c]
    monkeypatch.setattr(network, "shortest_paths_csr", paths_csr)
    return network


def check_graph_against_reference(nodes, indptr, idx, w, z):
    """(nodes, CSR) vs the reference's DiGraph G: same nodes, same edges, same weights (zeros with their sign)"""
    assert np.array_equal(nodes, np.sort(z["g_nodes"]))
    assert indptr[-1] == len(z["g_u"])
    ref = {(a, b): c for a, b, c in zip(z["g_u"], z["g_v"], z["g_w"])}
    rows = np.repeat(np.arange(len(nodes)), np.diff(indptr))
    mine = {(nodes[r], nodes[c]): x for r, c, x in zip(rows, idx, w)}
    assert mine.keys() == ref.keys()
    a = np.array([mine[k] for k in ref])
    b = np.array([ref[k] for k in ref])
    assert _ulp_close(a, b) and np.array_equal(np.signbit(a), np.signbit(b))


def test_building_edges_oracle_matches_reference_graph_and_numpy(oracle_lib):
    """the gravity model between buildings: the 83 728 building -> building edges of the reference's G, and the row sums
    against numpy's own pairwise `sum(axis=1)` for sizes on both sides of its block and split thresholds"""
    z = _fixture()
    build = z["build"]
    indptr, col, prob, _ = oracle_lib.building_edges(build[:, 4], build[:, 6], build[:, 7], 0.00161)
    bb = (z["g_u"] < 4000000) & (z["g_v"] < 4000000)
    row_of = {b: i for i, b in enumerate(build[:, 0])}
    r = np.array([row_of[a] for a in z["g_u"][bb]])
    c = np.array([row_of[a] for a in z["g_v"][bb]])
    order = np.lexsort((c, r))
    assert np.array_equal(np.repeat(np.arange(len(build)), np.diff(indptr)), r[order]) and np.array_equal(col, c[order])
    assert _ulp_close(-np.log(prob), z["g_w"][bb][order])
    rs = np.random.RandomState(0)
    for B in (5, 8, 100, 128, 129, 300, 1000, 3000):
        fs, x, y = rs.rand(B) * 1000 + 1, rs.rand(B) * 5000, rs.rand(B) * 5000
        if B > 20:
            x[5], y[5] = x[6], y[6]                               # two buildings on one spot: the 0 -> 0.00001 rule (:14)
        d = (np.sqrt((x[None, :] - x[:, None]) ** 2 + (y[None, :] - y[:, None]) ** 2)) ** 2
        d[d == 0] = 0.00001
        sc = fs / d
        np.fill_diagonal(sc, 0)
        assert np.array_equal(oracle_lib.building_edges(fs, x, y, 0.5)[3], sc.sum(axis=1)), B


def test_create_graph_is_the_reference_graph(oracle_lib, monkeypatch):
    """network.create_graph == communities.create_network: every node and edge of G (building gravity edges, agent -> home /
    work / other, agent -> agent), and the distance matrix over it == scipy's on the reference's own G"""
    network = _oracle_standins(monkeypatch, oracle_lib)
    z = _fixture()
    nodes, indptr, idx, w = network.create_graph(z["agents"], z["households"], z["build"])
    check_graph_against_reference(nodes, indptr, idx, w, z)
    gn = z["g_nodes"]
    pos = {v: i for i, v in enumerate(gn)}
    G = sp.csr_matrix((z["g_w"], (np.array([pos[a] for a in z["g_u"]]), np.array([pos[a] for a in z["g_v"]]))), shape=(len(gn), len(gn)))
    ref = shortest_path(G, directed=True, return_predecessors=False)
    perm = np.array([pos[v] for v in nodes])
    assert np.array_equal(network.all_pairs(indptr, idx, w, rows_per_call=500, diagonal_inf=False), ref[np.ix_(perm, perm)])


def test_whole_setup_chain_reproduces_the_routine_inputs(oracle_lib, monkeypatch):
    """agents, households, build -> graph -> distances -> the neighbourhood lists of the routine kernel: exactly the lists
    derived from the reference's own `dists` (tests/golden/routines_city120.npz), and then the reference's routines"""
    from dysturbd_b200 import routines
    network = _oracle_standins(monkeypatch, oracle_lib)
    z = np.load(os.path.join(GOLDEN, "routines_city120.npz"))
    nodes, indptr, idx, w = network.create_graph(z["agents"], z["households"], z["build"])
    dists = network.all_pairs(indptr, idx, w)
    nb = routines.neighbourhoods(nodes, dists, z["agents"], z["build"], 1, 6)
    assert np.array_equal(nb["bld_ids"], z["bld_ids"]) and np.array_equal(nb["home"], z["home"]) and np.array_equal(nb["anchors"], z["anchors"])
    for name, key in (("out", "out"), ("inn", "in"), ("near", "near")):
        assert np.array_equal(nb[name][0], z[key + "_ptr"]) and np.array_equal(nb[name][1], z[key + "_idx"]), name
    r = oracle_lib.routines(nb["home"], nb["anchors"], nb["out"], nb["inn"], nb["near"], int(z["seed"]), V=z["visits"].shape[1], k=int(z["k"]))
    assert np.array_equal(routines.to_visits(r, nb["bld_ids"]), z["visits"], equal_nan=True)
    # the same lists without the dense matrix: bounded searches listing building nodes only
    nb2 = routines.neighbourhoods_from_graph(nodes, indptr, idx, w, z["agents"], z["build"], 1, 6, rows_per_call=300)
    for key in ("bld_ids", "home", "anchors"):
        assert np.array_equal(nb2[key], nb[key]), key
    for key in ("out", "inn", "near"):
        assert np.array_equal(nb2[key][0], nb[key][0]) and np.array_equal(nb2[key][1], nb[key][1]), key


def test_shortest_paths_oracle_property_random_graphs(oracle_lib):
    """size-independent check of the fixed-point argument: on random directed graphs -- zero weights, tiny and
    huge weights side by side, isolated vertices, self loops -- the restatement's left-to-right sums equal scipy's Dijkstra
    bit for bit, and every finite distance satisfies d(v) <= fl(d(u) + w(u, v)) for every edge (a fixed point)"""
    rs = np.random.RandomState(11)
    for trial in range(40):
        n = int(rs.randint(2, 120))
        E = int(rs.randint(0, 6 * n))
        row, col = rs.randint(0, n, E), rs.randint(0, n, E)
        key = np.unique(row.astype(np.int64) * n + col)
        row, col = (key // n).astype(np.int32), (key % n).astype(np.int32)
        scale = rs.choice([1e-300, 1e-9, 1.0, 1e9], len(key))
        w = np.where(rs.rand(len(key)) < 0.25, 0.0, rs.rand(len(key)) * scale)
        indptr, idx, ww = _csr_from_edges(n, row, col, w)
        src = np.arange(n, dtype=np.int32)
        got = oracle_lib.shortest_paths(indptr, idx, ww, src)
        g = sp.csr_matrix((ww, idx, indptr), shape=(n, n))
        assert np.array_equal(got, shortest_path(g, directed=True, return_predecessors=False)), trial
        rows = np.repeat(np.arange(n), np.diff(indptr))
        with np.errstate(invalid="ignore"):
            assert np.all(got[:, idx] <= got[:, rows] + ww[None, :]), trial
