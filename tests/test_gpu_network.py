"""SURVEY.md 8f-2 on the GPU: dys_similarity_edges and dys_shortest_paths through the C ABI against the reference's own graph
(golden fixtures), the oracle and scipy's shortest_path (the call the reference makes)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.csgraph import shortest_path

from helpers import GOLDEN, REF_DIR
from test_network import _csr_from_edges, _fixture, _ulp_close, check_edges_against_reference, check_graph_against_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["city300", "full_city"])
def test_similarity_edges_match_the_reference_graph(which, oracle_lib):
    """every edge of the reference's G between agents, in np.where order, with agents_prob bit for bit (the oracle's value)"""
    from dysturbd_b200 import network
    z = _fixture() if which == "city300" else np.load(os.path.join(REF_DIR, "config1_inputs.npz"))
    cols = network.agent_columns(z["agents"], z["households"], z["build"])
    indptr, col, prob = network.similarity_edges(cols)
    check_edges_against_reference(indptr, col, prob, z)
    o_ptr, o_col, o_prob, _ = oracle_lib.similarity_edges(cols["age"], cols["income"], cols["x"], cols["y"], cols["household"], 0.95)
    assert np.array_equal(indptr, o_ptr) and np.array_equal(col, o_col) and np.array_equal(prob, o_prob)


@pytest.mark.parametrize("kw", [dict(min_prob=0.9, w_a=0.97, w_i=1.0, w_d=0.99),      # weights below 1: the pre-filter stays on
                                dict(min_prob=0.8, w_a=1.25, w_i=1.0, w_d=1.0),        # a weight above 1: every pair is evaluated
                                dict(min_prob=0.999)])
def test_similarity_edges_other_parameters(kw, oracle_lib):
    from dysturbd_b200 import network
    z = _fixture()
    cols = network.agent_columns(z["agents"], z["households"], z["build"])
    indptr, col, prob = network.similarity_edges(cols, **kw)
    o_ptr, o_col, o_prob, _ = oracle_lib.similarity_edges(cols["age"], cols["income"], cols["x"], cols["y"], cols["household"],
                                                           kw["min_prob"], kw.get("w_a", 1.0), kw.get("w_i", 1.0), kw.get("w_d", 1.0))
    assert indptr[-1] > 0
    assert np.array_equal(indptr, o_ptr) and np.array_equal(col, o_col) and np.array_equal(prob, o_prob)


def test_shortest_paths_equal_scipy_city300():
    from dysturbd_b200 import network
    z = _fixture()
    n = len(z["agents"])
    indptr, idx, w = _csr_from_edges(n, z["edge_row"], z["edge_col"], z["edge_w"])
    ref = shortest_path(sp.csr_matrix((w, idx, indptr), shape=(n, n)), directed=True, return_predecessors=False)
    got, sweeps = network.shortest_paths(indptr, idx, w, np.arange(n), return_sweeps=True)
    assert np.array_equal(got, ref) and sweeps > 3
    some = np.array([5, 5, 700, 0, 123], dtype=np.int32)                           # a short, unsorted list with a repeat
    assert np.array_equal(network.shortest_paths(indptr, idx, w, some), ref[some])


def test_shortest_paths_equal_scipy_full_city():
    """the shipped city's own graph (22 493 vertices, 311 512 edges, 236 components): 1 536 sources in two batches, grouped by
    component as the adapter sends them, and a strided sample in agent order"""
    from dysturbd_b200 import network
    z = np.load(os.path.join(REF_DIR, "config1_inputs.npz"))
    n = len(z["agents"])
    indptr, idx, w = _csr_from_edges(n, z["edge_row"], z["edge_col"], z["edge_w"])
    g = sp.csr_matrix((w, idx, indptr), shape=(n, n))
    order = np.argsort(network.components(indptr, idx), kind="stable").astype(np.int32)
    for src in (order[4000:5536], np.arange(0, n, n // 64)[:64].astype(np.int32)):
        assert np.array_equal(network.shortest_paths(indptr, idx, w, src), shortest_path(g, directed=True, indices=src))


def test_shortest_paths_csr_is_the_compacted_dense_result():
    from dysturbd_b200 import engine, network
    z = _fixture()
    n = len(z["agents"])
    indptr, idx, w = _csr_from_edges(n, z["edge_row"], z["edge_col"], z["edge_w"])
    src = np.random.RandomState(1).permutation(n)[:300].astype(np.int32)
    dense = network.shortest_paths(indptr, idx, w, src)
    for skip in (True, False):
        d = dense.copy()
        if skip:
            d[np.arange(len(src)), src] = np.inf
        r, c = np.nonzero(np.isfinite(d))
        ptr, col, dist = network.shortest_paths_csr(indptr, idx, w, src, capacity=len(r) + 7, skip_self=skip)
        assert np.array_equal(np.diff(ptr), np.bincount(r, minlength=len(src)))
        assert np.array_equal(col, c) and np.array_equal(dist, d[r, c])
    with pytest.raises(engine.DysError):
        network.shortest_paths_csr(indptr, idx, w, src, capacity=10)


def test_shortest_paths_directed_random_graph(oracle_lib):
    """one-way edges, zero weights, vertices without edges"""
    from dysturbd_b200 import network
    rs = np.random.RandomState(3)
    n, E = 3000, 9000
    row, col = rs.randint(0, n - 50, E), rs.randint(0, n - 50, E)            # the last 50 vertices are isolated
    keep = row != col
    key = np.unique(row[keep].astype(np.int64) * n + col[keep])              # no parallel edges (scipy would add them up)
    row, col = (key // n).astype(np.int32), (key % n).astype(np.int32)
    w = np.where(rs.rand(len(key)) < 0.2, 0.0, rs.rand(len(key)))
    indptr, idx, ww = _csr_from_edges(n, row, col, w)
    src = rs.permutation(n)[:700].astype(np.int32)
    got = network.shortest_paths(indptr, idx, ww, src)
    assert np.array_equal(got, oracle_lib.shortest_paths(indptr, idx, ww, src))
    g = sp.csr_matrix((ww, idx, indptr), shape=(n, n))
    assert np.array_equal(got, shortest_path(g, directed=True, indices=src))


def test_create_interaction_reproduces_interaction_prob():
    """agents, households, build -> interaction_prob, everything between create_data and the day loop: the reference's matrix"""
    from dysturbd_b200 import network
    z = _fixture()
    ip, (indptr, col, w) = network.create_interaction(z["agents"], z["households"], z["build"])
    assert np.array_equal(ip.indptr, z["ip_indptr"]) and np.array_equal(ip.indices, z["ip_indices"])
    assert _ulp_close(ip.data, z["ip_data"])
    ip2 = network.interaction_prob(indptr, col, w, rows_per_call=96)
    assert np.array_equal(ip2.indptr, ip.indptr) and np.array_equal(ip2.indices, ip.indices) and np.array_equal(ip2.data, ip.data)


def test_network_built_on_the_device_drives_the_day_loop():
    """closing the loop: agents, households, build -> interaction_prob on the GPU (no reference matrix involved) -> the day
    kernels -> the agent state and Stats of the literal reference run on every day of the golden fixture of the same city"""
    from dysturbd_b200 import engine as eng, network
    from dysturbd_b200.population import Population
    from helpers import Fixture, run_against_fixture
    z = _fixture()
    fx = Fixture("city300_philox_gradual_all")
    assert np.array_equal(fx.build, z["build"]) and np.array_equal(fx.agents[:, 0], z["agents"][:, 0])      # the same city
    ip, _ = network.create_interaction(z["agents"], z["households"], z["build"])
    pop = Population.from_reference(fx.agents, fx.build, fx.visits, ip)
    sim = eng.Engine(pop, fx.params)
    assert run_against_fixture(sim, fx) == fx.days
    sim.close()


def test_building_edges_match_the_oracle_and_the_reference_graph(oracle_lib):
    """the gravity model between buildings with numpy's pairwise row sums reproduced on the device: bld_prob bit for bit
    (shipped city's 1 015 buildings; random layouts on both sides of the pairwise block and split sizes)"""
    from dysturbd_b200 import network
    build = _fixture()["build"]
    indptr, col, prob = network.building_edges(build)
    o_ptr, o_col, o_prob, _ = oracle_lib.building_edges(build[:, 4], build[:, 6], build[:, 7], 0.00161)
    assert indptr[-1] == 83728
    assert np.array_equal(indptr, o_ptr) and np.array_equal(col, o_col) and np.array_equal(prob, o_prob)
    rs = np.random.RandomState(5)
    for B in (5, 8, 127, 128, 129, 300, 2500):
        b = np.zeros((B, 15))
        b[:, 4], b[:, 6], b[:, 7] = rs.rand(B) * 1000 + 1, rs.rand(B) * 5000, rs.rand(B) * 5000
        if B > 20:
            b[5, 6:8] = b[6, 6:8]
        got = network.building_edges(b, min_prob=0.5 / B)
        exp = oracle_lib.building_edges(b[:, 4], b[:, 6], b[:, 7], 0.5 / B)
        assert got[0][-1] > 0 and all(np.array_equal(g, e) for g, e in zip(got, exp[:3])), B


def test_whole_setup_on_the_device():
    """everything between create_data and the day loop: agents, households, build -> G (all four kinds of edges) -> dists
    -> neighbourhood lists -> routines, each stage against the reference's own (tests/golden: the whole of G for one city,
    the lists from the reference's dists and its routines under injected draws for another)"""
    from dysturbd_b200 import network, routines
    z = _fixture()
    check_graph_against_reference(*network.create_graph(z["agents"], z["households"], z["build"]), z)
    r = np.load(os.path.join(GOLDEN, "routines_city120.npz"))
    nodes, indptr, idx, w = network.create_graph(r["agents"], r["households"], r["build"])
    dists = network.all_pairs(indptr, idx, w)
    nb = routines.neighbourhoods(nodes, dists, r["agents"], r["build"], 1, 6)
    for name, key in (("out", "out"), ("inn", "in"), ("near", "near")):
        assert np.array_equal(nb[name][0], r[key + "_ptr"]) and np.array_equal(nb[name][1], r[key + "_idx"]), name
    visits = routines.to_visits(routines.generate(nb, int(r["seed"]), V=r["visits"].shape[1], k=int(r["k"])), nb["bld_ids"])
    assert np.array_equal(visits, r["visits"], equal_nan=True)
    # the same lists by bounded searches on the device (no dense distance matrix)
    nb2 = routines.neighbourhoods_from_graph(nodes, indptr, idx, w, r["agents"], r["build"], 1, 6, rows_per_call=300)
    for key in ("out", "inn", "near"):
        assert np.array_equal(nb2[key][0], nb[key][0]) and np.array_equal(nb2[key][1], nb[key][1]), key


def test_bounded_search_lists_exactly_the_entries_below_the_bound():
    from dysturbd_b200 import network
    z = _fixture()
    n = len(z["agents"])
    indptr, idx, w = _csr_from_edges(n, z["edge_row"], z["edge_col"], z["edge_w"])
    src = np.arange(0, n, 3, dtype=np.int32)
    dense = network.shortest_paths(indptr, idx, w, src)
    dense[np.arange(len(src)), src] = np.inf
    keep = (np.arange(n) % 5 != 0).astype(np.uint8)
    for bound in (0.02, 0.1, 0.5):
        m = (dense < bound) & keep.astype(bool)[None, :]
        r, c = np.nonzero(m)
        ptr, col, dist = network.shortest_paths_csr(indptr, idx, w, src, capacity=len(r) + 1, max_dist=bound, col_keep=keep)
        assert len(r) > 50 and np.array_equal(np.diff(ptr), np.bincount(r, minlength=len(src)))
        assert np.array_equal(col, c) and np.array_equal(dist, dense[r, c])


def test_network_calls_on_degenerate_inputs(oracle_lib):
    """no edges, one vertex, no sources, one agent, one household, one building, everybody on one spot"""
    from dysturbd_b200 import network
    empty = np.zeros(6, dtype=np.int64)
    d = network.shortest_paths(empty, np.zeros(0, dtype=np.int32), np.zeros(0), [3, 0])
    exp = np.full((2, 5), np.inf)
    exp[0, 3] = exp[1, 0] = 0.0
    assert np.array_equal(d, exp)
    assert network.shortest_paths(empty, np.zeros(0, dtype=np.int32), np.zeros(0), []).shape == (0, 5)
    assert np.array_equal(network.shortest_paths(np.zeros(2, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0), [0]), [[0.0]])
    ptr, col, dist = network.shortest_paths_csr(empty, np.zeros(0, dtype=np.int32), np.zeros(0), [1, 2], capacity=0)
    assert np.array_equal(ptr, [0, 0, 0]) and len(col) == 0
    one = dict(age=np.array([30.0]), income=np.array([1.0]), x=np.array([5.0]), y=np.array([6.0]), household=np.zeros(1, dtype=np.int64))
    indptr, col, prob = network.similarity_edges(one)
    assert np.array_equal(indptr, [0, 0]) and len(col) == 0
    fam = dict(age=np.array([30.0, 5.0, 61.0]), income=np.full(3, 7.0), x=np.full(3, 5.0), y=np.full(3, 6.0), household=np.zeros(3, dtype=np.int64))
    indptr, col, prob = network.similarity_edges(fam)                    # one household: all six ordered pairs, p = 1 (:39)
    assert np.array_equal(indptr, [0, 2, 4, 6]) and np.array_equal(col, [1, 2, 0, 2, 0, 1]) and np.all(prob == 1.0)
    two = dict(age=np.array([30.0, 30.0, 31.0]), income=np.full(3, 7.0), x=np.full(3, 5.0), y=np.full(3, 6.0), household=np.arange(3, dtype=np.int64))
    got = network.similarity_edges(two)                                  # zero income and distance spread: 0 / 0 = nan, no edge, as in numpy
    o = oracle_lib.similarity_edges(two["age"], two["income"], two["x"], two["y"], two["household"], 0.95)
    assert got[0][-1] == 0 and np.array_equal(got[0], o[0])
    b = np.zeros((1, 15))
    b[0, 4] = 100.0
    assert network.building_edges(b)[0][-1] == 0
    spot = np.zeros((4, 15))
    spot[:, 4] = [10.0, 20.0, 30.0, 40.0]                                # four buildings on one spot: every distance 0 -> 0.00001 (:14)
    got = network.building_edges(spot, min_prob=0.1)
    exp = oracle_lib.building_edges(spot[:, 4], spot[:, 6], spot[:, 7], 0.1)
    assert got[0][-1] > 0 and all(np.array_equal(g, e) for g, e in zip(got, exp[:3]))


def test_network_calls_reject_bad_input():
    from dysturbd_b200 import engine, network
    indptr, idx = np.array([0, 1, 2], dtype=np.int64), np.array([1, 0], dtype=np.int32)
    with pytest.raises(engine.DysError):
        network.shortest_paths(indptr, idx, np.array([0.5, -0.1]), [0])
    with pytest.raises(engine.DysError):
        network.shortest_paths(indptr, idx, np.array([0.5, 0.1]), [2])
    with pytest.raises(engine.DysError):
        network.shortest_paths(indptr, np.array([1, 7], dtype=np.int32), np.array([0.5, 0.1]), [0])
    cols = dict(age=np.array([1.0, np.nan]), income=np.zeros(2), x=np.zeros(2), y=np.zeros(2), household=np.zeros(2, dtype=np.int64))
    with pytest.raises(engine.DysError):
        network.similarity_edges(cols)
