"""Routine generation (contagious3.py:139-170) on the device.

The reference builds every agent's static visit list from the dense all-pairs shortest-path matrix `dists` of the
communities.py graph: before each anchor (agents[:,12], agents[:,18], home) up to k flexible stops, each taken on a coin
flip and drawn uniformly from (buildings within bld_dist TO the anchor  ∩  buildings within bld_dist FROM the current
position) ∪ (buildings within a_dist of the agent).  `neighbourhoods()` turns what that code reads from `dists` into three
families of sorted index lists -- all the routine kernel needs, and all a sparse network builder has to provide for cities
where the dense matrix is out of reach; `generate()` runs the kernel through the C ABI (dys_generate_routines).
Draws are keyed (Philox; agent row and coin index), so the result does not depend on the order agents are processed in."""
import ctypes as C

import numpy as np

from . import engine


def _csr(mask):
    """boolean [rows, cols] -> (ptr int64[rows + 1], idx int32) with ascending column indices per row"""
    rows, cols = np.nonzero(mask)
    ptr = np.zeros(mask.shape[0] + 1, dtype=np.int64)
    np.add.at(ptr, rows + 1, 1)
    return np.cumsum(ptr), cols.astype(np.int32)


def neighbourhoods(nodes, dists, agents, build, a_dist, bld_dist):
    """What contagious3.py:139-170 reads from `dists`, as sorted lists over the buildings IN ASCENDING ID ORDER (the order
    np.intersect1d / np.union1d put them in).  Returns dict(bld_ids (sorted), home[N], anchors[N,3] (-1 = NaN anchor),
    out / inn / near = (ptr, idx))."""
    nodes = np.asarray(nodes)
    is_bld = nodes < 4000000                                   # the reference's own test for "is a building", :148,:152
    bn = np.nonzero(is_bld)[0]
    bn = bn[np.argsort(nodes[bn], kind="stable")]
    bld_ids = nodes[bn].astype(np.float64)
    pos = {float(v): i for i, v in enumerate(nodes)}
    an = np.array([pos[float(a)] for a in agents[:, 0]])
    Dbb = dists[np.ix_(bn, bn)]
    out = _csr(Dbb < bld_dist)                                 # within bld_dist FROM row building (c_nodes, :152)
    inn = _csr(Dbb.T < bld_dist)                               # within bld_dist TO row building (i_nodes, :148)
    in_build = np.isin(bld_ids, build[:, 0])
    near = _csr((dists[np.ix_(an, bn)] < a_dist) & in_build[None, :])      # a_nodes, :144
    home, anchors = _anchors(bld_ids, agents)
    return dict(bld_ids=bld_ids, home=home, anchors=anchors, out=out, inn=inn, near=near)


def _anchors(bld_ids, agents):
    rank = {float(v): i for i, v in enumerate(bld_ids)}
    reg = agents[:, [12, 18, 7]]                                # agents_reg[:, 2:], :131
    anchors = np.array([[rank[float(x)] if not np.isnan(x) else -1 for x in row] for row in reg], dtype=np.int32)
    home = np.array([rank[float(x)] for x in agents[:, 7]], dtype=np.int32)
    return home, anchors


def neighbourhoods_from_graph(nodes, indptr, indices, weights, agents, build, a_dist, bld_dist, device=0, rows_per_call=4096):
    """The same lists as neighbourhoods(), cut on the device straight from the graph (network.create_graph) by BOUNDED
    searches -- no (N + B)^2 distance matrix: `out` = buildings nearer than bld_dist from each building
    (dys_shortest_paths_csr with max_dist = bld_dist, listing building nodes only), `inn` = its transpose (the distances TO a
    building are the same numbers read column-wise), `near` = buildings of `build` nearer than a_dist from each agent."""
    import scipy.sparse as sp
    from . import network
    nodes = np.asarray(nodes, dtype=np.float64)
    if np.any(np.diff(nodes) <= 0):
        raise ValueError("nodes must be in ascending id order (network.create_graph returns them so)")
    is_bld = nodes < 4000000                                   # the reference's own test for "is a building", :148,:152
    bn = np.nonzero(is_bld)[0].astype(np.int32)
    bld_ids = nodes[bn]
    rank = np.full(len(nodes), -1, dtype=np.int32)
    rank[bn] = np.arange(len(bn), dtype=np.int32)

    def lists(sources, bound, keep):
        ptrs, cols = [np.zeros(1, dtype=np.int64)], []
        for k0 in range(0, len(sources), rows_per_call):
            src = sources[k0:k0 + rows_per_call]
            ptr, col, _ = network.shortest_paths_csr(indptr, indices, weights, src, capacity=len(src) * int(keep.sum()), skip_self=True,
                                                     device=device, max_dist=bound, col_keep=keep)
            ptrs.append(ptrs[-1][-1] + ptr[1:])
            cols.append(rank[col])
        return np.concatenate(ptrs), (np.concatenate(cols) if cols else np.zeros(0, dtype=np.int32))
    out = lists(bn, bld_dist, is_bld)                                                           # c_nodes, :152
    m = sp.csr_matrix((np.ones(len(out[1]), dtype=np.int8), out[1], out[0]), shape=(len(bn), len(bn))).T.tocsr()
    m.sort_indices()
    inn = (m.indptr.astype(np.int64), m.indices.astype(np.int32))                                # i_nodes, :148
    an = np.searchsorted(nodes, agents[:, 0]).astype(np.int32)
    near = lists(an, a_dist, is_bld & np.isin(nodes, build[:, 0]))                              # a_nodes, :144
    home, anchors = _anchors(bld_ids, agents)
    return dict(bld_ids=bld_ids, home=home, anchors=anchors, out=out, inn=inn, near=near)


def generate(nb, seed, V=10, k=2, device=0):
    """the routine kernel: int32 [N, V] building ranks (indices into nb['bld_ids']), -1 padded"""
    lib = engine.load_library()
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    home, anchors = c(nb["home"], np.int32), c(nb["anchors"], np.int32)
    ptrs = [c(nb[x][0], np.int64) for x in ("out", "inn", "near")]
    idxs = [c(nb[x][1], np.int32) for x in ("out", "inn", "near")]
    out = np.zeros((len(home), V), dtype=np.int32)
    rc = lib.dys_generate_routines(device, len(home), len(ptrs[0]) - 1, V, k, home.ctypes.data, anchors.ctypes.data,
                                   ptrs[0].ctypes.data, idxs[0].ctypes.data, ptrs[1].ctypes.data, idxs[1].ctypes.data,
                                   ptrs[2].ctypes.data, idxs[2].ctypes.data, C.c_uint64(seed), out.ctypes.data)
    if rc != 0:
        raise engine.DysError(rc, (lib.dys_last_error(None) or b"").decode())
    return out


def to_visits(routine, bld_ids):
    """bld_visits_by_agents as the reference leaves it (:164-170): building ids, NaN padded"""
    return np.where(routine >= 0, np.asarray(bld_ids)[np.maximum(routine, 0)], np.nan)
