"""Network construction on the GPU: the host-side mirror of what the reference does between
`create_data` and the day loop --

    G = create_network(agents, households, build)                 communities.py:7-77
    dists = shortest_path(nx.to_scipy_sparse_matrix(G), ...)      contagious3.py:124-126
    interaction_prob = np.exp(-dists[agent rows][:, agent cols])  contagious3.py:127-130

-- for the agent block of that graph.  The reference materialises five N x N float64 matrices for the edges and an
(N + B)^2 distance matrix; here the edges come out of `dys_similarity_edges` as CSR (no N x N array anywhere) and the
distances out of `dys_shortest_paths` in row batches that are turned into interaction_prob rows at once.  Agent -> agent
paths only run through agent nodes (agents have no incoming edge from a building), so the agent block is the whole story
for interaction_prob.

What stays on the host, on purpose: `-np.log(p)` over the E edges and `np.exp(-d)` over the reachable pairs.  Both are
the reference's own numpy calls; a device log / exp is a different libm and would not be bit-identical, while the edge
probabilities and the path sums are IEEE add / multiply / divide / sqrt only and ARE bit-identical on the device.
"""
import ctypes as C

import numpy as np

from . import engine


def agent_columns(agents, households, build):
    """the per-agent columns create_network derives (communities.py:23-35), with its `argmax` look-ups: an id that is not
    found resolves to row 0, as there"""
    def rows_of(keys, table):
        order = np.argsort(table, kind="stable")
        pos = np.searchsorted(table[order], keys)
        pos = np.minimum(pos, len(table) - 1)
        hit = table[order][pos] == keys
        return np.where(hit, order[pos], 0)           # (argmax of an all-False column is 0)
    agent_hh = rows_of(agents[:, 1], households[:, 0])                                         # :23
    hh_home = households[agent_hh, 1]                                                           # :33
    agent_homes = rows_of(hh_home, build[:, 0])                                                 # :34
    return dict(age=np.ascontiguousarray(agents[:, 4], dtype=np.float64),
                income=np.ascontiguousarray(households[agent_hh, 2], dtype=np.float64),
                x=np.ascontiguousarray(build[agent_homes, 6], dtype=np.float64),
                y=np.ascontiguousarray(build[agent_homes, 7], dtype=np.float64),
                household=np.ascontiguousarray(agent_hh, dtype=np.int64))


def _check(lib, rc):
    if rc != 0:
        raise engine.DysError(rc, (lib.dys_last_error(None) or b"").decode())


def similarity_edges(cols, min_prob=0.95, w_a=1.0, w_i=1.0, w_d=1.0, device=0):
    """communities.py:22-40, 65-70 -> (indptr[n+1], col[E], prob[E]) with prob == agents_prob[i, j] bit for bit"""
    lib = engine.load_library()
    n = len(cols["age"])
    lib.dys_similarity_edges.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 5 + [C.c_double] * 4 + [C.c_void_p, C.c_int64] + [C.c_void_p] * 3
    p = lambda k, dt: np.ascontiguousarray(cols[k], dtype=dt)
    arr = [p("age", np.float64), p("income", np.float64), p("x", np.float64), p("y", np.float64), p("household", np.int64)]
    head = [device, n] + [a.ctypes.data for a in arr] + [float(w_a), float(w_i), float(w_d), float(min_prob)]
    indptr = np.zeros(n + 1, dtype=np.int64)
    _check(lib, lib.dys_similarity_edges(*head, indptr.ctypes.data, 0, None, None, None))      # sizing pass
    E = int(indptr[-1])
    col, prob = np.empty(E, dtype=np.int32), np.empty(E, dtype=np.float64)
    _check(lib, lib.dys_similarity_edges(*head, indptr.ctypes.data, E, col.ctypes.data, prob.ctypes.data, None))
    return indptr, col, prob


def building_edges(build, min_prob=0.00161, device=0):
    """communities.py:12-21, 53-54 -> (indptr[B+1], col[E], prob[E]) over the rows of `build`, prob == bld_prob[i, j] bit for bit"""
    lib = engine.load_library()
    lib.dys_building_edges.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 3 + [C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    fs, x, y = (np.ascontiguousarray(build[:, k], dtype=np.float64) for k in (4, 6, 7))
    B = len(fs)
    indptr = np.zeros(B + 1, dtype=np.int64)
    head = [device, B, fs.ctypes.data, x.ctypes.data, y.ctypes.data, float(min_prob), indptr.ctypes.data]
    _check(lib, lib.dys_building_edges(*head, 0, None, None))                                   # sizing pass
    E = int(indptr[-1])
    col, prob = np.empty(E, dtype=np.int32), np.empty(E, dtype=np.float64)
    _check(lib, lib.dys_building_edges(*head, E, col.ctypes.data, prob.ctypes.data))
    return indptr, col, prob


def create_graph(agents, households, build, a_min_prob=0.95, b_min_prob=0.00161, w_a=1.0, w_i=1.0, w_d=1.0, device=0):
    """communities.create_network(agents, households, build) as arrays: (nodes, indptr, indices, weights) -- the node ids
    (buildings and agents, ascending) and the CSR adjacency of the reference's DiGraph G with its weights -log(p):
    building -> building (gravity model, :52-54), agent -> home / workplace / other regular building (weight log 1 = 0,
    :56-61; a building named twice by an agent is one edge, as in a DiGraph) and agent -> agent (:65-70)."""
    cols = agent_columns(agents, households, build)
    b_ptr, b_col, b_prob = building_edges(build, b_min_prob, device=device)
    a_ptr, a_col, a_prob = similarity_edges(cols, a_min_prob, w_a, w_i, w_d, device=device)
    N, B = len(agents), len(build)
    ids_b, ids_a = build[:, 0], agents[:, 0]
    hh_home = households[cols["household"], 1]
    u = [np.repeat(ids_b, np.diff(b_ptr)), ids_a]                                               # :52-56
    v = [ids_b[b_col], hh_home]
    w = [-np.log(b_prob), np.full(N, np.log(1))]
    for c in (12, 18):                                                                          # :58-61
        has = ~np.isnan(agents[:, c])
        u.append(ids_a[has]); v.append(agents[has, c]); w.append(np.full(int(has.sum()), np.log(1)))
    u.append(np.repeat(ids_a, np.diff(a_ptr))); v.append(ids_a[a_col]); w.append(edge_weights(a_prob))   # :65-70
    u, v, w = np.concatenate(u), np.concatenate(v), np.concatenate(w)
    nodes = np.unique(np.concatenate([u, v]))
    r, c = np.searchsorted(nodes, u), np.searchsorted(nodes, v)
    key = r.astype(np.int64) * len(nodes) + c
    _, last = np.unique(key[::-1], return_index=True)                                           # a repeated (u, v): the LAST weight stands
    keep = np.sort(len(key) - 1 - last)
    r, c, w = r[keep], c[keep], w[keep]
    order = np.lexsort((c, r))
    indptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=len(nodes)))]).astype(np.int64)
    return nodes, indptr, c[order].astype(np.int32), w[order]


def all_pairs(indptr, indices, weights, device=0, rows_per_call=4096, diagonal_inf=True):
    """contagious3.py:125-126: the dense `dists` of a graph (every node a source), float64 [n, n], +inf where there is no
    path and -- as the reference leaves it for everything that follows (:126 np.fill_diagonal) -- on the diagonal"""
    n = len(indptr) - 1
    out = np.empty((n, n), dtype=np.float64)
    for k0 in range(0, n, rows_per_call):
        src = np.arange(k0, min(n, k0 + rows_per_call), dtype=np.int32)
        out[k0:k0 + len(src)] = shortest_paths(indptr, indices, weights, src, device=device)
    if diagonal_inf:
        np.fill_diagonal(out, np.inf)
    return out


def edge_weights(prob):
    """the weight create_network gives an edge (communities.py:70): -log(p), with numpy's log as there"""
    return -np.log(prob)


def shortest_paths(indptr, indices, weights, sources, device=0, return_sweeps=False, return_info=False):
    """scipy.sparse.csgraph.shortest_path(g, directed=True, indices=sources) -> float64 [len(sources), n]
    return_info: also dict(sweeps, prep_s, relax_s, download_s) as the library measured them"""
    lib = engine.load_library()
    lib.dys_shortest_paths.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 3
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    sources = np.ascontiguousarray(sources, dtype=np.int32)
    n = len(indptr) - 1
    out = np.empty((len(sources), n), dtype=np.float64)
    info = np.zeros(4)
    _check(lib, lib.dys_shortest_paths(device, n, indptr.ctypes.data, indices.ctypes.data, weights.ctypes.data, len(sources),
                                       sources.ctypes.data, out.ctypes.data, info.ctypes.data))
    if return_info:
        return out, dict(sweeps=int(info[0]), prep_s=info[1], relax_s=info[2], download_s=info[3])
    return (out, int(info[0])) if return_sweeps else out


def shortest_paths_csr(indptr, indices, weights, sources, capacity=None, skip_self=True, device=0, return_info=False, out=None,
                       max_dist=np.inf, col_keep=None):
    """the same distances as CSR over the reachable vertices only: (ptr[len(sources)+1], col, dist), compacted on the device.
    out = (col_buffer, dist_buffer, offset): write into caller-owned arrays from `offset` on (the views returned alias them).
    max_dist: list (and search) only distances < max_dist; col_keep: uint8 [n] mask of the vertices to list."""
    lib = engine.load_library()
    lib.dys_shortest_paths_csr.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 3
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    sources = np.ascontiguousarray(sources, dtype=np.int32)
    keep = None if col_keep is None else np.ascontiguousarray(col_keep, dtype=np.uint8)
    if keep is not None and keep.shape != (len(indptr) - 1,):
        raise ValueError("col_keep must have one entry per vertex")
    ptr = np.zeros(len(sources) + 1, dtype=np.int64)
    if out is None:
        col, dist, off = np.empty(int(capacity), dtype=np.int32), np.empty(int(capacity), dtype=np.float64), 0
    else:
        col, dist, off = out
        if col.dtype != np.int32 or dist.dtype != np.float64 or not (col.flags.c_contiguous and dist.flags.c_contiguous) or len(col) != len(dist):
            raise TypeError("out = (int32[m], float64[m], offset), both contiguous")
    room = len(col) - off if capacity is None else min(int(capacity), len(col) - off)
    info = np.zeros(4)
    _check(lib, lib.dys_shortest_paths_csr(device, len(indptr) - 1, indptr.ctypes.data, indices.ctypes.data, weights.ctypes.data,
                                           len(sources), sources.ctypes.data, int(bool(skip_self)), float(max_dist),
                                           None if keep is None else keep.ctypes.data, ptr.ctypes.data, room,
                                           col.ctypes.data + 4 * off, dist.ctypes.data + 8 * off, info.ctypes.data))
    m = int(ptr[-1])
    res = (ptr, col[off:off + m], dist[off:off + m])
    return res + (dict(sweeps=int(info[0]), prep_s=info[1], relax_s=info[2], download_s=info[3]),) if return_info else res


def components(indptr, indices):
    """weakly connected components of the CSR graph (labels [n]); sources of one component share the vertices they reach"""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    n = len(indptr) - 1
    g = sp.csr_matrix((np.ones(len(indices), dtype=np.int8), indices, indptr), shape=(n, n))
    return connected_components(g, directed=True, connection="weak")[1]


def interaction_prob(indptr, indices, weights, device=0, rows_per_call=2048, row_range=None):
    """contagious3.py:124-130 for the agent block: interaction_prob = exp(-shortest path), 0 on the diagonal and where no
    path exists, as a scipy CSR matrix (sorted indices).  The sources go to the device in agent order and come back as
    compacted (vertex, distance) lists written straight into the final arrays, so neither an n x n matrix nor dense rows
    ever exist on the host and nothing is reordered; exp runs in place.
    row_range = (lo, hi): only the rows of agents lo..hi-1, as a (hi - lo) x n matrix.  Rows are independent: several GPUs
    (one process each, `device` = its GPU) take disjoint ranges and nothing is exchanged; scipy.sparse.vstack of the
    pieces in rank order is the whole matrix."""
    import scipy.sparse as sp
    n = len(indptr) - 1
    lo, hi = (0, n) if row_range is None else (int(row_range[0]), int(row_range[1]))
    if not 0 <= lo <= hi <= n:
        raise ValueError("row_range outside [0, %d]" % n)
    labels = components(indptr, indices)
    total = int((np.bincount(labels)[labels[lo:hi]] - 1).sum())  # a source reaches at most its weak component, less itself
    col, val = np.empty(total, dtype=np.int32), np.empty(total, dtype=np.float64)
    out_ptr = np.zeros(hi - lo + 1, dtype=np.int64)
    pos = 0
    for k0 in range(lo, hi, rows_per_call):
        src = np.arange(k0, min(hi, k0 + rows_per_call), dtype=np.int32)
        ptr, _, d = shortest_paths_csr(indptr, indices, weights, src, skip_self=True, device=device, out=(col, val, pos))
        out_ptr[k0 - lo + 1:k0 - lo + 1 + len(src)] = pos + ptr[1:]
        np.negative(d, out=d)
        np.exp(d, out=d)                                                                        # :127
        pos += int(ptr[-1])
    return sp.csr_matrix((val[:pos], col[:pos], out_ptr), shape=(hi - lo, n))


def create_interaction(agents, households, build, min_prob=0.95, w_a=1.0, w_i=1.0, w_d=1.0, device=0):
    """agents, households, build (create_data's arrays) -> interaction_prob, i.e. communities.create_network's agent edges,
    the all-pairs shortest paths and contagious3.py:127 in one go.  Returns (interaction_prob CSR, (indptr, col, weight))."""
    indptr, col, prob = similarity_edges(agent_columns(agents, households, build), min_prob, w_a, w_i, w_d, device)
    w = edge_weights(prob)
    return interaction_prob(indptr, col, w, device=device), (indptr, col, w)
