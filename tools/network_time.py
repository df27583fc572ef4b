#!/usr/bin/env python
"""Developer tool: the network construction of the full shipped city (22 493 agents) on the GPU -- edges, all-pairs
shortest paths, interaction_prob -- timed next to the oracle (OpenMP Dijkstra) and scipy (the reference's call) on a sample."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import shortest_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dysturbd_b200 import network  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "config1", "config1_inputs.npz"))
n = len(z["agents"])
cols = network.agent_columns(z["agents"], z["households"], z["build"])
network.similarity_edges(dict((k, v[:64]) for k, v in cols.items()))          # (context creation, module load)
t = time.perf_counter()
indptr, col, prob = network.similarity_edges(cols)
t_edges = time.perf_counter() - t
w = network.edge_weights(prob)
res = {"agents": n, "edges": int(indptr[-1]), "edges_s": t_edges}
order = np.argsort(network.components(indptr, col), kind="stable").astype(np.int32)
def apsp(srcs, tag):
    t = time.perf_counter()
    tot = dict(sweeps=0, prep_s=0.0, relax_s=0.0, download_s=0.0)
    for k0 in range(0, n, 2048):
        d, info = network.shortest_paths(indptr, col, w, srcs[k0:k0 + 2048], return_info=True)
        for k in tot:
            tot[k] += info[k]
    res[tag] = dict(total_s=time.perf_counter() - t, **tot)


def apsp_csr(tag):
    labels = network.components(indptr, col)
    reach = np.bincount(labels)[labels] - 1
    t = time.perf_counter()
    tot = dict(sweeps=0, prep_s=0.0, relax_s=0.0, download_s=0.0)
    nnz = 0
    for k0 in range(0, n, 2048):
        src = order[k0:k0 + 2048]
        ptr, c, d, info = network.shortest_paths_csr(indptr, col, w, src, int(reach[src].sum()), return_info=True)
        nnz += len(c)
        for k in tot:
            tot[k] += info[k]
    res[tag] = dict(total_s=time.perf_counter() - t, pairs=nnz, **tot)


apsp_csr("apsp_csr_by_component")
apsp(order, "apsp_by_component")
apsp(np.arange(n, dtype=np.int32), "apsp_agent_order")
if "--ip" in sys.argv:
    t = time.perf_counter()
    ip = network.interaction_prob(indptr, col, w)
    res.update(interaction_prob_s=time.perf_counter() - t, ip_nnz=int(ip.nnz))
src = order[4000:4512]
g = sp.csr_matrix((w, col, indptr), shape=(n, n))
t = time.perf_counter()
ref = shortest_path(g, directed=True, indices=src)
res["scipy_512_sources_s"] = time.perf_counter() - t
t = time.perf_counter()
got = network.shortest_paths(indptr, col, w, src)
res["gpu_512_sources_s"] = time.perf_counter() - t
res["equal_scipy"] = bool(np.array_equal(got, ref))
try:
    from oracle import port
    t = time.perf_counter()
    o = port.shortest_paths(indptr, col, w, src)
    res.update(oracle_512_sources_s=time.perf_counter() - t, oracle_threads=port.max_threads(), equal_oracle=bool(np.array_equal(got, o)))
    t = time.perf_counter()
    port.similarity_edges(cols["age"], cols["income"], cols["x"], cols["y"], cols["household"], 0.95)
    res["oracle_edges_s"] = time.perf_counter() - t
except Exception as e:                                                        # noqa: BLE001
    res["oracle"] = repr(e)
print(json.dumps(res))
