"""Golden vectors for the network construction, from the UNMODIFIED reference (needs /root/reference; run
in the authoring container only):   python tests/golden/make_golden_network.py

A household sub-sample of the shipped city through the reference's own create_data -> communities.create_network ->
scipy shortest_path -> interaction_prob (contagious3.py:119-130, executed literally by oracle/literal.py).  Stored: the
three input tables, the agent-agent edges of the reference's graph G (agent row indices and the weights -log p exactly as
G holds them), interaction_prob as CSR, and the whole of G (node order, every edge by node id with its weight).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import literal  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    assert literal.available(), "needs /root/reference"
    ind, bld = literal.subsample_csv("/tmp/dys_golden_network", 300, 1, force_households=(0,))
    setup = literal.reference_setup(ind, bld, seed=7)
    agents, households, build, G = setup["agents"], setup["households"], setup["build"], setup["G"]
    row_of = {a: i for i, a in enumerate(agents[:, 0])}
    edges = [(row_of[u], row_of[v], d["weight"]) for u, v, d in G.edges(data=True) if u in row_of and v in row_of]
    edges.sort(key=lambda e: (e[0], e[1]))
    # every edge of G by node id, in G's own order (buildings < 4 000 000 <= agents), and G's node order (= row order of `dists`)
    all_edges = [(u, v, d["weight"]) for u, v, d in G.edges(data=True)]
    nodes = np.array(G.nodes(), dtype=np.float64)
    ip = sp.csr_matrix(setup["interaction_prob"])
    ip.sort_indices()
    np.savez_compressed(os.path.join(HERE, "network_city300.npz"), agents=agents, households=households, build=build,
                        edge_row=np.array([e[0] for e in edges], dtype=np.int32), edge_col=np.array([e[1] for e in edges], dtype=np.int32),
                        edge_w=np.array([e[2] for e in edges], dtype=np.float64),
                        g_nodes=nodes, g_u=np.array([e[0] for e in all_edges], dtype=np.float64),
                        g_v=np.array([e[1] for e in all_edges], dtype=np.float64), g_w=np.array([e[2] for e in all_edges], dtype=np.float64),
                        ip_indptr=ip.indptr.astype(np.int64), ip_indices=ip.indices.astype(np.int32), ip_data=ip.data)
    print("agents", len(agents), "households", len(households), "buildings", len(build), "agent-agent edges", len(edges), "ip nnz", ip.nnz)
