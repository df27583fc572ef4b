#!/usr/bin/env python
"""Developer tool (for ncu): one pass of the network construction on the shipped city -- the edge kernels, then the
shortest paths of 1 024 sources of the largest component."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dysturbd_b200 import network  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "config1", "config1_inputs.npz"))
cols = network.agent_columns(z["agents"], z["households"], z["build"])
indptr, col, prob = network.similarity_edges(cols)
w = network.edge_weights(prob)
labels = network.components(indptr, col)
big = np.nonzero(labels == np.bincount(labels).argmax())[0][:1024].astype(np.int32)
ptr, c, d, info = network.shortest_paths_csr(indptr, col, w, big, capacity=len(big) * int((labels == labels[big[0]]).sum()), return_info=True)
print(len(c), info)
