"""TEST INFRASTRUCTURE (oracle) -- generates the config-1 (shipped city) reference artefacts.

Runs ONLY in the authoring container (needs /root/reference).  Produces, under tests/golden/config1/ (committed; the
8 GB-class interaction_prob cache goes to oracle/_ref/local/, git-ignored):
  config1_inputs.npz          agents[N,29], build[B,15], bld_visits_by_agents[N,10], agent-agent edge list of
                              the communities.py graph (so interaction_prob can be rebuilt without the reference)
  config1_<scenario>_golden.npz   per-day state columns + the reference's own Stats/SAs/Buildings/IO_mat output,
                              produced by the literal contagious3.py loop under keyed Philox draws
  local/config1_ip_csr.npz    interaction_prob as CSR (1.8 GB; not shipped to the GPU box, see .gpurunignore)

usage: python -m oracle.make_config1_ref [--days 100] [--seed 2020] [--philox-seed 20201019] [--scenario noLockdown]
"""
import argparse
import json
import os
import time

import numpy as np

from . import literal

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
OUT = os.path.normpath(os.path.join(HERE, "..", "tests", "golden", "config1"))


def agent_edges(G, agents):
    ids = agents[:, 0]
    idset = {float(x): i for i, x in enumerate(ids)}
    r, c, w = [], [], []
    for u, v, d in G.edges(data="weight"):
        if u in idset and v in idset:
            r.append(idset[u]); c.append(idset[v]); w.append(d)
    return np.array(r, np.int32), np.array(c, np.int32), np.array(w, np.float64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=100)
    ap.add_argument("--seed", type=int, default=2020)
    ap.add_argument("--philox-seed", type=int, default=20201019)
    ap.add_argument("--scenario", default="noLockdown")
    args = ap.parse_args()
    os.makedirs(os.path.join(REF, "local"), exist_ok=True)
    os.makedirs(OUT, exist_ok=True)
    ind = os.path.join(literal.REFERENCE_DIR, "Data", "civ_withCar_bldg_np.csv")
    bld = os.path.join(literal.REFERENCE_DIR, "Data", "bldg_with_inst_orig.csv")
    t0 = time.time()
    s = literal.reference_setup(ind, bld, seed=args.seed, verbose=True)
    print("setup", time.time() - t0, s["timings"], flush=True)
    er, ec, ew = agent_edges(s["G"], s["agents"])
    np.savez_compressed(os.path.join(OUT, "config1_inputs.npz"), agents=s["agents"], build=s["build"],
                        bld_visits_by_agents=s["bld_visits_by_agents"], households=np.asarray(s["households"], np.float64),
                        edge_row=er, edge_col=ec, edge_w=ew, seed=args.seed,
                        timings=json.dumps(s["timings"]))
    import scipy.sparse as sp
    ip = sp.csr_matrix(s["interaction_prob"])
    np.savez(os.path.join(REF, "local", "config1_ip_csr.npz"), indptr=ip.indptr.astype(np.int64),
             indices=ip.indices.astype(np.int32), data=ip.data)
    print("saved inputs, ip nnz", ip.nnz, flush=True)
    del ip
    N = len(s["agents"])
    scenario = args.scenario.split("_")
    t0 = time.time()
    r = literal.reference_days(s, scenario, args.days, literal.PhiloxDraws(N, args.philox_seed), verbose=False)
    print("days", len(r["snapshots"]), time.time() - t0, flush=True)
    snaps = r["snapshots"]
    out = {
        "status_x2": np.array([np.round(x["status"] * 2) for x in snaps]).astype(np.uint8),
        "t_inf": np.array([x["t_inf"] for x in snaps]).astype(np.int16),
        "t_q": np.array([x["t_q"] for x in snaps]).astype(np.int16),
        "t_adm": np.array([x["t_adm"] for x in snaps]).astype(np.int16),
        "src_id": np.array([x["src_id"] for x in snaps]).astype(np.int64),
        "open_next": np.array([x["open_next"] for x in snaps]).astype(np.uint8),
        "day_seconds": np.array(r["day_seconds"]),
        "philox_seed": args.philox_seed,
        "outputs_json": json.dumps(r["outputs"], default=lambda o: o.item() if hasattr(o, "item") else str(o)),
    }
    np.savez_compressed(os.path.join(OUT, "config1_%s_golden.npz" % args.scenario), **out)
    print("done", flush=True)


if __name__ == "__main__":
    main()
