"""Generates tests/golden/routines_city120.npz from the UNMODIFIED reference (needs /root/reference; authoring container
only):   python tests/golden/make_golden_routines.py
The reference's own create_data -> create_network -> shortest_path -> routine loop (contagious3.py:119-170) run on a
120-household sub-sample of the shipped city with keyed draws injected into the routine loop (oracle/literal._RoutineRandom);
stored: the neighbourhood lists dysturbd_b200.routines.neighbourhoods() derives from the reference's `dists`, the
reference's bld_visits_by_agents, and the three input tables (agents, households, build)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dysturbd_b200 import routines  # noqa: E402
from oracle import literal  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main(name="routines_city120", n_households=120, csv_seed=5, setup_seed=11, routine_seed=777):
    ind, bld = literal.subsample_csv("/tmp/dys_golden_" + name, n_households, csv_seed)
    s = literal.reference_setup(ind, bld, seed=setup_seed, routine_seed=routine_seed)
    prm = literal._import_reference()[0]
    nb = routines.neighbourhoods(s["nodes"], s["dists"], s["agents"], s["build"], prm.a_dist, prm.bld_dist)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), bld_ids=nb["bld_ids"], home=nb["home"], anchors=nb["anchors"],
                        out_ptr=nb["out"][0], out_idx=nb["out"][1], in_ptr=nb["inn"][0], in_idx=nb["inn"][1],
                        near_ptr=nb["near"][0], near_idx=nb["near"][1], seed=routine_seed, k=prm.k,
                        visits=s["bld_visits_by_agents"],
                        # the inputs of create_network, so that the whole chain (graph -> distances -> lists -> routines) can be rebuilt
                        agents=s["agents"], households=s["households"], build=s["build"])
    print(name, "agents", len(nb["home"]), "buildings", len(nb["bld_ids"]), "list entries",
          len(nb["out"][1]), len(nb["inn"][1]), len(nb["near"][1]))


if __name__ == "__main__":
    main()
