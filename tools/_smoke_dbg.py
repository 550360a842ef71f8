import sys, os, faulthandler
faulthandler.enable()
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
which = sys.argv[1]
from eon_b200 import B200Potential, Walkers, search_params, workloads as WL
if which in ("eam", "both"):
    s = WL.load_structure("akmc_al.json")
    R = WL.jitter(s.positions, 4, 0.05, seed=7)
    with B200Potential("eam", device=0) as pot:
        F, U, _ = pot.forceBatch(4, s.n_atoms, R.reshape(4, -1), s.atomic_nrs, s.cell.reshape(9))
    print("eam done", flush=True)
if which in ("morse", "both"):
    import oracle_lib as O
    d = O.load_fixture("morse_pt_saddle.json")
    with B200Potential("morse_pt", device=0) as pot:
        wk = Walkers(pot, d["atomic_nrs"], d["fixed"], d["cell"], d["displacement_positions"].reshape(1, -1), d["direction"].reshape(1, -1))
        wk.saddle_search(search_params(), -1462.0)
        h = wk.harvest()
        if which == "morse" and len(sys.argv) > 2: wk.close()
    print("morse done", flush=True)
if which == "oracle":
    import oracle_lib as O
    opot = O.potential(O.POT_EAM)
    print("oracle done", flush=True)
