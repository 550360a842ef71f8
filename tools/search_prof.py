"""One launch of 256 resident dimer + LBFGS searches on the bench's search workload (for ncu).  Development aid.
    python tools/search_prof.py [lanczos] [walkers=W] [maxit=K]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
import bench  # noqa: E402
from eon_b200 import B200Potential, Walkers, search_params, workloads as WL  # noqa: E402

opt = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
W = int(opt.get("walkers", 256))
s2, scale = bench.search_workload()
N = s2.n_atoms
spot = B200Potential("eam", device=0, skin=1.0, conservative=True)
wk = Walkers(spot, s2.atomic_nrs, s2.fixed, s2.cell, np.broadcast_to(s2.positions.reshape(1, -1), (W, 3 * N)).copy())
wk.relax(search_params(opt_method=1, opt_converged_force=1e-4))
hmin = wk.harvest()
Rmin, E0 = hmin.positions[0], float(hmin.energies[0])
cfg = dict(WL.AKMC_AL)
wk.displace(Rmin, N - 1, cfg["radius"] * scale, cfg["magnitude"], cfg["seed"], first_walker_id=0)
params = dict(bench.SEARCH_PARAMS if "lanczos" in sys.argv else bench.SEARCH_PARAMS_DIMER)
if "maxit" in opt:
    params["saddle_max_iterations"] = int(opt["maxit"])
wk.saddle_search(search_params(**params), E0)
h = wk.harvest(geometry=False)
print("ms", wk.last_ms, "good", int((h.status == 0).sum()), "force calls", int(h.force_calls.sum()),
      "atom-force-evals/s %.3e" % (h.force_calls.sum() * N / (wk.last_ms * 1e-3)))
