"""Epicentre selection on the device neighbour list (client/EpiCenters.cpp:52-112,188-258) against the reference's
O(N^2) minimum-image rule restated in numpy."""
import numpy as np

from eon_b200 import B200Potential
from eon_b200 import displace as DP
from eon_b200 import workloads as WL


def _coordination_bruteforce(R, cell, cutoff):
    L = np.diag(cell)
    d = R[:, None, :] - R[None, :, :]
    d -= np.round(d / L) * L
    r = np.linalg.norm(d, axis=2)
    np.fill_diagonal(r, np.inf)
    return (r < cutoff).sum(axis=1)


def test_coordination_and_least_coordinated_epicentre(backend):
    s = WL.load_structure("akmc_pt.json")
    ref = _coordination_bruteforce(s.positions, s.cell, 3.3)
    with B200Potential("morse_pt", lib=backend) as pot:
        c = DP.coordination(pot, s.positions, s.cell, 3.3)
        assert np.array_equal(c, ref)
        cand = DP.min_coordinated_atoms(pot, s.positions, s.cell, s.fixed, 3.3)
        free = s.fixed.sum(axis=1) == 0
        assert len(cand) >= 1 and free[cand].all() and (ref[cand] == ref[free].min()).all()
        assert set(cand) == set(np.where(free & (ref <= ref[free].min()))[0])
        picks = {DP.min_coordinated_epicentre(pot, s.positions, s.cell, s.fixed, np.random.default_rng(k)) for k in range(8)}
        assert picks <= set(int(i) for i in cand)
    assert DP.last_atom_epicentre(s.positions) == s.n_atoms - 1
    assert free[DP.random_free_atom_epicentre(s.fixed, s.n_atoms, np.random.default_rng(1))]
    # displacement around the epicentre: only free atoms within the radius move, the mode is the unit displacement
    e = int(cand[0])
    Rd, mode = DP.displace(s, s.positions, e, 3.3, 0.1, seed=5)
    moved = np.abs(Rd - s.positions).sum(axis=1) > 0
    assert moved[e] and free[moved].all() and abs(np.linalg.norm(mode) - 1.0) < 1e-12
