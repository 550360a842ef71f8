"""FD Hessian through one batched force call (client/Hessian.cpp:150-231) against the same formula evaluated with the
CPU oracle one configuration at a time."""
import numpy as np
import pytest

from eon_b200 import B200Potential
from eon_b200.hessian import fd_hessian


class _OraclePot:
    def __init__(self, oracle, kind):
        self.o, self.p = oracle, oracle.potential(kind)

    def forceBatch(self, S, N, R, Z, box):
        F = np.empty((S, N, 3))
        U = np.empty(S)
        for s in range(S):
            U[s], F[s] = self.o.force(self.p, np.asarray(R[s]).reshape(N, 3), Z, np.asarray(box).reshape(3, 3))
        return F, U, np.zeros(S)


@pytest.mark.parametrize("central", [False, True])
def test_lj13_hessian_matches_oracle(backend, oracle, central):
    d = oracle.load_fixture("lj13.json")
    R, Z, box = d["positions"], d["atomic_nrs"], d["cell"]
    atoms = [0, 3, 7]
    masses = np.linspace(1.0, 2.0, len(Z))
    H0 = fd_hessian(_OraclePot(oracle, oracle.POT_LJ), R, Z, box, atoms, masses=masses, central=central)
    with B200Potential("lj", lib=backend) as pot:
        H = fd_hessian(pot, R, Z, box, atoms, masses=masses, central=central)
        assert pot.forceCallCounter == 1 + (2 if central else 1) * 9  # one batched call, one count per system
    assert H.shape == (9, 9) and np.array_equal(H, H.T)
    assert np.abs(H - H0).max() <= 1e-7 * np.abs(H0).max()  # forces agree to 1e-10 relative, divided by dr = 0.01


def test_fixed_atoms_are_masked(backend, oracle):
    d = oracle.load_fixture("morse_pt_saddle.json")
    R, Z, box, fixed = d["positions"], d["atomic_nrs"], d["cell"], d["fixed"]
    free = [int(i) for i in np.where(fixed.sum(axis=1) == 0)[0]]
    H0 = fd_hessian(_OraclePot(oracle, oracle.POT_MORSE_PT), R, Z, box, free, fixed=fixed)
    with B200Potential("morse_pt", lib=backend) as pot:
        H = fd_hessian(pot, R, Z, box, free, fixed=fixed)
    assert H.shape == (3 * len(free), 3 * len(free))
    assert np.abs(H - H0).max() <= 1e-7 * max(np.abs(H0).max(), 1e-12)
    assert np.linalg.eigvalsh(H).min() > -1e-6  # the fixture's reactant is a minimum for its one free atom
