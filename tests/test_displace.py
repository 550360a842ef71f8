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


def _philox4x32_10(c, k):
    """Reference Philox4x32-10 (Salmon et al. 2011) in Python integers."""
    c = list(c)
    k = list(k)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors): zero input and the 0xff.. input."""
    assert _philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_device_displacements(backend, oracle):
    """`Walkers.displace` (k_displace): the semantics of eon/displace.py:180-263 -- only free axes of the epicentre and
    of the atoms within the radius move, normal deviates of the requested standard deviation, mode = normalised
    displacement -- and the exact numbers of a Philox4x32-10 + Box-Muller restatement in Python, per (walker, atom),
    independent of how the walkers are split over handles."""
    from eon_b200 import Walkers
    d = oracle.load_fixture("akmc_pt.json")
    Z, box, fixed, R0 = d["atomic_nrs"], d["cell"], d["fixed"], d["positions"]
    N = len(Z)
    W, radius, sigma, seed = 6, 3.3, 0.1, 0x1234567890ABCDEF
    free_atoms = np.where(fixed.sum(axis=1) == 0)[0]
    epi = free_atoms[[0, 5, 9, 20, 33, 47]].astype(np.int32)
    with B200Potential("morse_pt", lib=backend) as pot:
        wk = Walkers(pot, Z, fixed, box, np.broadcast_to(R0.reshape(1, -1), (W, 3 * N)).copy())
        wk.displace(R0, epi, radius, sigma, seed, first_walker_id=100)
        h = wk.harvest()
        wk2 = Walkers(pot, Z, fixed, box, np.broadcast_to(R0.reshape(1, -1), (2, 3 * N)).copy())
        wk2.displace(R0, epi[4:], radius, sigma, seed, first_walker_id=104)  # walkers 104, 105 on another handle
        h2 = wk2.harvest()
        wk.close()
        wk2.close()
    assert np.array_equal(h.positions[4:], h2.positions) and np.array_equal(h.modes[4:], h2.modes)
    L = np.diag(box)
    for w in range(W):
        disp = h.positions[w] - R0
        dd = R0 - R0[epi[w]]
        dd -= np.round(dd / L) * L
        near = (np.linalg.norm(dd, axis=1) < radius)
        near[epi[w]] = True
        moved = np.abs(disp).sum(axis=1) > 0
        assert np.all(disp[fixed.astype(bool)] == 0.0)
        assert not np.any(moved & ~near) and np.all(moved[near & (fixed.sum(axis=1) == 0)])
        assert abs(np.linalg.norm(h.modes[w]) - 1.0) < 1e-12
        assert np.abs(h.modes[w] - disp / np.linalg.norm(disp)).max() < 1e-9
        # exact numbers: counter (atom, walker id lo, walker id hi, 0), key = seed
        for i in np.where(near)[0][:8]:
            r = _philox4x32_10([int(i), (100 + w) & 0xFFFFFFFF, (100 + w) >> 32, 0], [seed & 0xFFFFFFFF, seed >> 32])
            u = [(x + 0.5) * 2.0 ** -32 for x in r]
            ra, rb = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
            g = np.array([ra * np.cos(2 * np.pi * u[1]), ra * np.sin(2 * np.pi * u[1]), rb * np.cos(2 * np.pi * u[3])])
            assert np.abs(disp[i] - sigma * g * (1 - fixed[i])).max() < 1e-12
    allmoved = np.concatenate([(h.positions[w] - R0)[np.abs(h.positions[w] - R0).sum(axis=1) > 0].ravel() for w in range(W)])
    allmoved = allmoved[allmoved != 0.0]
    assert abs(allmoved.std() / sigma - 1.0) < 0.15 and abs(allmoved.mean()) < 0.03
