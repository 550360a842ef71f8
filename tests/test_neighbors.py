"""Bit-exact neighbour sets: the device cell-list build (eonb200_neighbors) against the reference's own vesin
(oracle/_ref/libvesin_ref.so, compiled from client/thirdparty/vesin) and the oracle's brute-force restatement
of vesin's final predicate.  Pair order is unspecified in vesin (sorted by i only), so sets are compared."""
import numpy as np
import pytest

from eon_b200 import B200Potential


def canon(pairs, shifts):
    return set(map(tuple, np.concatenate([np.asarray(pairs, dtype=np.int64), np.asarray(shifts).astype(np.int64)], axis=1).tolist()))


def reference_list(oracle, R, box, rc, full, per):
    if oracle.ref_vesin_available():
        return oracle.ref_vesin_neighbors(R, box, rc, full=full, periodic=per)[:2]
    return oracle.neighbors(R, box, rc, full=full, periodic=per)


def cases(oracle):
    rng = np.random.default_rng(3)
    out = []
    for fixture, rc in (("lj13.json", 15.0), ("morse_pt_saddle.json", 9.5), ("akmc_al.json", 5.555),
                        ("akmc_pt.json", 9.5), ("neb_al.json", 5.555)):
        d = oracle.load_fixture(fixture)
        out.append((fixture, d["positions"], d["cell"], rc, (1, 1, 1)))
    H = np.array([[6.0, 0, 0], [1.5, 5.0, 0], [0.7, -1.2, 7.0]])
    out.append(("triclinic-unwrapped", rng.uniform(-1, 2, (30, 3)) @ H, H, 4.5, (1, 1, 1)))
    out.append(("triclinic-slab", rng.uniform(0, 1, (30, 3)) @ H, H, 4.5, (1, 0, 1)))
    out.append(("open", rng.uniform(0, 1, (40, 3)) @ H, H, 3.0, (0, 0, 0)))
    small = np.diag([3.0, 3.5, 4.0])
    out.append(("box-smaller-than-cutoff", rng.uniform(0, 1, (6, 3)) @ small, small, 7.0, (1, 1, 1)))
    out.append(("single-atom-self-images", np.array([[0.3, 0.2, 0.1]]), small, 7.0, (1, 1, 1)))
    # atoms exactly on cell faces and at exactly the cutoff distance (strict < must reject them)
    cube = np.diag([8.0, 8.0, 8.0])
    edge = np.array([[0.0, 0.0, 0.0], [8.0, 0.0, 0.0], [4.0, 0.0, 0.0], [0.0, 3.0, 0.0], [0.0, 0.0, 8.0 - 3.0]])
    out.append(("faces-and-exact-cutoff", edge, cube, 3.0, (1, 1, 1)))
    return out


@pytest.mark.parametrize("full", [True, False])
def test_sets_match_vesin(backend, oracle, full):
    with B200Potential("lj", lib=backend) as pot:
        for name, R, box, rc, per in cases(oracle):
            Pr, Sr = reference_list(oracle, R, box, rc, full, per)
            P, S = pot.neighbors(R, box, rc, full=full, periodic=per)
            assert len(P) == len(Pr), name
            assert canon(P, S) == canon(Pr, Sr), name


def test_batched_systems_and_vectors(backend, oracle):
    """Several different systems in one launch; distances / vectors equal vesin's bit for bit."""
    d = oracle.load_fixture("akmc_al.json")
    rng = np.random.default_rng(8)
    S = 4
    R = d["positions"][None] + rng.uniform(-0.3, 0.3, (S,) + d["positions"].shape)
    with B200Potential("eam", lib=backend) as pot:
        P, Sh, Dist, Vec = pot.neighborsBatch(R, d["cell"], 5.555, full=True, return_vectors=True)
    for s in range(S):
        if oracle.ref_vesin_available():
            Pr, Sr, Dr, Vr = oracle.ref_vesin_neighbors(R[s], d["cell"], 5.555, full=True, vectors=True)
        else:
            Pr, Sr = oracle.neighbors(R[s], d["cell"], 5.555, full=True)
            Dr = Vr = None
        assert canon(P[s], Sh[s]) == canon(Pr, Sr)
        if Dr is not None:
            key = lambda p, sh: [tuple(a) + tuple(b) for a, b in zip(p.tolist(), sh.tolist())]
            mine = dict(zip(key(P[s], Sh[s]), zip(Dist[s].tolist(), map(tuple, Vec[s].tolist()))))
            for k, dd, vv in zip(key(Pr, Sr), Dr.tolist(), map(tuple, Vr.tolist())):
                assert mine[k] == (dd, vv)


def test_sortedness_and_symmetry_at_full_size(backend, oracle):
    """Size-independent properties on the C3-sized cell (2047 atoms): rows ordered by i, full list symmetric
    ((i,j,S) <-> (j,i,-S)), half list is exactly half of it, every atom of the perfect lattice has 54
    neighbours except the vacancy's shell."""
    from eon_b200 import workloads as WL
    s = WL.fcc_vacancy(8, 3.615, 13)
    with B200Potential("eam", lib=backend) as pot:
        P, S = pot.neighbors(s.positions, s.cell, 5.555, full=True)
        Ph, _ = pot.neighbors(s.positions, s.cell, 5.555, full=False)
    assert np.all(np.diff(P[:, 0]) >= 0)
    fwd = canon(P, S)
    back = canon(P[:, ::-1], -S)
    assert fwd == back
    assert 2 * len(Ph) == len(P)
    counts = np.bincount(P[:, 0], minlength=s.n_atoms)
    assert counts.max() == 54 and (counts == 54).sum() == s.n_atoms - 54 and counts.min() == 53
