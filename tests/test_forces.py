"""Energy / force parity of the batched potentials against the CPU oracle and the reference goldens.

Tolerances are the north star's: forces 1e-10 relative (to the largest force component of the
configuration), energies 1e-8 eV.
"""
import numpy as np
import pytest

from eon_b200 import B200Potential, PotentialError, eam_row

F_RTOL = 1e-10
E_ATOL = 1e-8

CASES = [("lj", "POT_LJ", "lj13.json", {}), ("lj_cluster", "POT_LJCLUSTER", "lj13.json", {}),
         ("morse_pt", "POT_MORSE_PT", "morse_pt_saddle.json", {}), ("morse_pt", "POT_MORSE_PT", "akmc_pt.json", {}),
         ("eam", "POT_EAM", "al_fcc.json", {}), ("eam", "POT_EAM", "akmc_al.json", {}),
         ("eam", "POT_EAM", "akmc_al.json", {"conservative": True}), ("eam", "POT_EAM", "neb_al.json", {})]


def assert_parity(F, U, F0, U0):
    assert abs(U - U0) < max(E_ATOL, 8 * np.spacing(abs(U0)))  # 1e-8 eV, or a few ulp for overlapping atoms
    assert np.abs(F - F0).max() <= F_RTOL * max(np.abs(F0).max(), 1e-300)


@pytest.mark.parametrize("kind,okind,fixture,kw", CASES)
@pytest.mark.parametrize("skin", [0.0, 1.0])
def test_force_matches_oracle(backend, oracle, kind, okind, fixture, kw, skin):
    d = oracle.load_fixture(fixture)
    R, Z, box = d["positions"], d["atomic_nrs"], d["cell"]
    U0, F0 = oracle.force(oracle.potential(getattr(oracle, okind), **kw), R, Z, box)
    with B200Potential(kind, lib=backend, skin=skin, **kw) as pot:
        F, U, var = pot.force(len(Z), R, Z, box)
        assert var == 0.0
        assert_parity(F, U, F0, U0)
        assert pot.forceCallCounter == 1


def test_goldens(backend, oracle):
    """Known-answer values of the reference's own tests (tests/golden/goldens.json)."""
    g = oracle.goldens()
    d = oracle.load_fixture("lj13.json")
    with B200Potential("lj", lib=backend) as pot:
        F, U, _ = pot.force(13, d["positions"], d["atomic_nrs"], d["cell"])
        assert U == pytest.approx(g["lj_point"]["energy"], abs=5e-7)
        assert np.linalg.norm(F, axis=1).max() == pytest.approx(g["lj_point"]["max_force"], abs=5e-7)
    with B200Potential("lj_cluster", lib=backend) as pot:
        _, U, _ = pot.force(13, d["positions"], d["atomic_nrs"], d["cell"])
        assert U == pytest.approx(g["ljcluster_point"]["energy"], abs=5e-7)
    d = oracle.load_fixture("morse_pt_saddle.json")
    with B200Potential("morse_pt", lib=backend) as pot:
        F, U, _ = pot.force(len(d["atomic_nrs"]), d["positions"], d["atomic_nrs"], d["cell"])
        assert U == pytest.approx(g["morse_pt_point"]["energy"], abs=5e-7)
        free = d["fixed"].sum(axis=1) == 0
        assert np.linalg.norm(F[free], axis=1).max() == pytest.approx(g["morse_pt_point"]["max_force"], abs=5e-7)


def test_batch_equals_single(backend, oracle):
    """Batch contract of client/unit_tests/RgpotAdapterTest.cpp:168-218: two different systems through
    forceBatch equal two force calls (1e-10 there).  Not bit-identical by design: a cached candidate list
    built from a neighbouring geometry visits the same pairs in another order."""
    a = oracle.load_fixture("lj13.json")
    b = oracle.load_fixture("lj13_product.json")
    with B200Potential("lj", lib=backend) as pot:
        Fb, Ub, _ = pot.forceBatch(2, 13, np.stack([a["positions"], b["positions"]]).reshape(2, -1),
                                   a["atomic_nrs"], a["cell"])
        for k, d in enumerate((a, b)):
            F, U, _ = pot.force(13, d["positions"], d["atomic_nrs"], d["cell"])
            assert np.abs(F - Fb[k]).max() <= 1e-12 * np.abs(F).max() and abs(U - Ub[k]) < 1e-10
        assert pot.forceCallCounter == 4


@pytest.mark.parametrize("kind,okind,fixture,amp", [("eam", "POT_EAM", "akmc_al.json", 0.08),
                                                    ("morse_pt", "POT_MORSE_PT", "akmc_pt.json", 0.05),
                                                    ("lj", "POT_LJ", "lj13.json", 0.05)])
def test_jittered_batch_and_cache_reuse(backend, oracle, kind, okind, fixture, amp):
    """Seeded jitter, several walkers, repeated calls: the Verlet-cached list (reused, then rebuilt after a
    large move) never changes the result (NL cache invariance, client/unit_tests/PotTest.cpp:60-105)."""
    d = oracle.load_fixture(fixture)
    Z, box, N = d["atomic_nrs"], d["cell"], len(d["atomic_nrs"])
    rng = np.random.default_rng(11)
    opot = oracle.potential(getattr(oracle, okind))
    S = 5
    base = d["positions"][None] + rng.uniform(-amp, amp, (S, N, 3))
    with B200Potential(kind, lib=backend, skin=1.0) as pot:
        seen_reuse = seen_rebuild = False
        for step, scale in enumerate((0.0, 0.02, 0.05, 0.9, 0.01)):
            R = base + scale * rng.uniform(-1, 1, base.shape)
            F, U, _ = pot.forceBatch(S, N, R.reshape(S, -1), Z, box)
            st = pot.nl_stats()
            seen_reuse |= st["reused"] > 0
            seen_rebuild |= step > 0 and st["rebuilt"] > 0
            for s in range(S):
                U0, F0 = oracle.force(opot, R[s], Z, box)
                assert_parity(F[s], U[s], F0, U0)
        assert seen_reuse and seen_rebuild


def test_eam_multi_element_and_custom_rows(backend, oracle):
    """Mixed Al/Pd/Au systems exercise the per-atom rows (parameters of atom i, EAM.cpp:140; pair part with
    the row of the lower index, :200); a caller-supplied table replaces the built-in rows (the Cu-row hook)."""
    d = oracle.load_fixture("akmc_al.json")
    Z = d["atomic_nrs"].copy()
    rng = np.random.default_rng(5)
    Z[rng.random(len(Z)) < 0.3] = 46
    Z[rng.random(len(Z)) < 0.2] = 79
    R = d["positions"] + rng.uniform(-0.05, 0.05, d["positions"].shape)
    for cons in (False, True):
        U0, F0 = oracle.force(oracle.potential(oracle.POT_EAM, conservative=cons), R, Z, d["cell"])
        with B200Potential("eam", lib=backend, conservative=cons) as pot:
            F, U, _ = pot.force(len(Z), R, Z, d["cell"])
            assert_parity(F, U, F0, U0)
    coeff = [-0.1, 9.0, 20.0, -30.0, 12.0, -2.0, -1.0, 0.5, -0.05]
    orow = oracle.EamRow()
    orow.Z, orow.Dm, orow.alphaM, orow.Rm, orow.beta1, orow.beta2, orow.r_cut = 29, 1.9, 1.6, 2.3, 3.4, 3.5, 5.1
    for k, c in enumerate(coeff):
        orow.func_coeff[k] = c
    Zc = np.full_like(Z, 29)
    U0, F0 = oracle.force(oracle.potential(oracle.POT_EAM, rows=[orow]), R, Zc, d["cell"])
    with B200Potential("eam", lib=backend, eam_rows=[eam_row(29, 1.9, 1.6, 2.3, 3.4, 3.5, 5.1, coeff)]) as pot:
        F, U, _ = pot.force(len(Zc), R, Zc, d["cell"])
        assert_parity(F, U, F0, U0)


def test_eam_matches_compiled_reference(backend, oracle):
    """Against the reference's own EAM.cpp compiled from the reference tree (oracle/_ref/libeam_ref.so)."""
    if not oracle.ref_eam_available():
        pytest.skip("oracle/_ref/libeam_ref.so not built")
    d = oracle.load_fixture("akmc_al.json")
    rng = np.random.default_rng(2)
    R = d["positions"] + rng.uniform(-0.05, 0.05, d["positions"].shape)
    U0, F0 = oracle.ref_eam_force(R, d["atomic_nrs"], d["cell"])
    with B200Potential("eam", lib=backend) as pot:
        F, U, _ = pot.force(len(R), R, d["atomic_nrs"], d["cell"])
        assert_parity(F, U, F0, U0)


def test_unknown_element_is_14324(backend, oracle):
    d = oracle.load_fixture("al_fcc.json")
    Z = d["atomic_nrs"].copy()
    Z[1] = 29
    with B200Potential("eam", lib=backend) as pot:
        with pytest.raises(PotentialError) as exc:
            pot.force(len(Z), d["positions"], Z, d["cell"])
        assert exc.value.code == 14324


def test_positions_outside_cell_and_negative(backend, oracle):
    """Unwrapped input: EAM shifts / wraps like EAM.cpp:56-90, the pair potentials take vesin's images."""
    for kind, okind, fixture in (("eam", "POT_EAM", "akmc_al.json"), ("morse_pt", "POT_MORSE_PT", "akmc_pt.json")):
        d = oracle.load_fixture(fixture)
        rng = np.random.default_rng(9)
        L = np.diag(d["cell"])
        R = d["positions"] + rng.integers(-1, 2, d["positions"].shape) * L * (rng.random(d["positions"].shape) < 0.1)
        if kind == "eam":
            R = d["positions"] - 0.4  # negative coordinates: the reference shifts everything by |min|
        U0, F0 = oracle.force(oracle.potential(getattr(oracle, okind)), R, d["atomic_nrs"], d["cell"])
        with B200Potential(kind, lib=backend) as pot:
            F, U, _ = pot.force(len(R), R, d["atomic_nrs"], d["cell"])
            assert_parity(F, U, F0, U0)


def test_triclinic_pair_potential(backend, oracle):
    H = np.array([[21.0, 0, 0], [3.0, 20.5, 0], [-2.0, 1.5, 22.0]])
    rng = np.random.default_rng(4)
    n = 60
    frac = rng.random((n, 3))
    R = frac @ H
    Z = np.full(n, 78, dtype=np.int32)
    U0, F0 = oracle.force(oracle.potential(oracle.POT_MORSE_PT), R, Z, H)
    with B200Potential("morse_pt", lib=backend) as pot:
        F, U, _ = pot.force(n, R, Z, H)
        assert_parity(F, U, F0, U0)


def test_even_and_odd_atom_counts(backend, oracle):
    """3N even takes the bulk-copy (TMA) staging path, 3N odd the plain one; both must agree with the oracle."""
    d = oracle.load_fixture("akmc_al.json")
    for n in (384, 385, 2, 1):
        R, Z = d["positions"][:n], d["atomic_nrs"][:n]
        U0, F0 = oracle.force(oracle.potential(oracle.POT_EAM), R, Z, d["cell"])
        with B200Potential("eam", lib=backend) as pot:
            F, U, _ = pot.forceBatch(3, n, np.stack([R] * 3).reshape(3, -1), Z, d["cell"])
            for s in range(3):
                assert_parity(F[s], U[s], F0, U0)


@pytest.mark.parametrize("kw", [
    # ordered-pair kernel (eam_pair_mode=1): cluster sizes and atoms per thread
    dict(eam_pair_mode=1, tiles_override=1, threads_per_walker=64),   # several atoms per thread, one CTA
    dict(eam_pair_mode=1, tiles_override=2, threads_per_walker=32),   # ... in a two-CTA cluster
    dict(eam_pair_mode=1, tiles_override=4), dict(eam_pair_mode=1, tiles_override=8),  # one atom per thread, wide clusters
    # half-list kernel (default): CTA sizes below, at and above the atom count (the half list is laid out per size)
    dict(threads_per_walker=64), dict(threads_per_walker=224), dict(), dict(threads_per_walker=512),
    dict(threads_per_walker=1024)])
def test_fused_eam_launch_shapes(backend, oracle, kw):
    """Every shape of the two EAM kernels gives the oracle's forces, with a fresh list, a reused list, a list whose
    atoms drifted beyond skin / 4 (every candidate screened again) and a list rebuilt inside the kernel after a large
    displacement."""
    from eon_b200 import workloads as WL
    s = WL.load_structure("akmc_al.json")
    opot = oracle.potential(oracle.POT_EAM)
    R = WL.jitter(s.positions, 4, 0.04, seed=11)
    R[2] += 0.9   # beyond skin / 2: in-kernel rebuild; the rigid shift also moves atoms across the cell faces
    R[3] = R[0] + 0.18  # between skin / 4 and skin / 2 of the list built from R[0]: every row is screened again
    with B200Potential("eam", lib=backend, skin=1.0, **kw) as pot:
        for w in (0, 1, 3, 2, 1):
            F, U, _ = pot.forceBatch(2, s.n_atoms, np.stack([R[w], R[0]]).reshape(2, -1), s.atomic_nrs, s.cell.reshape(9))
            for k, ww in enumerate((w, 0)):
                U0, F0 = oracle.force(opot, R[ww], s.atomic_nrs, s.cell)
                assert_parity(F[k], U[k], F0, U0)


def test_caller_owned_pinned_buffers(backend, oracle):
    """pin_host_buffers + out=: the page-locked direct path returns what the staged path returns."""
    from eon_b200 import workloads as WL
    s = WL.load_structure("akmc_al.json")
    R = np.ascontiguousarray(WL.jitter(s.positions, 5, 0.05, seed=3).reshape(5, -1))

    def same(a, b):  # the half-list kernel's shared-memory accumulation order is free: equal to rounding
        a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
        return np.abs(a - b).max() <= 1e-13 * np.abs(b).max()

    with B200Potential("eam", lib=backend) as pot:
        F0, U0, _ = pot.forceBatch(5, s.n_atoms, R, s.atomic_nrs, s.cell.reshape(9))
    out = np.full((5, 3 * s.n_atoms), np.nan)
    with B200Potential("eam", lib=backend, pin_host_buffers=1) as pot:
        for _ in range(2):  # second call: ranges already registered
            F1, U1, _ = pot.forceBatch(5, s.n_atoms, R, s.atomic_nrs, s.cell.reshape(9), out=out)
            assert F1.base is out or np.shares_memory(F1, out)
            assert same(F1, F0) and same(U1, U0)
        # per-system topologies and boxes still work next to the shared-row shortcut
        Zs = np.tile(s.atomic_nrs, (5, 1))
        Bs = np.tile(s.cell.reshape(9), (5, 1))
        F2, U2, _ = pot.forceBatch(5, s.n_atoms, R, Zs, Bs)
        assert same(F2, F0) and same(U2, U0)
    with pytest.raises(ValueError):
        with B200Potential("eam", lib=backend) as pot:
            pot.forceBatch(5, s.n_atoms, R, s.atomic_nrs, s.cell.reshape(9), out=np.zeros(7))


@pytest.mark.gpu
def test_measurement_aids(cuda_backend):
    """eonb200_l2_flush and eonb200_phase_clocks: per-CTA clocks are monotonic through the phases."""
    from eon_b200 import workloads as WL
    s = WL.load_structure("akmc_al.json")
    R = WL.jitter(s.positions, 8, 0.05, seed=5).reshape(8, -1)
    with B200Potential("eam", lib=cuda_backend) as pot:
        pot.forceBatch(8, s.n_atoms, R, s.atomic_nrs, s.cell.reshape(9))
        pot.phase_clocks(True)
        pot.l2_flush(64 << 20)
        pot.forceBatch(8, s.n_atoms, R, s.atomic_nrs, s.cell.reshape(9))
        c = pot.phase_clocks(False)
        assert c.shape[0] >= 8 and c.shape[1] == 16
        cyc = c[:, 2:11]
        assert (np.diff(cyc[:, [0, 1, 2, 4, 5, 6, 7]], axis=1) >= 0).all()
        assert (c[:, 1] >= 0).all() and (c[:, 1] < 1024).all()


def test_c3_sized_cell(backend, oracle):
    """BASELINE.json configs[2] geometry: fcc 8 x 8 x 8 with a vacancy (2047 atoms, a = 3.615 A, 54 neighbours inside
    the EAM cutoff), jittered walkers, against the oracle -- the shape where a walker spans several CTAs with one
    atom per thread of 512-thread CTAs.  Plus a size-independent property: with F = -grad U the net force vanishes."""
    from eon_b200 import workloads as WL
    s = WL.fcc_vacancy(8, 3.615, 13)
    R = WL.jitter(s.positions, 2, 0.05, seed=21)
    opot = oracle.potential(oracle.POT_EAM_NL)
    with B200Potential("eam", lib=backend) as pot:
        F, U, _ = pot.forceBatch(2, s.n_atoms, R.reshape(2, -1), s.atomic_nrs, s.cell.reshape(9))
        for w in range(2):
            U0, F0 = oracle.force(opot, R[w], s.atomic_nrs, s.cell)
            assert_parity(F[w], U[w], F0, U0)
        assert pot.nl_stats()["ordered_pairs"] > 2 * 2047 * 50
    with B200Potential("eam", lib=backend, conservative=True) as pot:
        F, U, _ = pot.forceBatch(2, s.n_atoms, R.reshape(2, -1), s.atomic_nrs, s.cell.reshape(9))
        assert np.abs(F.sum(axis=1)).max() < 1e-9 * np.abs(F).max() * s.n_atoms
