"""The CPU oracle pinned against the reference: golden / known-answer values of the reference's own tests,
the reference's EAM.cpp and vesin compiled from the reference tree (oracle/_ref).  CPU only."""
import numpy as np
import pytest


def canon(pairs, shifts):
    return set(map(tuple, np.concatenate([pairs, shifts.astype(np.int64)], axis=1).tolist()))


def test_point_goldens(oracle):
    g = oracle.goldens()
    d = oracle.load_fixture("lj13.json")
    U, F = oracle.force(oracle.potential(oracle.POT_LJ), d["positions"], d["atomic_nrs"], d["cell"])
    assert U == pytest.approx(g["lj_point"]["energy"], abs=5e-7)  # PotTest.cpp:38-43
    assert np.linalg.norm(F, axis=1).max() == pytest.approx(g["lj_point"]["max_force"], abs=5e-7)
    U, _ = oracle.force(oracle.potential(oracle.POT_LJCLUSTER), d["positions"], d["atomic_nrs"], d["cell"])
    assert U == pytest.approx(g["ljcluster_point"]["energy"], abs=5e-7)  # PotTest.cpp:130-132
    d = oracle.load_fixture("morse_pt_saddle.json")
    U, F = oracle.force(oracle.potential(oracle.POT_MORSE_PT), d["positions"], d["atomic_nrs"], d["cell"])
    assert U == pytest.approx(g["morse_pt_point"]["energy"], abs=5e-7)  # data/reference/point_morse_pt.dat
    free = d["fixed"].sum(axis=1) == 0
    assert np.linalg.norm(F[free], axis=1).max() == pytest.approx(g["morse_pt_point"]["max_force"], abs=5e-7)


def test_forces_are_minus_gradient(oracle):
    """Sign convention check: finite differences of the energy."""
    for kind, fixture, kw in ((oracle.POT_LJ, "lj13.json", {}), (oracle.POT_MORSE_PT, "akmc_pt.json", {}),
                              (oracle.POT_EAM, "al_fcc.json", {"conservative": True})):
        d = oracle.load_fixture(fixture)
        pot = oracle.potential(kind, **kw)
        rng = np.random.default_rng(0)
        R = d["positions"] + rng.uniform(-0.03, 0.03, d["positions"].shape)
        _, F = oracle.force(pot, R, d["atomic_nrs"], d["cell"])
        for (i, k) in ((0, 0), (len(R) - 1, 2)):
            h = 1e-5
            Rp, Rm = R.copy(), R.copy()
            Rp[i, k] += h
            Rm[i, k] -= h
            fd = -(oracle.force(pot, Rp, d["atomic_nrs"], d["cell"])[0] -
                   oracle.force(pot, Rm, d["atomic_nrs"], d["cell"])[0]) / (2 * h)
            assert F[i, k] == pytest.approx(fd, rel=1e-5, abs=1e-7)


def test_eam_restatement_is_bit_identical_to_reference_binary(oracle):
    if not oracle.ref_eam_available():
        pytest.skip("oracle/_ref/libeam_ref.so not built (reference checkout absent)")
    for fixture in ("al_fcc.json", "akmc_al.json"):
        d = oracle.load_fixture(fixture)
        rng = np.random.default_rng(1)
        for R in (d["positions"], d["positions"] + rng.uniform(-0.1, 0.1, d["positions"].shape), d["positions"] - 0.7):
            U0, F0 = oracle.ref_eam_force(R, d["atomic_nrs"], d["cell"])
            U, F = oracle.force(oracle.potential(oracle.POT_EAM), R, d["atomic_nrs"], d["cell"])
            assert U == U0 and np.array_equal(F, F0)
            U2, F2 = oracle.force(oracle.potential(oracle.POT_EAM_NL), R, d["atomic_nrs"], d["cell"])
            assert abs(U2 - U0) < 1e-9 and np.abs(F2 - F0).max() <= 1e-11 * np.abs(F0).max()


def test_neighbour_predicate_matches_vesin(oracle):
    if not oracle.ref_vesin_available():
        pytest.skip("oracle/_ref/libvesin_ref.so not built (reference checkout absent)")
    rng = np.random.default_rng(3)
    cases = []
    for fixture, rc in (("lj13.json", 15.0), ("morse_pt_saddle.json", 9.5), ("akmc_al.json", 5.555)):
        d = oracle.load_fixture(fixture)
        cases.append((d["positions"], d["cell"], rc, (1, 1, 1)))
    H = np.array([[6.0, 0, 0], [1.5, 5.0, 0], [0.7, -1.2, 7.0]])
    cases.append((rng.uniform(-1, 2, (30, 3)) @ H, H, 4.5, (1, 1, 1)))
    cases.append((rng.uniform(0, 1, (30, 3)) @ H, H, 4.5, (1, 0, 1)))
    cases.append((rng.uniform(0, 1, (6, 3)) @ np.diag([3.0, 3.5, 4.0]), np.diag([3.0, 3.5, 4.0]), 7.0, (1, 1, 1)))
    for R, box, rc, per in cases:
        for full in (True, False):
            for algo in (0, 2):  # vesin's automatic choice and its cell list give the same set
                Pv, Sv = oracle.ref_vesin_neighbors(R, box, rc, full=full, periodic=per, algorithm=algo)
                Po, So = oracle.neighbors(R, box, rc, full=full, periodic=per)
                assert len(Pv) == len(Po) and canon(Pv, Sv) == canon(Po, So)


def test_saddle_search_known_answers(oracle):
    """client/unit_tests/data/reference/{saddle_search_morse_pt,process_search_morse_pt}.dat."""
    g = oracle.goldens()
    d = oracle.load_fixture("morse_pt_saddle.json")
    pot = oracle.potential(oracle.POT_MORSE_PT)
    Z, box, fixed = d["atomic_nrs"], d["cell"], d["fixed"]
    E0, _ = oracle.force(pot, d["positions"], Z, box)
    res, _, _ = oracle.saddle_search(pot, oracle.search_params(), d["displacement_positions"], Z, fixed, box,
                                     d["direction"], E0)
    k = g["saddle_dimer_cg"]
    assert res.status == 0 and res.iterations == k["iterations"]
    assert res.force_calls == k["force_calls_saddle"]
    assert res.energy == pytest.approx(k["energy_saddle"], abs=5e-7)
    assert res.eigenvalue == pytest.approx(k["eigenvalue"], abs=5e-7)
    res, _, _ = oracle.saddle_search(pot, oracle.search_params(opt_method=oracle.OPT_LBFGS), d["displacement_positions"],
                                     Z, fixed, box, d["direction"], E0)
    k = g["process_search_lbfgs"]
    assert res.status == 0 and res.force_calls <= k["force_calls_saddle"]
    assert res.energy == pytest.approx(k["energy_saddle"], abs=2e-5)


def test_lanczos_known_answer(oracle):
    """saddle_search_lanczos_morse_pt.dat: E_saddle, status, <= 44 force calls (JobIntegrationTest.cpp:418-460)."""
    g = oracle.goldens()["saddle_lanczos_cg"]
    d = oracle.load_fixture("morse_pt_saddle.json")
    pot = oracle.potential(oracle.POT_MORSE_PT)
    E0, _ = oracle.force(pot, d["positions"], d["atomic_nrs"], d["cell"])
    res, _, _ = oracle.saddle_search(pot, oracle.search_params(minmode_method=oracle.MINMODE_LANCZOS),
                                     d["displacement_positions"], d["atomic_nrs"], d["fixed"], d["cell"],
                                     d["direction"], E0)
    assert res.status == 0 and res.force_calls + 1 <= g["total_force_calls"]
    assert res.energy == pytest.approx(g["energy_saddle"], abs=2e-5)
    assert res.eigenvalue < 0


def test_lj_minimisation_known_answers(oracle):
    """minimization_lj.dat (LBFGS), min_lj_cg.dat, min_lj_fire.dat: final energy and force-call counts."""
    g = oracle.goldens()
    d = oracle.load_fixture("lj13.json")
    pot = oracle.potential(oracle.POT_LJ)
    for key, method in (("lj_minimization", oracle.OPT_LBFGS), ("lj_min_cg", oracle.OPT_CG),
                        ("lj_min_fire", oracle.OPT_FIRE)):
        res, _ = oracle.relax(pot, oracle.search_params(opt_method=method, opt_converged_force=1e-3), d["positions"],
                              d["atomic_nrs"], d["fixed"], d["cell"])
        assert res.converged == 1
        assert res.energy == pytest.approx(g[key]["energy"], abs=2e-6)
        assert res.force_calls <= g[key]["force_calls"]
