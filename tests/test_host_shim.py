"""The eonclient-side shim (eon_b200/host/B200Potential.hpp, an eonc::Potential subclass over the C ABI) driven
by a C++ harness the way eonclient drives a potential: forceBatch with arrays of per-system pointers
(client/ImprovedDimer.cpp:176, client/NudgedElasticBand.cpp:515) and force (client/Matter.cpp:523)."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HARNESS = os.path.join(HERE, "host", "harness")
LIB = os.path.join(ROOT, "eon_b200", "libeonb200.so")


def build_harness():
    out = subprocess.run(["make", "-C", os.path.join(HERE, "host"), "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout


def run(kind, systems, *extra):
    S, N = len(systems), len(systems[0][1])
    text = []
    for R, Z, box in systems:
        text.append(" ".join(repr(float(x)) for x in np.asarray(R).reshape(-1)))
        text.append(" ".join(str(int(z)) for z in Z))
        text.append(" ".join(repr(float(x)) for x in np.asarray(box).reshape(-1)))
    return subprocess.run([HARNESS, LIB, str(kind), str(S), str(N)] + list(extra), input="\n".join(text) + "\n", text=True,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)


def test_shim_compiles_and_fails_loudly_without_gpu():
    build_harness()
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("a GPU is present")
    out = run(3, [(np.zeros((1, 3)), [13], np.eye(3) * 10)])
    assert out.returncode == 3 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_shim_matches_oracle(oracle):
    build_harness()
    for kind, okind, fixture in ((3, "POT_EAM", "akmc_al.json"), (2, "POT_MORSE_PT", "akmc_pt.json"), (0, "POT_LJ", "lj13.json")):
        d = oracle.load_fixture(fixture)
        rng = np.random.default_rng(6)
        N = len(d["atomic_nrs"])
        systems = [(d["positions"] + rng.uniform(-0.05, 0.05, d["positions"].shape), d["atomic_nrs"], d["cell"]) for _ in range(3)]
        out = run(kind, systems)
        assert out.returncode == 0, out.stderr
        lines = out.stdout.split("\n")
        assert lines[0] == "calls 4"  # 3 by the callee in forceBatch (Potential.h:119-120) + 1 by the caller (Matter.cpp:525)
        vals = lines[1:]
        opot = oracle.potential(getattr(oracle, okind))
        p = 0
        for s in range(4):
            R, Z, box = systems[s % 3] if s < 3 else systems[0]
            E, var = map(float, vals[p].split())
            F = np.array([float(x) for x in vals[p + 1:p + 1 + 3 * N]]).reshape(N, 3)
            p += 1 + 3 * N
            U0, F0 = oracle.force(opot, R, Z, box)
            assert var == 0.0 and abs(E - U0) < 1e-8
            assert np.abs(F - F0).max() <= 1e-10 * np.abs(F0).max()


@pytest.mark.gpu
def test_shim_scattered_blocks_direct_path(oracle):
    """pin_host_buffers with one heap block per system (what Matter objects hand over): from the second call on the
    kernel reads and writes the caller's blocks directly (per-system addresses, page-locked span); forces are poisoned
    before every call, so whatever comes back was written by the last one."""
    build_harness()
    d = oracle.load_fixture("akmc_al.json")
    rng = np.random.default_rng(9)
    N = len(d["atomic_nrs"])
    systems = [(d["positions"] + rng.uniform(-0.05, 0.05, d["positions"].shape), d["atomic_nrs"], d["cell"]) for _ in range(5)]
    out = run(3, systems, "check", "-", "pin", "repeat=3")
    assert out.returncode == 0, out.stderr
    lines = out.stdout.split("\n")
    assert lines[0] == "calls 16"  # 3 x 5 in forceBatch + 1 by the caller
    vals = lines[1:]
    opot = oracle.potential(oracle.POT_EAM)
    p = 0
    for s in range(5):
        R, Z, box = systems[s]
        E, var = map(float, vals[p].split())
        F = np.array([float(x) for x in vals[p + 1:p + 1 + 3 * N]]).reshape(N, 3)
        p += 1 + 3 * N
        U0, F0 = oracle.force(opot, R, Z, box)
        assert var == 0.0 and abs(E - U0) < 1e-8
        assert np.abs(F - F0).max() <= 1e-10 * np.abs(F0).max()


@pytest.mark.gpu
def test_shim_rethrows_14324():
    build_harness()
    out = run(3, [(np.random.default_rng(0).uniform(0, 5, (4, 3)), [13, 13, 29, 13], np.eye(3) * 10)])
    assert out.returncode == 0 and out.stdout.strip() == "thrown 14324"


@pytest.mark.gpu
def test_shim_neb_band_forces(oracle):
    """`B200Potential::nebBandForces` (one call for the whole updateForces step) against the Python wrapper."""
    from eon_b200 import B200Potential, neb as NEB
    build_harness()
    a, b = oracle.load_fixture("lj13.json"), oracle.load_fixture("lj13_product.json")
    M, N = 5, len(a["atomic_nrs"])
    lam = np.linspace(0.0, 1.0, M + 2).reshape(-1, 1, 1)
    band = a["positions"][None] + lam * (b["positions"] - a["positions"])[None]
    out = run(0, [(band[i], a["atomic_nrs"], a["cell"]) for i in range(M + 2)], "neb")
    assert out.returncode == 0, out.stderr
    lines = out.stdout.split("\n")
    with B200Potential("lj") as pot:
        ref = NEB.band_forces_resident(pot, band.reshape(1, M + 2, -1), a["atomic_nrs"], a["cell"], climbing_image=1)
    assert lines[0] == "neb %d calls %d" % (ref["climbing"][0], M + 2)
    vals = np.array([float(x) for x in lines[1:] if x])
    assert np.array_equal(vals[:M + 2], ref["energies"][0])
    assert np.array_equal(vals[M + 2:], ref["projected"].reshape(-1))
