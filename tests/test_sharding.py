"""N > 1 path on CPU: walker sharding and the host-side harvest over a 2-rank gloo group
(walkers shard with no collective on the data path; the harvest is the only exchange)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

from eon_b200 import sharding
from eon_b200.walkers import Harvest


def test_shard_ranges_partition_the_walkers():
    for n, world, band in [(256, 1, 1), (256, 2, 1), (4096, 8, 1), (70, 4, 7), (9, 4, 1), (3, 8, 1)]:
        seen = []
        for r in range(world):
            lo, hi = sharding.shard_range(n, r, world, band)
            assert lo % band == 0 and hi % band == 0 and lo <= hi
            seen.extend(range(lo, hi))
        assert seen == list(range(n))
        for w in (0, n // 2, n - 1):
            lo, hi = sharding.shard_range(n, sharding.owner_of(w, n, world, band), world, band)
            assert lo <= w < hi
    with pytest.raises(ValueError):
        sharding.shard_range(10, 0, 2, band=3)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _fake_harvest(lo, hi, n_atoms=5):
    ids = np.arange(lo, hi)
    return Harvest(status=(ids % 3).astype(np.int32), iterations=ids.astype(np.int64) * 2,
                   force_calls=ids.astype(np.int64) * 7, energies=-1.0 * ids, eigenvalues=-0.5 * ids,
                   convergence=1e-3 * ids, positions=np.tile(ids[:, None].astype(float), (1, 3 * n_atoms)),
                   modes=None)


def _worker(rank, world, port, n_walkers, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(n_walkers, rank, world)
        got = sharding.harvest(_fake_harvest(lo, hi), dst=0)
        if rank == 0:
            ref = _fake_harvest(0, n_walkers)
            ok = (np.array_equal(got.status, ref.status) and np.array_equal(got.force_calls, ref.force_calls)
                  and np.array_equal(got.energies, ref.energies) and np.array_equal(got.positions, ref.positions)
                  and got.modes is None)
            q.put(("rank0", bool(ok), int(got.status.shape[0])))
        else:
            q.put(("other", got is None, 0))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_harvest_over_two_gloo_ranks():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_walkers = 37  # ragged split: 18 + 19
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_walkers, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out = dict((o[0], o[1:]) for o in out)
    assert out["rank0"] == (True, n_walkers)
    assert out["other"][0] is True


def test_harvest_without_process_group_is_identity():
    h = _fake_harvest(0, 4)
    assert sharding.harvest(h) is h


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under a 2-rank launch: rank 1 exits 0 without work or output."""
    import subprocess
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                          "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def _band_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    import emu_lib
    import oracle_lib
    from eon_b200 import B200Potential, neb as NEB
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = oracle_lib.load_fixture("lj13.json"), oracle_lib.load_fixture("lj13_product.json")
        B, M = 5, 3
        lam = np.linspace(0.0, 1.0, M + 2).reshape(1, -1, 1, 1)
        R = a["positions"][None, None] + lam * (b["positions"] - a["positions"])[None, None] \
            + np.random.default_rng(9).normal(0.0, 0.02, (B, M + 2, len(a["atomic_nrs"]), 3))
        R = R.reshape(B, M + 2, -1)
        lo, hi = sharding.shard_range(B, rank, world)
        with B200Potential("lj", lib=emu_lib.lib()) as pot:
            mine = NEB.band_forces_resident(pot, R[lo:hi], a["atomic_nrs"], a["cell"], climbing_image=1)
            got = sharding.gather_bands(mine, dst=0)
            if rank == 0:
                ref = NEB.band_forces_resident(pot, R, a["atomic_nrs"], a["cell"], climbing_image=1)
                ok = all(np.array_equal(got[k], ref[k]) for k in ("projected", "energies", "climbing")) \
                    and got["tangents"] is None and got["projected"].shape[0] == B
                q.put(("rank0", bool(ok), hi - lo))
            else:
                q.put(("other", got is None, hi - lo))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_bands_shard_over_two_gloo_ranks():
    """NEB bands shard like walkers: each rank evaluates its own bands (emulation backend here), one gather at the end."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out = dict((o[0], o[1:]) for o in out)
    assert out["rank0"][0] is True and out["other"][0] is True
    assert out["rank0"][1] + out["other"][1] == 5
