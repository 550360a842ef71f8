#!/usr/bin/env python
"""bench.py -- atom-force-evals/s (and saddle searches/hour) of the batched EAM hot path.

    python bench.py --gpus N --steps K --warmup W            # this repository's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU EAM.cpp

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): examples/akmc-al -- 385 Al atoms
(128 fixed), box 22.845 x 22.845 x 64 A, in-tree EAM Al row (rc 5.555 A), 256 concurrent walkers
per GPU, positions jittered U(-0.05, 0.05) A per walker.  One "step" = one batched energy+force
evaluation of all walkers of a rank (256 x 385 atom-force-evals).  Walkers are independent, so
ranks shard them with no collective on the data path (weak scaling: 256 walkers per GPU).

Timed region of `value`: inputs resident in HBM, L2 flushed before every step (a 512 MiB memset queued on
the same stream ahead of the step, outside the CUDA events), CUDA events on the launching stream around each step's kernels, summed
over the K steps, max over ranks.  `e2e`: the same metric through B200Potential.forceBatch with
host buffers (H2D + kernels + D2H inside the timed region).  The saddle-search leg runs the whole
device-resident dimer/LBFGS search of 256 walkers once and reports searches/hour.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from eon_b200 import workloads as WL  # noqa: E402

N_WALKERS = 256
JITTER = 0.05
N_POSITION_SETS = 4
FLUSH_BYTES = 512 << 20  # > 4x the 126 MB L2
# dram__bytes_read.sum + dram__bytes_write.sum of one k_eam_fused launch of this workload (ncu --set full capture)
NCU_DRAM_BYTES_PER_LAUNCH = 66.6e6
NCU_TRAFFIC_SOURCE = ("profiles/r2z_n3_ncu_raw_selected.csv (k_eam_n3, one launch of 256 x 385 atoms under ncu --set full, "
                      "cold L2): 56.4 MB read + 10.1 MB written against 33.7 MB algorithmic (64 N + 8 P_o, SURVEY.md 8d). "
                      "The surplus: list rows are fetched whole up to the longest row of the walker (lanes with shorter "
                      "rows never read the tail), and part of the 39 MB pair scratch (g'/r and the Morse term, written "
                      "by the density pass and read back by the force pass) leaves L2 under the profiler's cache "
                      "control. 66.6 MB is 10 us at the HBM peak, a sixth of the kernel's duration: not the bound")
SEARCH_ROW_Z = WL.SEARCH_ROW_Z  # the in-tree Au row: see DESIGN.md "Search workload"
A0_AU = WL.A0_AU  # equilibrium fcc lattice constant of the in-tree EAM Au row (oracle)


def eam_flops(ordered_pairs: float, atoms: float) -> float:
    """flops_EAM = 107 * P_o + 40 * N (exp = 20 flop convention)."""
    return 107.0 * ordered_pairs + 40.0 * atoms


def eam_min_bytes(ordered_pairs: float, atoms: float) -> float:
    """read R 24 + write F 24 + 2 passes x 4 B per neighbour index + 16 B rho/F' per atom."""
    return 64.0 * atoms + 8.0 * ordered_pairs


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            tok = [t.strip() for t in r.split(",")]
            if len(tok) < 7:
                continue
            try:
                sm.append(float(tok[0]))
                smax.append(float(tok[1]))
            except ValueError:
                continue
            for name, val in zip(names, tok[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(n_atoms: int, world: int, skin: float, rebuilt: int) -> dict:
    """The `config` both arms print (the reference arm times the same walkers on the host cores)."""
    return {"workload": "examples/akmc-al: 385 Al atoms (128 fixed), in-tree EAM Al row, "
                        f"{N_WALKERS} walkers per GPU, jitter U(-{JITTER},{JITTER}) A, {N_POSITION_SETS} position sets",
            "walkers_per_gpu": N_WALKERS, "n_atoms": n_atoms,
            "l2": "flushed before every step (512 MiB memset queued on the same stream, outside the events)",
            "neighbour_list": f"Verlet skin {skin} A, cached; lists rebuilt inside the timed region: {rebuilt}",
            "parallelism": f"walker-sharded x{world}, no collective"}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def bind_to_gpu_numa_node(local_rank: int, world: int) -> dict:
    """Pin this rank to its own slice of the cores of the NUMA node its GPU hangs off (one host thread per GPU; eight
    Python processes wandering over two sockets cost the round-1 e2e number a quarter of its scaling)."""
    info = {"bound": False}
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = -1
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        allowed = sorted(os.sched_getaffinity(0))
        cpus = allowed
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
                txt = fh.read().strip()
            nodecpus = set()
            for part in txt.split(","):
                a, _, b = part.partition("-")
                nodecpus.update(range(int(a), int(b or a) + 1))
            cpus = [c for c in allowed if c in nodecpus] or allowed
        # ranks whose GPUs share a node split its cores evenly
        per = max(1, len(cpus) // max(1, world))
        mine = cpus[(local_rank % max(1, world)) * per:(local_rank % max(1, world) + 1) * per] or cpus
        os.sched_setaffinity(0, mine)
        info = {"bound": True, "numa_node": node, "cpus": len(mine)}
    except Exception as exc:  # no sysfs / no permission: run unbound
        info["why"] = str(exc)[:80]
    return info


# ------------------------------------------------------------------------------------------------------
# inputs
def force_workload(first_walker: int):
    s = WL.load_structure("akmc_al.json")
    sets = [WL.jitter(s.positions, N_WALKERS, JITTER, seed=1000 * k + first_walker, fixed=None)
            for k in range(N_POSITION_SETS)]
    return s, np.stack(sets)  # [SETS][W][N][3]


def search_workload():
    """akmc-al geometry scaled to the lattice constant of the in-tree EAM Au row; all atoms Z = 79."""
    return WL.akmc_al_on_au_row()


# examples/akmc-al/config.ini: LBFGS to 1e-4, Lanczos with tolerance 0.05, max_energy 10 (workloads.AKMC_AL); the dimer
# variant of the same searches is reported next to it
SEARCH_PARAMS = dict(WL.AKMC_AL["search"])
SEARCH_PARAMS_DIMER = dict(SEARCH_PARAMS, minmode_method=0)
STREAM_SLOTS, STREAM_QUEUE, STREAM_GROUPS = 512, 2048, 32


# ------------------------------------------------------------------------------------------------------
# CPU arms (the only place outside tests/ and smoke() that touches oracle/)
def _cpu_worker(args):
    """One pinned process: evaluate walkers in a loop for `budget_s` seconds, return the number of evaluations."""
    kind, core, positions, Z, cell, budget_s = args
    try:
        os.sched_setaffinity(0, {core})
    except OSError:
        pass
    import oracle_lib as O
    if kind == "reference":
        def one(w):
            O.ref_eam_force(positions[w % len(positions)], Z, cell)
    else:
        pot = O.potential(O.POT_EAM if kind == "port" else O.POT_EAM_NL)

        def one(w):
            O.force(pot, positions[w % len(positions)], Z, cell)
    one(0)
    n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        one(n)
        n += 1
    return n, time.perf_counter() - t0


CPU_KINDS = {"reference": "EAM.cpp compiled from the reference tree (builds cell lists it never reads, then an O(N^2) loop)",
             "port": "oracle restatement of EAM.cpp (bit-identical results, O(N^2) loop only)",
             "port-nl": "oracle with a cell-list neighbour search (what a non-wasteful CPU loop does)"}


def cpu_force_baseline(struct, positions, budget_s: float, use_reference: bool, kind: str = None):
    """Time a CPU EAM on a bounded sample of the same walkers: ONE PINNED PROCESS PER HOST CORE (BASELINE.md section
    3), each evaluating single walkers for `budget_s` seconds."""
    import multiprocessing as mp
    import oracle_lib as O
    cores_avail = sorted(os.sched_getaffinity(0))
    cores = len(cores_avail)
    if kind is None:
        kind = "reference" if (use_reference and O.ref_eam_available()) else "port"
    if kind == "reference" and not O.ref_eam_available():
        kind = "port"
    O.lib()  # build / load before forking
    jobs = [(kind, c, positions[:8], struct.atomic_nrs, struct.cell, budget_s) for c in cores_avail]
    ctx = mp.get_context("spawn")  # (the parent may hold a CUDA context)
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        out = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    n_evals = sum(o[0] for o in out)
    dt = max(o[1] for o in out)
    rate = n_evals * struct.n_atoms / dt
    return {"value": rate, "unit": "atom-force-evals/s", "cores": cores, "kind": "reference" if kind == "reference" else "port",
            "variant": kind,
            "sample": f"{n_evals} single-walker evaluations of the {struct.n_atoms}-atom akmc-al cell ({CPU_KINDS[kind]}), "
                      f"{cores} pinned processes, {dt:.1f} s each (wall {wall:.1f} s with start-up)"}, dt, n_evals


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    struct, sets = force_workload(0)
    positions = sets[0]
    per_step_budget = min(4.0, 100.0 / max(1, args.steps))  # the whole K-step run stays within a couple of minutes
    times = []
    evals = 0
    info = None
    for step in range(args.warmup + args.steps):
        info, dt, n = cpu_force_baseline(struct, positions, per_step_budget if step >= args.warmup else 0.5, True)
        if step >= args.warmup:
            times.append(dt)
            evals += n
    total = sum(times)
    value = evals * struct.n_atoms / total
    line = {"impl": "reference", "metric": "atom-force-evals/sec", "value": value, "unit": "atom-force-evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(struct.n_atoms, max(1, world), args.skin, 0),
                           reference_arm="EAM.cpp of the reference tree on the host cores, bounded sample of the same "
                                         "walkers per step (no device, no lists)"),
            "cpu_baseline": dict(info, value=value),
            "e2e": {"value": value, "unit": "atom-force-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def cpu_search_baseline(struct, Rmin, E0, n_walkers: int):
    import oracle_lib as O
    cores = os.cpu_count() or 1
    pot = O.potential(O.POT_EAM_NL, conservative=True)
    prm = O.search_params(**SEARCH_PARAMS)
    cfg = dict(WL.AKMC_AL)
    R, M = [], []
    for k in range(n_walkers):
        a, b = WL.displace(struct, Rmin, struct.n_atoms - 1, cfg["radius"] * A0_AU / 4.05, cfg["magnitude"], cfg["seed"] + k)
        R.append(a)
        M.append(b)

    def one(k):
        res, _, _ = O.saddle_search(pot, prm, R[k], struct.atomic_nrs, struct.fixed, struct.cell, M[k], E0)
        return res.status, res.force_calls, res.energy

    with ThreadPoolExecutor(cores) as ex:
        t0 = time.perf_counter()
        out = list(ex.map(one, range(n_walkers)))
        dt = time.perf_counter() - t0
    good = sum(1 for o in out if o[0] == 0)
    return {"searches_per_hour": good / dt * 3600.0, "completed": n_walkers, "status_good": good, "cores": cores,
            "kind": "port (oracle dimer/LBFGS over the cell-list EAM variant)", "seconds": dt,
            "force_calls": int(sum(o[1] for o in out))}, out


# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    from eon_b200 import B200Potential, Walkers, search_params

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(local_rank, world)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        return float(t.item())

    from eon_b200 import sharding
    first, last = sharding.shard_range(N_WALKERS * world, rank, world)  # weak scaling: 256 walkers per GPU
    assert last - first == N_WALKERS
    struct, sets = force_workload(first)
    N, W = struct.n_atoms, N_WALKERS
    pot = B200Potential("eam", device=local_rank, skin=args.skin, tiles_override=args.tiles, lpa_override=args.lpa)
    d_R = torch.from_numpy(sets.reshape(N_POSITION_SETS, W, 3 * N)).to(dev)
    d_Z = torch.from_numpy(np.broadcast_to(struct.atomic_nrs, (W, N)).copy()).to(dev)
    d_box = torch.from_numpy(np.broadcast_to(struct.cell.reshape(9), (W, 9)).copy()).to(dev)
    d_F = torch.empty((W, 3 * N), dtype=torch.float64, device=dev)
    d_U = torch.empty((W,), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step(k):
        pot.forceBatchDevice(W, N, d_R[k % N_POSITION_SETS].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(),
                             d_box.data_ptr(), sync=True)
        nl_ms, f_ms = pot.last_force_ms()
        return nl_ms + f_ms

    for k in range(max(3, args.warmup)):
        step(k)
    stats = pot.nl_stats()
    pairs_per_step = float(stats["ordered_pairs"])
    # FP64 roofline denominator, measured live (MEASURED_PEAKS.json has no FP64 figure)
    fp64_peak = pot.measure_fp64_peak(5)

    # Ranks reach the barrier below at different times, and a GPU that waits idle for the others drops its clocks: the
    # whole timed region (a few milliseconds) then runs at half speed on that rank (seen once at 8 GPUs: 0.135 ms per
    # step on one rank, 0.062 on the rest).  So: meet once, keep every GPU busy for 0.15 s with untimed steps, meet again.
    barrier()
    t_spin = time.perf_counter() + 0.15
    k_spin = 0
    while time.perf_counter() < t_spin:
        step(k_spin)
        k_spin += 1
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = pot.kernelLaunches
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    rebuilt = 0
    for k in range(args.steps):
        pot.l2_flush(FLUSH_BYTES)  # queued on the potential's stream ahead of the step: L2 cold, stream never dry
        dev_ms += step(k)
        rebuilt += pot.nl_stats()["rebuilt"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = pot.kernelLaunches - launches0
    clocks = sampler.stop()
    dev_ms_max = reduce_max(dev_ms)
    if os.environ.get("EONB200_BENCH_TRACE"):  # development aid: every rank's own device time
        print("[bench] rank %d: %.4f ms per step (device), nl %s" % (rank, dev_ms / args.steps, pot.nl_stats()), file=sys.stderr, flush=True)
    units = reduce_sum(float(W) * N * args.steps)
    value = units / (dev_ms_max * 1e-3)

    # hot-L2 variant (no flush) and a fresh-list variant (skin 0: cell-list rebuild on every step)
    t_hot = 0.0
    for k in range(args.steps):
        t_hot += step(k)
    pot0 = B200Potential("eam", device=local_rank, skin=0.0, tiles_override=args.tiles, lpa_override=args.lpa)
    for k in range(3):
        pot0.forceBatchDevice(W, N, d_R[k % N_POSITION_SETS].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(),
                              d_box.data_ptr(), sync=True)
    t_fresh = 0.0
    for k in range(args.steps):
        pot0.l2_flush(FLUSH_BYTES)
        pot0.forceBatchDevice(W, N, d_R[k % N_POSITION_SETS].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(),
                              d_box.data_ptr(), sync=True)
        a, b = pot0.last_force_ms()
        t_fresh += a + b
    pot0.close()

    # e2e: host buffers through the plugin call (H2D + kernels + D2H inside the timed region)
    # The caller owns its buffers, as eOn's Matter does: position sets and one force array live for the whole run, and
    # the handle is asked to page-lock them (pin_host_buffers) instead of staging every call.
    epot = B200Potential("eam", device=local_rank, skin=args.skin, tiles_override=args.tiles, lpa_override=args.lpa,
                         pin_host_buffers=1)
    h_R = [np.ascontiguousarray(sets[k].reshape(W, 3 * N)) for k in range(N_POSITION_SETS)]
    h_F = np.empty((W, 3 * N), dtype=np.float64)
    for k in range(N_POSITION_SETS + 1):
        epot.forceBatch(W, N, h_R[k % N_POSITION_SETS], struct.atomic_nrs, struct.cell.reshape(9), out=h_F)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        _, e_U, _ = epot.forceBatch(W, N, h_R[k % N_POSITION_SETS], struct.atomic_nrs, struct.cell.reshape(9), out=h_F)
    torch.cuda.synchronize()
    e2e_s = reduce_max(time.perf_counter() - t0)
    # same position set as the last device-resident step above: the host path must return the same energies
    assert np.allclose(e_U, d_U.cpu().numpy(), rtol=0.0, atol=1e-8), "e2e energies differ from the device-resident run"
    # the same call sustained for 400 more steps: with the GPU idle two thirds of every call its clocks settle lower than
    # right after the device-resident leg, and the call time with them (reported beside the K timed steps)
    t0 = time.perf_counter()
    for k in range(400):
        epot.forceBatch(W, N, h_R[k % N_POSITION_SETS], struct.atomic_nrs, struct.cell.reshape(9), out=h_F)
    torch.cuda.synchronize()
    e2e_sustained_ms = 1e3 * reduce_max(time.perf_counter() - t0) / 400
    epot.close()
    e2e_value = units / e2e_s
    # The same call through the C++ shim an eonclient build would link (eon_b200/host/B200Potential.hpp), every system
    # in its own heap block like the Matter objects of a dimer or a band (client/Matter.cpp:523): staged copies.
    e2e_shim = None
    exe = os.path.join(ROOT, "tests", "host", "harness")
    if os.path.exists(exe):
        text = []
        for w in range(W):
            text.append(" ".join(repr(float(x)) for x in sets[0][w].reshape(-1)))
            text.append(" ".join(str(int(z)) for z in struct.atomic_nrs))
            text.append(" ".join(repr(float(x)) for x in struct.cell.reshape(-1)))
        barrier()
        try:
            out = subprocess.run([exe, os.path.join(ROOT, "eon_b200", "libeonb200.so"), "3", str(W), str(N), "bench",
                                  str(max(10, args.steps)), "pin", f"device={local_rank}"], input="\n".join(text) + "\n",
                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
            tok = out.stdout.split()
            ms_mean = reduce_max(float(tok[1])) if out.returncode == 0 and len(tok) >= 3 else float("nan")
            e2e_shim = {"ms_per_step": ms_mean, "value": W * N * world / (ms_mean * 1e-3), "unit": "atom-force-evals/s",
                        "how": "tests/host/harness bench: B200Potential::forceBatch with one heap block per system "
                               "(what Matter objects hand over): the span of the heap that holds the blocks is "
                               "page-locked once, the kernel gets one device-visible address per system and reads "
                               "its coordinates / writes its forces there directly (bulk copies over PCIe, no "
                               "staging memcpy, one launch); wall clock around the call"}
        except Exception as exc:
            e2e_shim = {"error": str(exc)[:120]}
    # bytes the call moves: positions in (one shared topology row and box), forces + energies + list flags out
    h2d = W * 3 * N * 8 + N * 4 + 72
    d2h = W * 3 * N * 8 + W * 8 + 36

    # roofline of the dominant kernel (k_force_batch): algorithmic flops / measured kernel time
    ms_kernel = dev_ms / args.steps
    flops = eam_flops(pairs_per_step, W * N)
    achieved = flops / (ms_kernel * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    hbm_achieved = eam_min_bytes(pairs_per_step, W * N) / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "fp64", "kernel": "k_eam_n3 (one launch per step: staging, density over the owned half of the pairs, embedding, owner and receiver force passes)", "achieved": achieved, "peak": fp64_peak / 1e12,
                "unit": "TFLOP/s", "frac": achieved / (fp64_peak / 1e12) if fp64_peak > 0 else None, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                "traffic_source": NCU_TRAFFIC_SOURCE,
                "peak_source": "DFMA microbenchmark k_dfma_peak measured in this run (no FP64 figure in MEASURED_PEAKS.json)",
                "algorithmic_flops_per_launch": flops, "ordered_pairs_per_launch": pairs_per_step,
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "peak_source": hbm_src}}

    # device-resident saddle searches (secondary metric): examples/akmc-al settings on the stable Au-row geometry
    search = None
    if not args.no_search:
        s2, scale = search_workload()
        spot = B200Potential("eam", device=local_rank, skin=args.skin, conservative=True)
        wk = Walkers(spot, s2.atomic_nrs, s2.fixed, s2.cell, np.broadcast_to(s2.positions.reshape(1, -1), (W, 3 * N)).copy())
        wk.relax(search_params(opt_method=1, opt_converged_force=1e-4))
        hmin = wk.harvest()
        Rmin, E0 = hmin.positions[0], float(hmin.energies[0])
        cfg = dict(WL.AKMC_AL)
        Q = STREAM_QUEUE
        # start geometries and modes of the whole queue on the device (k_displace: Philox per (walker, atom); the
        # reference's server draws them with numpy, eon/displace.py:180-263), read back once for the streamed leg
        gen = Walkers(spot, s2.atomic_nrs, s2.fixed, s2.cell, np.broadcast_to(Rmin.reshape(1, -1), (Q, 3 * N)).copy())
        gen.displace(Rmin, N - 1, cfg["radius"] * scale, cfg["magnitude"], cfg["seed"], first_walker_id=Q * rank)
        hq = gen.harvest()
        Rq, Mq = hq.positions, hq.modes
        gen.close()

        def one_launch(params):
            wk.set(positions=Rq[:W], modes=Mq[:W])
            prm = search_params(**params)
            l0 = spot.kernelLaunches
            barrier()
            wk.saddle_search(prm, E0)
            ms = reduce_max(wk.last_ms)
            st = spot.nl_stats()
            h_local = wk.harvest(geometry=False)
            good = reduce_sum(float((h_local.status == 0).sum()))
            fcalls = reduce_sum(float(h_local.force_calls.sum()))
            rebuilt, reused = reduce_sum(float(st["rebuilt"])), reduce_sum(float(st["reused"]))
            h = sharding.harvest(h_local, dst=0)  # the path's only exchange: results of all ranks, walker order, on rank 0
            if h is None:
                h = h_local
            elif world > 1:
                assert h.status.shape[0] == W * world and int((h.status == 0).sum()) == int(good)
            return {"searches_per_hour": good / (ms * 1e-3) * 3600.0, "unit": "status-0 saddle searches/hour",
                    "walkers": int(W * world), "status_good": int(good), "ms": ms,
                    "status_histogram": {int(k): int(v) for k, v in zip(*np.unique(h.status, return_counts=True))},
                    "force_calls": int(fcalls), "atom_force_evals_per_s": fcalls * N / (ms * 1e-3),
                    "list_rebuild_fraction": rebuilt / max(1.0, rebuilt + reused),
                    "gpu_launches": int(spot.kernelLaunches - l0)}

        search = one_launch(SEARCH_PARAMS)
        search["config"] = ("akmc-al geometry scaled to a0 of the in-tree EAM Au row; examples/akmc-al/config.ini settings: "
                            "Lanczos (tolerance 0.05) + LBFGS 1e-4, displace radius 5.0 / magnitude 0.2 around the adatom; "
                            "one launch, one CTA per walker: lasts as long as its slowest walker")
        search["dimer"] = one_launch(SEARCH_PARAMS_DIMER)
        if rank == 0 and not args.no_cpu:
            cpu_s, _ = cpu_search_baseline(s2, Rmin, E0, max(8, os.cpu_count() or 1))
            search["cpu_baseline"] = cpu_s
        wk.close()
        # Throughput with every slot kept busy: a queue of searches through STREAM_SLOTS slots in STREAM_GROUPS groups,
        # each group refilled from the queue as soon as its walkers are done (eon/communicator.py:525-559 does the same
        # with cores); wall clock from the first submit to the last harvest, uploads and harvests included.
        wks = Walkers(spot, s2.atomic_nrs, s2.fixed, s2.cell, np.broadcast_to(Rmin.reshape(1, -1), (STREAM_SLOTS, 3 * N)).copy())
        prm = search_params(**SEARCH_PARAMS)
        wks.stream_search(prm, Rq[:STREAM_SLOTS], Mq[:STREAM_SLOTS], E0, groups=STREAM_GROUPS, geometry=False)  # warm-up
        barrier()
        t0 = time.perf_counter()
        hs = wks.stream_search(prm, Rq, Mq, E0, groups=STREAM_GROUPS, geometry=False)
        wall = reduce_max(time.perf_counter() - t0)
        goods = reduce_sum(float((hs.status == 0).sum()))
        fcs = reduce_sum(float(hs.force_calls.sum()))
        streamed = {"queue": int(Q * world), "slots_per_gpu": STREAM_SLOTS, "groups": STREAM_GROUPS,
                    "wall_s": wall, "status_good": int(goods),
                    "searches_per_hour": goods / wall * 3600.0,
                    "atom_force_evals_per_s": fcs * N / wall,
                    "how": "Walkers.stream_search: eonb200_walkers_submit / poll / harvest_group on one stream "
                           "per group; a walker at max_iterations only holds up its own 16-slot group"}
        # the same queue through ONE launch: every CTA takes the next entry from a device-side counter when its walker is done
        wks.queue_search(prm, Rq[:STREAM_SLOTS], Mq[:STREAM_SLOTS], E0, geometry=False)  # warm-up
        barrier()
        t0 = time.perf_counter()
        hq2 = wks.queue_search(prm, Rq, Mq, E0, geometry=False)
        wall_q = reduce_max(time.perf_counter() - t0)
        good_q = reduce_sum(float((hq2.status == 0).sum()))
        fc_q = reduce_sum(float(hq2.force_calls.sum()))
        differ_q = int(reduce_sum(float((hq2.status != hs.status).sum())))  # (0 expected: an entry's result does not depend on its slot)
        queued = {"queue": int(Q * world), "slots_per_gpu": STREAM_SLOTS, "wall_s": wall_q, "status_good": int(good_q),
                  "searches_per_hour": good_q / wall_q * 3600.0, "atom_force_evals_per_s": fc_q * N / wall_q,
                  "gpu_launches": 1, "outcomes_differing_from_streamed": differ_q,
                  "how": "Walkers.queue_search: eonb200_walkers_queue, one launch, resident CTAs pull the next search "
                         "from a device counter (wall clock includes the upload of the start geometries and the "
                         "read-back of every result)"}
        # The headline of the search leg is the queue (every slot refilled as soon as its walker or group finishes -- how a
        # production run keeps the GPU busy, and what scales with the number of GPUs); the one-launch runs, which last
        # as long as their slowest walker, stay beside it.
        one = search
        best = queued if queued["searches_per_hour"] >= streamed["searches_per_hour"] else streamed
        search = {"searches_per_hour": best["searches_per_hour"], "unit": "status-0 saddle searches/hour",
                  "mode": "queue of %d searches per GPU through %d resident slots (%s)"
                          % (Q, STREAM_SLOTS, "device-side queue, one launch" if best is queued else "streamed groups"),
                  "wall_s": best["wall_s"], "status_good": best["status_good"], "config": one.pop("config"),
                  "queued": queued, "streamed": streamed, "dimer": one.pop("dimer"),
                  "cpu_baseline": one.pop("cpu_baseline", None), "one_launch": one}
        wks.close()
        spot.close()

    # BASELINE.json configs[2] as a second data point: 4096 walkers of the 2047-atom vacancy cell sharded over the
    # ranks (the reference has no Cu EAM row -- SURVEY.md section 7 -- so the Al row runs on the Cu lattice, n = 54)
    c3 = None
    if not args.no_c3:
        s3 = WL.fcc_vacancy(8, 3.615, 13)
        N3 = s3.n_atoms
        lo3, hi3 = sharding.shard_range(4096, rank, world)
        W3 = hi3 - lo3
        sets3 = [torch.from_numpy(WL.jitter(s3.positions, W3, JITTER, seed=7000 * k + lo3).reshape(W3, 3 * N3)).to(dev)
                 for k in range(2)]
        d_Z3 = torch.from_numpy(np.broadcast_to(s3.atomic_nrs, (W3, N3)).copy()).to(dev)
        d_box3 = torch.from_numpy(np.broadcast_to(s3.cell.reshape(9), (W3, 9)).copy()).to(dev)
        d_F3 = torch.empty((W3, 3 * N3), dtype=torch.float64, device=dev)
        d_U3 = torch.empty((W3,), dtype=torch.float64, device=dev)
        pot3 = B200Potential("eam", device=local_rank, skin=args.skin)

        def step3(k):
            pot3.forceBatchDevice(W3, N3, sets3[k % 2].data_ptr(), d_Z3.data_ptr(), d_F3.data_ptr(), d_U3.data_ptr(),
                                  d_box3.data_ptr(), sync=True)
            return sum(pot3.last_force_ms())

        for k in range(3):
            step3(k)
        pairs3 = float(pot3.nl_stats()["ordered_pairs"])
        barrier()
        ms3 = 0.0
        n3 = 5
        for k in range(n3):
            pot3.l2_flush(FLUSH_BYTES)
            ms3 += step3(k)
        ms3 = reduce_max(ms3) / n3
        flops3 = reduce_sum(eam_flops(pairs3, W3 * N3))
        c3 = {"workload": "fcc 8x8x8 vacancy cell, 2047 atoms, a = 3.615 A, in-tree EAM Al row, 4096 walkers over all ranks",
              "walkers": 4096, "n_atoms": N3, "ms_per_step": ms3, "atom_force_evals_per_s": 4096.0 * N3 / (ms3 * 1e-3),
              "tflops": flops3 / (ms3 * 1e-3) / 1e12,
              "fp64_roofline_frac": flops3 / (ms3 * 1e-3) / (fp64_peak * world) if fp64_peak > 0 else None}
        pot3.close()
        del sets3, d_Z3, d_box3, d_F3, d_U3
        # ... and as resident saddle searches (BASELINE: "4096 walkers sharded over 1/2/4/8 B200"): settings of
        # examples/akmc-cu-vacancy/config.ini (dimer, rotations_max 20, LBFGS memory 25, converged_force 0.01,
        # max_move 0.05) on the in-tree Au row at its own lattice constant (the reference has no Cu EAM row and the Al
        # row has no stable crystal), every walker displaced around a neighbour of the vacancy
        if not args.no_search:
            s3s = WL.fcc_vacancy(8, WL.A0_AU, WL.SEARCH_ROW_Z)
            spot3 = B200Potential("eam", device=local_rank, skin=args.skin, conservative=True)
            wk3 = Walkers(spot3, s3s.atomic_nrs, s3s.fixed, s3s.cell, np.broadcast_to(s3s.positions.reshape(1, -1), (W3, 3 * N3)).copy())
            wk3.set_active([1] + [0] * (W3 - 1))
            wk3.relax(search_params(opt_method=1, opt_converged_force=1e-3, lbfgs_memory=25))
            wk3.set_active(None)
            h3 = wk3.harvest()
            Rmin3, E03 = h3.positions[0], float(h3.energies[0])
            R3 = np.empty((W3, N3, 3))
            M3 = np.empty((W3, N3, 3))
            for k in range(W3):
                R3[k], M3[k] = WL.displace(s3s, Rmin3, 0, 3.0 * WL.A0_AU / 3.615, 0.3, 42 + lo3 + k)
            wk3.set(positions=R3, modes=M3)
            prm3 = search_params(opt_method=1, opt_converged_force=0.01, lbfgs_memory=25, max_move=0.05, minmode_method=0,
                                 dimer_rotations_max=20, dimer_converged_angle=1.0, saddle_max_energy=300.0,
                                 saddle_max_iterations=1000)
            barrier()
            wk3.saddle_search(prm3, E03)
            ms3s = reduce_max(wk3.last_ms)
            hh3 = wk3.harvest(geometry=False)
            good3 = reduce_sum(float((hh3.status == 0).sum()))
            fc3 = reduce_sum(float(hh3.force_calls.sum()))
            st3 = np.zeros(24)
            for k_, v_ in zip(*np.unique(hh3.status, return_counts=True)):
                st3[int(k_)] += v_
            st3 = [reduce_sum(float(v)) for v in st3]
            c3["search"] = {"walkers": 4096, "ms": ms3s, "status_good": int(good3),
                            "searches_per_hour": good3 / (ms3s * 1e-3) * 3600.0,
                            "force_calls": int(fc3), "atom_force_evals_per_s": fc3 * N3 / (ms3s * 1e-3),
                            "status_histogram": {k_: int(v_) for k_, v_ in enumerate(st3) if v_ > 0},
                            "barrier_eV_median": float(np.median(hh3.energies[hh3.status == 0] - E03)) if (hh3.status == 0).any() else None,
                            "config": "2047-atom fcc vacancy cell on the Au row (a0 %.4f A), dimer rotations_max 20 + LBFGS "
                                      "memory 25, 0.01 eV/A, max_move 0.05; displacement sigma 0.3 A within 3.0 A (scaled) of "
                                      "a neighbour of the vacancy" % WL.A0_AU}
            wk3.close()
            spot3.close()

    # BASELINE.json configs[0]: the LJ13 cluster, LJ, dimer + CG with the reference's own saddle-search fixture
    # (client/unit_tests/SaddleSearchTest.cpp:33-54: atom 0 displaced by (+0.3, -0.2, 0), a random unit mode).  One
    # walker is what the reference runs; a queue of the same search from different random modes shows the batched rate.
    c1 = None
    if not args.no_c1:
        s1 = WL.load_structure("lj13.json")
        N1 = s1.n_atoms
        R1 = s1.positions.copy()
        R1[0, 0] += 0.3
        R1[0, 1] -= 0.2
        W1 = 4096
        rng1 = np.random.default_rng(13 + first)
        M1 = rng1.uniform(-1.0, 1.0, (W1, N1, 3))  # AtomMatrix::Random: uniform in [-1, 1], then normalised
        M1 /= np.linalg.norm(M1.reshape(W1, -1), axis=1)[:, None, None]
        prm1 = search_params(opt_method=0, opt_converged_force=0.01, max_move=0.1, dimer_converged_angle=0.01,
                             minmode_method=0, saddle_max_iterations=100, saddle_converged_force=0.05,
                             saddle_max_energy=20.0)
        ljpot = B200Potential("lj", device=local_rank)
        _, e1, _ = ljpot.forceBatch(1, N1, R1.reshape(1, -1), s1.atomic_nrs, s1.cell.reshape(9))
        E01 = float(e1[0])  # (the fixture takes the displaced structure's energy as the reactant energy)
        one = {}
        for nw, key in ((1, "single_walker"), (W1, "batched")):
            wk1 = Walkers(ljpot, s1.atomic_nrs, s1.fixed, s1.cell, np.broadcast_to(R1.reshape(1, -1), (nw, 3 * N1)).copy(),
                          M1[:nw].reshape(nw, -1).copy())
            if nw > 1:
                barrier()
            t0 = time.perf_counter()
            wk1.saddle_search(prm1, E01)
            torch.cuda.synchronize()
            wall1 = time.perf_counter() - t0
            if nw > 1:
                wall1 = reduce_max(wall1)
            h1 = wk1.harvest()
            wk1.close()
            if nw == 1:
                one = {"status": int(h1.status[0]), "iterations": int(h1.iterations[0]), "force_calls": int(h1.force_calls[0]),
                       "eigenvalue": float(h1.eigenvalues[0]), "energy_eV": float(h1.energies[0]), "wall_ms": 1e3 * wall1}
            else:
                good1 = reduce_sum(float((h1.status == 0).sum()))
                fc1 = reduce_sum(float(h1.force_calls.sum()))
                cnt1 = np.zeros(24)
                for k_, v_ in zip(*np.unique(h1.status, return_counts=True)):
                    cnt1[int(k_)] += v_
                hist1 = {k_: int(reduce_sum(float(v_))) for k_, v_ in enumerate(cnt1)}
                c1 = {"workload": "docs/lj13.con: 13-atom LJ cluster, dimer + CG (SaddleSearchTest.cpp:33-54 settings and "
                                  "displacement), one walker and a queue of %d random start modes per GPU" % W1,
                      "single_walker": one, "walkers": int(W1 * world), "wall_s": wall1, "status_good": int(good1),
                      "searches_per_hour": good1 / wall1 * 3600.0, "force_calls": int(fc1),
                      "atom_force_evals_per_s": fc1 * N1 / wall1,
                      "status_histogram": {k_: v_ for k_, v_ in hist1.items() if v_ > 0}}
        ljpot.close()

    # BASELINE.json configs[3] as a third data point: examples/akmc-pt, Morse_Pt, dimer + CG saddle searches from the
    # least-coordinated atoms followed by the minimisation of both minima per walker, classified like ProcessSearchJob
    c4 = None
    if not args.no_c4:
        from eon_b200 import displace as DP, process_search as PS
        cfg4 = dict(WL.AKMC_PT)
        s4 = WL.load_structure(cfg4["fixture"])
        N4 = s4.n_atoms
        mpot = B200Potential("morse_pt", device=local_rank, skin=args.skin)
        wk4 = Walkers(mpot, s4.atomic_nrs, s4.fixed, s4.cell, np.broadcast_to(s4.positions.reshape(1, -1), (W, 3 * N4)).copy())
        wk4.relax(search_params(opt_method=1, opt_converged_force=1e-4))
        hmin4 = wk4.harvest()
        Rmin4, E04 = hmin4.positions[0].reshape(N4, 3), float(hmin4.energies[0])
        cand = DP.min_coordinated_atoms(mpot, Rmin4, s4.cell, s4.fixed, 3.3)
        R4 = np.empty((W, N4, 3))
        M4 = np.empty((W, N4, 3))
        for k in range(W):
            rng = np.random.default_rng(cfg4["seed"] + first + k)
            R4[k], M4[k] = WL.displace(s4, Rmin4, int(cand[int(rng.integers(len(cand)))]), cfg4["radius"], cfg4["magnitude"],
                                       cfg4["seed"] + first + k)
        wk4.set(positions=R4, modes=M4)
        prm4 = search_params(**cfg4["search"])
        barrier()
        t0 = time.perf_counter()
        out4 = wk4.process_search(prm4, E04)
        torch.cuda.synchronize()
        wall4 = reduce_max(time.perf_counter() - t0)
        res4 = PS.process_results(out4, Rmin4, s4.cell, s4.fixed, max_energy=10.0, displacement_positions=R4,
                                  potential_type="MORSE_PT")
        st4 = np.array([r["termination_reason"] for r in res4])
        good4 = reduce_sum(float((st4 == 0).sum()))
        fc4 = reduce_sum(float(sum(r["total_force_calls"] for r in res4)))
        barr = [r["barrier_reactant_to_product"] for r in res4 if r["termination_reason"] == 0]
        cnt4 = np.zeros(24)
        for k_, v_ in zip(*np.unique(st4, return_counts=True)):
            cnt4[int(k_)] += v_
        hist4 = {k_: int(reduce_sum(float(v_))) for k_, v_ in enumerate(cnt4)}
        hist4 = {k_: v_ for k_, v_ in hist4.items() if v_ > 0}  # all ranks
        c4 = {"workload": "examples/akmc-pt: 343 Pt atoms (168 fixed), Morse_Pt, least-coordinated epicentres, sigma 0.1 A "
                          "within 3.3 A, dimer + CG 1e-3, then both minima relaxed per walker; host classification as "
                          "ProcessSearchJob (connectivity, barriers)",
              "walkers": int(W * world), "wall_s": wall4, "processes_found": int(good4),
              "process_searches_per_hour": good4 / wall4 * 3600.0,
              "status_histogram": hist4,
              "force_calls": int(fc4), "atom_force_evals_per_s": fc4 * N4 / wall4,
              "barrier_eV_min_median": [float(np.min(barr)), float(np.median(barr))] if barr else None}
        wk4.close()
        # the batched Morse_Pt force evaluation on its own (k_prepare + k_pair), against the FP64 roofline like the headline:
        # 53 flop per UNORDERED in-cutoff pair; the kernels visit every ordered pair
        sets4 = [torch.from_numpy(WL.jitter(Rmin4, W, JITTER, seed=4100 + k + first).reshape(W, 3 * N4)).to(dev) for k in range(2)]
        d_Z4 = torch.from_numpy(np.broadcast_to(s4.atomic_nrs, (W, N4)).copy()).to(dev)
        d_box4 = torch.from_numpy(np.broadcast_to(s4.cell.reshape(9), (W, 9)).copy()).to(dev)
        d_F4 = torch.empty((W, 3 * N4), dtype=torch.float64, device=dev)
        d_U4 = torch.empty((W,), dtype=torch.float64, device=dev)

        def step4(k):
            mpot.forceBatchDevice(W, N4, sets4[k % 2].data_ptr(), d_Z4.data_ptr(), d_F4.data_ptr(), d_U4.data_ptr(),
                                  d_box4.data_ptr(), sync=True)
            return mpot.last_force_ms()

        for k in range(3):
            step4(k)
        pairs4 = float(mpot.nl_stats()["ordered_pairs"])
        t4 = np.zeros(2)
        n4 = 10
        for k in range(n4):
            mpot.l2_flush(FLUSH_BYTES)
            t4 += np.array(step4(k))
        t4 = np.array([reduce_max(float(t4[0])), reduce_max(float(t4[1]))]) / n4
        flops4 = 53.0 * pairs4 / 2.0
        c4["force_kernel"] = {"ms_k_prepare": float(t4[0]), "ms_k_pair": float(t4[1]),
                              "atom_force_evals_per_s": W * N4 * world / (float(t4.sum()) * 1e-3),
                              "ordered_pairs_per_launch": pairs4, "tflops": flops4 / (float(t4.sum()) * 1e-3) / 1e12,
                              "fp64_roofline_frac": flops4 / (float(t4.sum()) * 1e-3) / fp64_peak if fp64_peak > 0 else None,
                              "note": "256 x 343 atoms per GPU, cutoff 9.5 A (~175 neighbours per atom), L2 flushed; both "
                                      "kernels walk every ordered pair of the cached list"}
        mpot.close()

    # BASELINE.json configs[4] as a fourth data point: examples/neb-al bands (reactant -> product, 7 movable images of 601
    # atoms), every movable image of every band in one forceBatch call, then the band kernel; host buffers throughout
    c5 = None
    if not args.no_c5:
        from eon_b200 import neb as NEB
        with open(os.path.join(WL.GOLDEN_DIR, "neb_al.json")) as fh:
            d5 = json.load(fh)
        s5 = WL.load_structure("neb_al.json")
        N5, M5, B5 = s5.n_atoms, 7, 64
        prod = np.array(d5["product_positions"], dtype=np.float64)
        inv5 = np.linalg.inv(s5.cell)
        dfrac = (prod - s5.positions) @ inv5
        path = (dfrac - np.floor(dfrac + 0.5)) @ s5.cell
        lam = np.linspace(0.0, 1.0, M5 + 2).reshape(1, -1, 1, 1)
        free5 = 1.0 - s5.fixed.reshape(1, 1, N5, 3)
        bands = np.empty((B5, M5 + 2, N5, 3))
        for b in range(B5):
            rng = np.random.default_rng(500 + first + b)
            bands[b] = s5.positions[None] + lam[0] * path[None] + rng.uniform(-0.02, 0.02, (M5 + 2, N5, 3)) * free5[0]
        bands[:, 0], bands[:, M5 + 1] = s5.positions, s5.positions + path
        R5 = bands.reshape(B5, M5 + 2, 3 * N5)
        npot = B200Potential("eam", device=local_rank, skin=args.skin, pin_host_buffers=1)
        with B200Potential("eam", device=local_rank, skin=args.skin) as hpot:  # two-call host path, for comparison
            _, Ue, _ = hpot.forceBatch(2, N5, np.ascontiguousarray(R5[0, [0, M5 + 1]]), s5.atomic_nrs, s5.cell)
            ends = np.broadcast_to(np.asarray(Ue).reshape(1, 2), (B5, 2))
            ref5 = NEB.band_forces(hpot, R5, s5.atomic_nrs, s5.cell, s5.fixed, endpoint_energies=ends, climbing_image=1)
            t0 = time.perf_counter()
            for _ in range(3):
                NEB.band_forces(hpot, R5, s5.atomic_nrs, s5.cell, s5.fixed, endpoint_energies=ends, climbing_image=1)
            wall5_host = (time.perf_counter() - t0) / 3
        out5 = None
        for _ in range(3):
            out5 = NEB.band_forces_resident(npot, R5, s5.atomic_nrs, s5.cell, s5.fixed, climbing_image=1)
        dev5 = float(np.abs(out5["projected"] - ref5["projected"]).max())
        barrier()
        reps5 = 10
        t0 = time.perf_counter()
        for _ in range(reps5):
            out5 = NEB.band_forces_resident(npot, R5, s5.atomic_nrs, s5.cell, s5.fixed, climbing_image=1)
        wall5 = reduce_max(time.perf_counter() - t0) / reps5
        c5 = {"workload": "examples/neb-al: 601 Al atoms (200 fixed), in-tree EAM, %d bands x %d movable images, linear "
                          "reactant -> product paths with U(-0.02,0.02) A jitter; eonb200_neb_band_forces: positions up, one "
                          "batched EAM launch over all images, one band-kernel launch, projected forces + energies back "
                          "(host buffers, page-locked by the handle)" % (B5, M5),
              "bands": int(B5 * world), "images_per_band": M5, "ms_per_band_evaluation_batch": wall5 * 1e3,
              "ms_two_call_host_path": wall5_host * 1e3, "max_abs_dev_vs_host_path": dev5,
              "image_force_evals_per_s": B5 * world * M5 / wall5, "atom_force_evals_per_s": B5 * world * M5 * N5 / wall5,
              "climbing_images": [int(x) for x in np.unique(out5["climbing"])],
              "max_projected_force": float(NEB.convergence_force(out5["projected"]).max()),
              "barrier_eV": float(out5["energies"][0].max() - out5["energies"][0, 0])}
        npot.close()

    # honest list accounting: the headline runs with every list valid (jitter well inside skin / 4); a search rebuilds
    # its lists at the rate observed in the search leg, and a step that rebuilds costs what the skin-0 run costs
    ms_reuse, ms_fresh = dev_ms_max / args.steps, t_fresh / args.steps
    frac = search["one_launch"]["list_rebuild_fraction"] if search else None
    build_bytes = 24.0 * W * N + 4.0 * pairs_per_step  # positions in, one 4-byte code per ordered pair out (no skin shell)
    ms_build = max(ms_fresh - ms_reuse, 1e-6)
    list_acc = {"rebuild_fraction_in_search_leg": frac,
                "ms_step_lists_valid": ms_reuse, "ms_step_rebuilding_every_list": ms_fresh,
                "value_search_mix": (W * N * world / (((1.0 - frac) * ms_reuse + frac * ms_fresh) * 1e-3)) if frac is not None else None,
                "nl_build": {"ms_per_batch_of_rebuilds": ms_build, "algorithmic_bytes": build_bytes,
                             "achieved_GBs": build_bytes / (ms_build * 1e-3) / 1e9,
                             "frac_of_hbm_peak": build_bytes / (ms_build * 1e-3) / 1e9 / hbm_peak,
                             "note": "rebuild time = (step at skin 0, every list rebuilt in the kernel) - (step with valid "
                                     "lists); bytes = 24 N + 4 P_o per system (SURVEY.md 8d)"}}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))  # the CPU arm gets every host core back
        cpu, _, _ = cpu_force_baseline(struct, sets[0], 6.0, True)
        cpu["variants"] = {}
        for kind in ("port", "port-nl"):
            v, _, _ = cpu_force_baseline(struct, sets[0], 4.0, False, kind=kind)
            cpu["variants"][kind] = {"value": v["value"], "sample": v["sample"]}
        cpu["variants"]["reference"] = {"value": cpu["value"], "sample": cpu["sample"]}

    if rank == 0:
        line = {"metric": "atom-force-evals/sec", "value": value, "unit": "atom-force-evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": dev_ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(N, world, args.skin, rebuilt),
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": "atom-force-evals/s", "h2d_bytes_per_step": h2d,
                                          "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                                          "ms_per_step_sustained_400_calls": e2e_sustained_ms,
                                          "how": "B200Potential.forceBatch on caller-owned contiguous numpy blocks, "
                                                 "page-locked by the handle; coordinates copied up in two chunks while the "
                                                 "first chunk's kernel runs, forces and energies written by the kernel "
                                                 "straight into the caller's page-locked blocks (bulk stores over PCIe, "
                                                 "no read-back copies); the d2h bytes are those posted writes",
                                          "cpp_shim": e2e_shim},
                "gpu_launches": int(launches), "roofline": roofline,
                "extra": {"wall_s_timed_region": t_wall, "value_hot_l2": units / world / (t_hot * 1e-3),
                          "value_fresh_list_every_step": units / world / (t_fresh * 1e-3),
                          "list_accounting": list_acc, "affinity": affinity,
                          "fp64_peak_tflops_measured": fp64_peak / 1e12, "c1": c1, "c3": c3, "c4": c4, "c5": c5},
                "search": search}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    pot.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skin", type=float, default=1.0)
    ap.add_argument("--tiles", type=int, default=0, help="CTAs per walker in the force passes (0 = automatic)")
    ap.add_argument("--lpa", type=int, default=0, help="lanes per atom (0 = automatic)")
    ap.add_argument("--no-search", action="store_true")
    ap.add_argument("--no-c1", action="store_true", help="skip the LJ13 dimer-search data point")
    ap.add_argument("--no-c3", action="store_true", help="skip the 4096 x 2047-atom data point")
    ap.add_argument("--no-c4", action="store_true", help="skip the akmc-pt Morse process-search data point")
    ap.add_argument("--no-c5", action="store_true", help="skip the neb-al band-force data point")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
