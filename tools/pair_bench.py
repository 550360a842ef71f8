"""Device-resident throughput of the LJ / Morse pair kernels (k_prepare + k_pair) on a batch of jittered walkers.
   python tools/pair_bench.py [morse_pt|lj] [walkers] [tiles] [lanes per atom]"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from eon_b200 import B200Potential, workloads as WL
kind = sys.argv[1] if len(sys.argv) > 1 else "morse_pt"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
s = WL.load_structure("akmc_pt.json" if kind == "morse_pt" else "lj13.json")
N = s.n_atoms
dev = torch.device("cuda", 0)
sets = [torch.from_numpy(WL.jitter(s.positions, W, 0.05, seed=k).reshape(W, 3 * N)).to(dev) for k in range(4)]
d_Z = torch.from_numpy(np.broadcast_to(s.atomic_nrs, (W, N)).copy()).to(dev)
d_box = torch.from_numpy(np.broadcast_to(s.cell.reshape(9), (W, 9)).copy()).to(dev)
d_F = torch.empty((W, 3 * N), dtype=torch.float64, device=dev); d_U = torch.empty((W,), dtype=torch.float64, device=dev)
tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lpa = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pot = B200Potential(kind, device=0, tiles_override=tiles, lpa_override=lpa)
def step(k):
    pot.forceBatchDevice(W, N, sets[k % 4].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(), d_box.data_ptr(), sync=True)
    return pot.last_force_ms()
for k in range(4): step(k)
pairs = pot.nl_stats()["ordered_pairs"]
t = np.zeros(2)
n = 20
for k in range(n):
    pot.l2_flush(512 << 20)
    t += np.array(step(k))
t /= n
flops = (53.0 if kind == "morse_pt" else 35.0) * pairs / 2
print(f"tiles {tiles} lpa {lpa} | {kind}: {W} walkers x {N} atoms, {pairs} ordered pairs; k_prepare {1e3*t[0]:.1f} us + k_pair {1e3*t[1]:.1f} us per step; "
      f"{W*N/(t.sum()*1e-3):.3e} atom-force-evals/s; {flops/(t.sum()*1e-3)/1e12:.2f} TFLOP/s (SURVEY 8d count)")
