import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench, ctypes
from eon_b200 import B200Potential, _capi
lib = _capi.bind(ctypes.CDLL(os.path.abspath("build/v/lib_dbg.so")))
struct, sets = bench.force_workload(0)
N, W = struct.n_atoms, 256
dev = torch.device("cuda", 0)
pot = B200Potential("eam", device=0, skin=1.0, lib=lib)
d_R = torch.from_numpy(sets.reshape(4, W, 3 * N)).to(dev)
d_Z = torch.from_numpy(np.broadcast_to(struct.atomic_nrs, (W, N)).copy()).to(dev)
d_box = torch.from_numpy(np.broadcast_to(struct.cell.reshape(9), (W, 9)).copy()).to(dev)
d_F = torch.empty((W, 3 * N), dtype=torch.float64, device=dev); d_U = torch.empty((W,), dtype=torch.float64, device=dev)
def step(k):
    pot.forceBatchDevice(W, N, d_R[k % 4].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(), d_box.data_ptr(), sync=True)
for k in range(4): step(k)
pot.phase_clocks(True); pot.l2_flush(512 << 20); step(5)
c = pot.phase_clocks(False).astype(np.float64)
for a, b, nm in [(9, 15, "t0 leaves recv"), (9, 13, "last warp leaves recv"), (9, 6, "last warp arrives at reduction barrier"), (9, 7, "t0 passes barrier"), (9, 14, "t0 after reduction"), (9, 10, "exit")]:
    d = (c[:, b] - c[:, a]) / 1.965e3
    print("%-40s med %.2f p90 %.2f max %.2f" % (nm, np.median(d), np.percentile(d, 90), d.max()))
