"""Device-resident timing of the NEB band kernel (k_neb_band) on the C5 shape: B bands x 7 movable images x 601 atoms.
Algorithmic traffic per movable image: its positions and forces read once, tangent and projected force written once
(4 x 3 N x 8 bytes; the neighbouring images' positions are shared with their own CTAs)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from eon_b200 import B200Potential, neb as NEB


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    M, N = 7, 601
    D = 3 * N
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    cell = np.array([28.638, 0, 0, 0, 28.638, 0, 0, 0, 40.125])
    base = torch.rand(B, 1, D, device=dev, dtype=torch.float64, generator=g) * 28.0
    R = (base + torch.linspace(0, 1, M + 2, device=dev, dtype=torch.float64).reshape(1, -1, 1)
         * torch.randn(B, 1, D, device=dev, dtype=torch.float64, generator=g) * 0.3).contiguous()
    F = torch.randn(B, M + 2, D, device=dev, dtype=torch.float64, generator=g)
    E = torch.randn(B, M + 2, device=dev, dtype=torch.float64, generator=g)
    P = torch.empty(B, M, D, device=dev, dtype=torch.float64)
    T = torch.empty(B, M, D, device=dev, dtype=torch.float64)
    ci = torch.zeros(B, device=dev, dtype=torch.int32)
    out = {}
    with B200Potential("eam", device=0) as pot:
        for name, kw in (("neb", dict(climbing_image=1)), ("dneb", dict(climbing_image=1, projection=1))):
            prm = NEB.neb_params(pot._lib, **kw)
            cb = cell.ctypes.data_as(C.POINTER(C.c_double))

            def launch():
                pot._check(pot._lib.eonb200_neb_forces_device(pot._h, B, M, N, 401, R.data_ptr(), F.data_ptr(), E.data_ptr(),
                                                              cb, C.byref(prm), P.data_ptr(), T.data_ptr(), ci.data_ptr()))
            for _ in range(3):
                launch()
            torch.cuda.synchronize()
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                launch()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / reps * 1e3
            alg = 4.0 * D * 8 * B * M
            out[name] = {"ms_per_launch": ms, "images_per_s": B * M / ms * 1e3, "algorithmic_GBps": alg / ms / 1e6,
                         "bytes_algorithmic": alg}
        # parity of the resident path against the host-buffer entry on a slice
        Ph, Th, cih = NEB.neb_forces(pot, R[:4].cpu().numpy(), F[:4].cpu().numpy(), E[:4].cpu().numpy(), cell, 401,
                                     climbing_image=1, projection=1)
        out["device_vs_host_max_abs"] = float(np.abs(Ph - P[:4].cpu().numpy()).max())
    out["config"] = {"bands": B, "images": M, "atoms": N, "working_set_MB": (R.numel() + F.numel() + P.numel() + T.numel()) * 8 / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
