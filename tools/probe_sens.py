"""Timing probe of the sensitivity passes (development aid): python tools/probe_sens.py [2d|3d] [n] [reps]

For every physics: the element-per-thread pass with atomic nodal adds (no plan needed) and the tile-plan pass on a
scalar handle (node-window staging, plan-ordered sums, deterministic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pyfem_gpu_testflight_b200 as pf

dim = sys.argv[1] if len(sys.argv) > 1 else "2d"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
c = pf.ProblemCreator(n + 1, n + 1, n + 1) if dim == "3d" else pf.ProblemCreator(n + 1, n + 1)
d = 3 if dim == "3d" else 2
for phys, comps in (("elasticity", d), ("poisson", 1)):
    for path in ("atomic", "tile"):
        if path == "atomic":  # a handle without plan: the element-per-thread kernel
            mesh = pf.DeviceMesh(c.X, c.conn, comps, build_gather_plan=False)
        else:
            mesh = pf.DeviceMesh(c.X, c.conn, 1)
        rho = torch.rand(mesh.nnodes, dtype=torch.float64, device="cuda") * 0.9 + 0.1
        phi = torch.rand(mesh.nnodes * comps, dtype=torch.float64, device="cuda")
        psi = torch.rand(mesh.nnodes * comps, dtype=torch.float64, device="cuda")
        out = torch.empty(mesh.nnodes, dtype=torch.float64, device="cuda")
        for _ in range(3):
            mesh.k_dv_sens(phys, rho, 3.0, phi, psi, out=out, deterministic=(path == "tile"))
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mesh.k_dv_sens(phys, rho, 3.0, phi, psi, out=out, deterministic=(path == "tile")); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        alg = mesh.nelems * mesh.nnodes_per_elem * 4 + mesh.nnodes * 8 * (d + 2 + 2 * comps)  # conn, X, rho, phi, psi, out
        print(f"k_dv_sens {phys} {dim} n={n} [{path}]: best {min(ts):.3f} ms median {np.median(ts):.3f} ms "
              f"-> {mesh.nelems / min(ts) / 1e6:.2f} G elem/s, {alg / min(ts) / 1e6:.0f} GB/s algorithmic, "
              f"checksum {float(out.sum()):.6e}", flush=True)
        del mesh
