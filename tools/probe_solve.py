"""Timing probe of the device consumers (development aid): python tools/probe_solve.py [n] [reps]

SpMV / transposed SpMV on the CSR values of an n x n plane-stress assembly (bytes moved = values + column blocks +
vectors), and the per-iteration cost of the Jacobi-preconditioned CG on an n x n Poisson problem."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pyfem_gpu_testflight_b200 as pf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


c = pf.ProblemCreator(n + 1, n + 1)
for m, name in ((2, "elasticity"), (1, "poisson")):
    mesh = pf.DeviceMesh(c.X, c.conn, m)
    vals = mesh.assemble_elasticity(1.0, 0.0) if m == 2 else mesh.assemble_poisson(1.0, 0.0)
    x = torch.rand(mesh.ncols, dtype=torch.float64, device="cuda")
    y = mesh.new_vector()
    alg = mesh.nnz * 8 + (mesh.nnz // (m * m)) * 4 + mesh.nrows * 16  # values, node-level columns, x and y once
    for label, fn in (("spmv", lambda: mesh.spmv(vals, x, out=y)), ("spmv_t", lambda: mesh.spmv_t(vals, x, out=y))):
        ms = timed(fn)
        print(f"{label} {name} n={n}: {ms:.3f} ms, {alg / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    del mesh, vals
conn, X, dof_fixed = c.create_poisson_problem()
q = pf.QuadratureBilinear2D()
model = pf.LinearPoisson(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), lambda Xq: 1.0)
vals = model.compute_jacobian_device()
rhs = torch.as_tensor(model.compute_rhs()).to("cuda")
model.mesh.apply_dirichlet(vals, rhs, model.dof_fixed, None, enforce_symmetric=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
try:
    u, iters, resid = model.mesh.cg(vals, rhs, rtol=1e-8, max_iter=400)
except RuntimeError:
    iters = 400
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"cg poisson n={n}: {iters} iterations in {dt * 1e3:.1f} ms = {dt / iters * 1e6:.1f} us per iteration "
      f"({model.mesh.nnz * 8 / (dt / iters) / 1e9:.0f} GB/s of CSR values)", flush=True)
