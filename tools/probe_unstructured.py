"""Development aid: the tile kernel on a mesh whose chunking is irregular (randomly jittered coordinates defeat the
lattice detection, so chunks come from the plain sort-tile-recursive split and hardly any two share a template):
timing, template statistics, and agreement of the gather and atomic paths.  python tools/probe_unstructured.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pyfem_gpu_testflight_b200 as pf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
c = pf.ProblemCreator(n + 1, n + 1)
h = 1.0 / n
X = c.X + np.random.default_rng(0).uniform(-0.2 * h, 0.2 * h, size=c.X.shape)
conn = c.conn[np.random.default_rng(1).permutation(c.conn.shape[0])]  # element order carries no structure either
mesh = pf.DeviceMesh(X, conn, 2)
print(f"jittered + permuted {n}^2 quads: nchunks={mesh.nchunks} templates={mesh.ntemplates} "
      f"halo={mesh.chunk_elems / mesh.nelems:.3f} plan_read={mesh.plan_bytes / 1e6:.1f} MB", flush=True)
vals = mesh.new_values()
rho = torch.rand(mesh.nnodes, dtype=torch.float64, device="cuda") * 0.9 + 0.1
for mode in ("gather", "atomic"):
    for _ in range(3):
        mesh.assemble_elasticity(rho, 3.0, out=vals, mode=mode)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mesh.assemble_elasticity(rho, 3.0, out=vals, mode=mode); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"  {mode}: best {min(ts):.3f} ms -> {mesh.nelems / min(ts) / 1e6:.2f} G elem/s", flush=True)
    if mode == "gather":
        ref = vals.clone()
print(f"  max|gather - atomic| / max|K| = {float((vals - ref).abs().max() / ref.abs().max()):.2e}", flush=True)
