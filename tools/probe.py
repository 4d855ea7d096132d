"""Quick device timing probe (development aid): python tools/probe.py [physics] [n] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pyfem_gpu_testflight_b200 as pf
from pyfem_gpu_testflight_b200 import _lib

phys = sys.argv[1] if len(sys.argv) > 1 else "elasticity2d"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["gather", "atomic"]

t0 = time.time()
if phys.endswith("3d"):
    c = pf.ProblemCreator(n + 1, n + 1, n + 1)
else:
    c = pf.ProblemCreator(n + 1, n + 1)
m = {"elasticity2d": 2, "elasticity3d": 3}.get(phys, 1)
print(f"mesh: {c.conn.shape[0]} elems, host gen {time.time()-t0:.2f}s", flush=True)
torch.cuda.synchronize()
t0 = time.time()
mesh = pf.DeviceMesh(c.X, c.conn, m)
torch.cuda.synchronize()
print(f"create: {time.time()-t0:.2f}s nnz={mesh.nnz} nchunks={mesh.nchunks} chunk_elems={mesh.chunk_elems} "
      f"redundancy={mesh.chunk_elems/max(1,mesh.nelems):.3f} plan_bytes={mesh.plan_bytes/1e6:.1f}MB "
      f"dev_bytes={mesh.info(_lib.INFO_DEVICE_BYTES)/1e9:.2f}GB maxk={mesh.info(_lib.INFO_MAX_ROW_BLOCKS)}", flush=True)
E = mesh.nelems
vals = mesh.new_values()
vals2 = mesh.new_values() if phys.startswith("helm") else None
u = torch.rand(mesh.nnodes, dtype=torch.float64, device="cuda")
rho = torch.rand(mesh.nnodes, dtype=torch.float64, device="cuda") * 0.9 + 0.1
res = mesh.new_vector()

def run(mode, use_rho):
    r = rho if use_rho else 1.0
    if phys.startswith("elasticity"):
        mesh.assemble_elasticity(r, 5.0 if use_rho else 0.0, out=vals, mode=mode)
    elif phys.startswith("poisson"):
        mesh.assemble_poisson(r, 5.0 if use_rho else 0.0, out=vals, mode=mode)
    elif phys.startswith("helm"):
        mesh.assemble_helmholtz(0.05, out_K=vals, out_R=vals2, mode=mode)
    elif phys.startswith("nl"):
        mesh.assemble_nlpoisson(np.ones(10) / 10, u, out_K=vals, out_res=res, mode=mode)

for mode in modes:
    for use_rho in (False, True):
        if use_rho and not (phys.startswith("elasticity") or phys.startswith("poisson")):
            continue
        for _ in range(3):
            run(mode, use_rho)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(mode, use_rho); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = np.array(ts)
        print(f"{phys} n={n} mode={mode} rho={'field' if use_rho else 'const'}: best {ts.min():.3f} ms median {np.median(ts):.3f} ms "
              f"-> {E/ts.min()/1e6:.2f} G elem/s (best), checksum {float(vals.sum()):.6e}", flush=True)
