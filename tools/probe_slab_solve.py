"""Timing probe of the row-slab conjugate gradients (development aid):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/probe_slab_solve.py [n] [iters] [graph|eager]
Poisson on an n x n quad mesh split into N row slabs (structured_slab, no global mesh on any rank), the first mesh
line fixed; a fixed number of iterations (the tolerance is unreachable on purpose) timed with CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import pyfem_gpu_testflight_b200 as pf
from pyfem_gpu_testflight_b200.partition import slab_node_ranges, structured_slab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
part = structured_slab(n + 1, n + 1, None, rank, world)
ranges = slab_node_ranges(n + 1, n + 1, None, world)
q = pf.QuadratureBilinear2D()
fixed = np.arange(n + 1)
model = pf.LinearPoisson(part.X, part.conn, fixed, None, q, pf.BasisBilinear2D(q), lambda Xq: 1.0, partition=part,
                         node_ranges=ranges, device=dev)
vals = model.compute_jacobian_device(1.0)
rhs = model.compute_rhs()
graph = None if len(sys.argv) <= 3 else (sys.argv[3] == "graph")
if graph is not None and model._slab_cg is None and world > 1:
    model._slab_solver()
if graph is not None and world > 1:
    import functools
    model._slab_cg.cg = functools.partial(model._slab_cg.cg, graph=graph)
for rep in range(2):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        model.solve_device(vals.clone(), rhs, rtol=1e-300, max_iter=iters)
    except RuntimeError as exc:  # "cg failed": max_iter reached, as intended
        pass
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0 and rep == 1:
        ex = model._slab_cg.exchange.bytes_per_refresh if model._slab_cg is not None else 0
        mode = "" if graph is None else (" [graph replay]" if graph and model._slab_cg._graph_ok else " [eager]")
        print(f"slab cg poisson n={n} ranks={world}{mode}: {iters} iterations {ms.item():.2f} ms -> {ms.item() / iters:.4f} ms / iteration, "
              f"{(n + 1) ** 2 / 1e6:.1f} M unknowns, halo {ex} B sent per rank and iteration", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
