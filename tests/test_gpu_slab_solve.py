"""Device solves on a matrix distributed by row slabs (slab_solve.py / pfg_cg_dist): every rank is a process with its
own slab handle, the model API call is the single-GPU one (Assembler.solve(method="cg", device=True),
model.compliance(rho, device=True)) and the gathered solution must equal the oracle's dense solve, as in the
reference's tests/test_linear_poisson.py:18-40 and tests/test_elasticity.py:22-51.

Two transports: NCCL with one GPU per rank (needs >= 2 GPUs), and gloo with all ranks SHARING GPU 0 (host-staged
exchange), which exercises the same device path on a one-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dense_solve(K, rhs, fixed):
    A = np.array(K.todense())
    b = np.array(rhs, dtype=float)
    A[fixed, :] = 0.0
    A[:, fixed] = 0.0
    A[fixed, fixed] = 1.0
    b[fixed] = 0.0
    return np.linalg.solve(A, b)


def _worker(rank, world, port, backend, halo, out_dir):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import pyfem_oracle as orc
    from make_golden import test_gfunc as gfunc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import pyfem_gpu_testflight_b200 as pf
    ok = True
    q = pf.QuadratureBilinear2D()
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    rho = 0.2 + 0.8 * np.random.default_rng(0).random(32 * 32)

    # thermal problem: Assembler.solve on the device, every rank holds its rows of u
    conn, X, dof_fixed = creator.create_poisson_problem()
    model = pf.LinearPoisson(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), gfunc, p=3.0, group=dist.group.WORLD,
                             halo=halo, device=dev)
    asm = pf.Assembler(model)
    u = model.gather_vector(asm.solve(method="cg", device=True))
    c_dev, _ = model.compliance(rho, solver="cg", device=True)
    if rank == 0:
        fixed = np.asarray(dof_fixed)
        rhs = orc.assemble_poisson_rhs(X, conn, gfunc)
        u_ref = _dense_solve(orc.assemble_poisson(X, conn), rhs, fixed)
        ok &= np.max(np.abs(u - u_ref)) <= 1e-7 * np.max(np.abs(u_ref)) and 0 < asm.last_iterations < 2000
        u_rho = _dense_solve(orc.assemble_poisson(X, conn, rho, 3.0), rhs, fixed)
        b = np.array(rhs, dtype=float)
        b[fixed] = 0.0
        ok &= abs(c_dev - b @ u_rho) <= 1e-6 * abs(b @ u_rho)

    # plane stress, two dofs per node
    conn, X, dof_fixed, nodal_force = creator.create_linear_elasticity_problem()
    model = pf.LinearElasticity(X, conn, dof_fixed, None, nodal_force, q, pf.BasisBilinear2D(q), p=3.0,
                                group=dist.group.WORLD, halo=halo, device=dev)
    c_dev, u_own = model.compliance(rho, solver="cg", device=True)
    u = model.gather_vector(u_own)
    # the gradient from the rank's own rows of u (ghost entries travel in one halo exchange) and from the global u
    grad = model.gather_vector(model.compliance_grad(rho, u_own))
    u_all = [None]
    if rank == 0:
        u_all = [u]
    dist.broadcast_object_list(u_all, src=0)
    grad_g = model.gather_vector(model.compliance_grad(rho, u_all[0]))
    if rank == 0:
        fixed = np.asarray(dof_fixed)
        rhs = orc.elasticity_point_loads(2 * X.shape[0], 2, nodal_force)
        u_ref = _dense_solve(orc.assemble_elasticity(X, conn, rho, 3.0), rhs, fixed)
        b = np.array(rhs, dtype=float)
        b[fixed] = 0.0
        ok &= np.max(np.abs(u - u_ref)) <= 1e-6 * np.max(np.abs(u_ref))
        ok &= abs(c_dev - b @ u_ref) <= 1e-6 * abs(b @ u_ref)
        g_ref = -orc.elasticity_K_dv_sens(X, conn, rho, 3.0, u, u)  # pyfem.py:1835-1847 on the same u
        ok &= np.max(np.abs(grad - g_ref)) <= 1e-10 * np.max(np.abs(g_ref))
        ok &= np.max(np.abs(grad_g - g_ref)) <= 1e-10 * np.max(np.abs(g_ref))

    # hex8, three dofs per node: the slab solve against the single-handle device solve of the same system
    X3, conn3 = orc.structured_mesh(7, 6, 9)
    X3 = X3 + np.random.default_rng(3).uniform(-0.01, 0.01, size=X3.shape)
    fixed3 = np.arange(3 * 7 * 6)  # the z = 0 plane
    force3 = {int(X3.shape[0] - 1): [0.0, 0.0, -1.0]}
    q3 = pf.QuadratureBlock3D()
    model = pf.LinearElasticity(X3, conn3, fixed3, None, force3, q3, pf.BasisBlock3D(q3), group=dist.group.WORLD,
                                halo=halo, device=dev)
    u = model.gather_vector(pf.Assembler(model).solve(method="cg", device=True))
    if rank == 0:
        rhs = orc.elasticity_point_loads(3 * X3.shape[0], 3, force3)
        u_ref = _dense_solve(orc.assemble_elasticity(X3, conn3), rhs, fixed3)
        ok &= np.max(np.abs(u - u_ref)) <= 1e-6 * np.max(np.abs(u_ref))
    # Helmholtz filter and its transposed application over the slabs (tests/test_helmholtz.py:11-44 compares the
    # filtered field at 1e-8)
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=17)
    conn, X = creator.create_helmhotz_problem()[:2]
    hm = pf.Helmholtz(0.07, X, conn, q, pf.BasisBilinear2D(q), group=dist.group.WORLD, halo=halo, device=dev)
    xf = np.random.default_rng(11).random(X.shape[0])
    gb, ge = hm.slab.owned_nodes
    rho_f = hm.gather_vector(hm.apply_device(xf).cpu().numpy())
    grad_f = hm.gather_vector(hm.apply_gradient_device(xf[gb:ge]).cpu().numpy())
    if rank == 0:
        Kh, Rh = orc.assemble_helmholtz(X, conn, 0.07)
        Kd = np.array(Kh.todense())
        ok &= np.max(np.abs(rho_f - np.linalg.solve(Kd, Rh @ xf))) <= 1e-7
        ok &= np.max(np.abs(grad_f - Rh.T @ np.linalg.solve(Kd, xf))) <= 1e-7
    # Newton loop of the nonlinear Poisson problem over the slabs (BiCGStab over all ranks) against the host loop of
    # a single handle with a direct solve, as tests/test_nonlinear_poisson.py:12-42 compares p.u
    import contextlib
    import io
    creator = pf.ProblemCreator(nnodes_x=24, nnodes_y=19)
    conn, X, dof_fixed = creator.create_poisson_problem()
    X = X / X.max(axis=0)
    xdv = 0.2 + np.random.default_rng(7).random(10)
    model = pf.NonlinearPoisson2D(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), group=dist.group.WORLD, halo=halo,
                                  device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        u = model.gather_vector(pf.Assembler(model).solve_nonlinear(xdv=xdv, device=True))
    if rank == 0:
        single = pf.NonlinearPoisson2D(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), device=dev)
        with contextlib.redirect_stdout(io.StringIO()):
            u_ref = pf.Assembler(single).solve_nonlinear(method="direct", xdv=xdv)
        ok &= np.max(np.abs(u - u_ref)) <= 1e-6 * np.max(np.abs(u_ref)) and np.max(np.abs(u_ref)) > 0
    flag = torch.tensor([1 if ok else 0])
    if backend == "nccl":
        flag = flag.to(dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write(str(int(flag.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, "ghost"), (3, "ghost")])
def test_slab_solve_ranks_sharing_one_gpu(tmp_path, world, halo):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, "gloo", halo, str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "ok").read() == "1"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("halo", ["ghost", "nccl"])
def test_slab_solve_two_gpus_nccl(tmp_path, halo):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, "nccl", halo, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
