"""GPU tests shaped like the reference's own tests (tests/test_linear_poisson.py:18-40, test_elasticity.py:22-51,
test_nonlinear_poisson.py:12-42, test_helmholtz.py:11-44): solve on the 32 x 32-node ProblemCreator mesh through
Assembler / the model API and compare the scalar p.u for a seeded random p.  The reference compares against its
hand-written ref_*.py solvers; here the second solution comes from the numpy oracle's matrices and a dense solve.
"""
import numpy as np
import pytest

import pyfem_oracle as orc
from make_golden import test_gfunc as gfunc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pf():
    import pyfem_gpu_testflight_b200 as pf
    return pf


def _dense_solve(K, rhs, fixed, vals=None):
    A = np.array(K.todense())
    b = np.array(rhs, dtype=float)
    u0 = np.zeros(len(b))
    if vals is not None:
        u0[fixed] = vals
        b = b - A @ u0
    A[fixed, :] = 0.0
    A[:, fixed] = 0.0
    A[fixed, fixed] = 1.0
    b[fixed] = 0.0 if vals is None else vals
    return np.linalg.solve(A, b)


def _ptu(u):
    np.random.seed(123)  # as the reference tests do
    p = np.random.rand(len(u))
    return float(p @ u)


def test_linear_poisson_case(pf):
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    conn, X, dof_fixed = creator.create_poisson_problem()
    q = pf.QuadratureBilinear2D()
    model = pf.LinearPoisson(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), gfunc)
    u = pf.Assembler(model).solve(method="direct")
    u_ref = _dense_solve(orc.assemble_poisson(X, conn), orc.assemble_poisson_rhs(X, conn, gfunc), np.asarray(dof_fixed))
    assert abs(_ptu(u) - _ptu(u_ref)) <= 1e-10 * abs(_ptu(u_ref))


def test_plane_stress_case(pf):
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    conn, X, dof_fixed, nodal_force = creator.create_linear_elasticity_problem()
    q = pf.QuadratureBilinear2D()
    model = pf.LinearElasticity(X, conn, dof_fixed, None, nodal_force, q, pf.BasisBilinear2D(q))
    u = pf.Assembler(model).solve(method="direct")
    rhs = orc.elasticity_point_loads(2 * X.shape[0], 2, nodal_force)
    u_ref = _dense_solve(orc.assemble_elasticity(X, conn), rhs, np.asarray(dof_fixed))
    assert abs(_ptu(u) - _ptu(u_ref)) <= 1e-10 * abs(_ptu(u_ref))


def test_nonlinear_poisson_newton(pf, capsys):
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    conn, X, dof_fixed = creator.create_poisson_problem()
    X = X / X.max(axis=0)
    q = pf.QuadratureBilinear2D()
    model = pf.NonlinearPoisson2D(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q))
    xdv = np.ones(10) / 10.0
    u = pf.Assembler(model).solve_nonlinear(method="direct", xdv=xdv)
    assert "pyfem" in capsys.readouterr().out  # the reference's residual log line (pyfem.py:2345)
    # the same Newton iteration on the oracle's matrices
    fixed = np.asarray(dof_fixed)
    v = np.zeros(X.shape[0])
    r0 = None
    for k in range(10):
        K, res = orc.assemble_nlpoisson(X, conn, xdv, v)
        res = res.copy()
        res[fixed] = 0.0
        nrm = np.linalg.norm(res)
        if k == 0:
            r0 = nrm
        elif nrm < 1e-10 * r0 or nrm < 1e-12:
            break
        v -= _dense_solve(K, res, fixed)
    assert abs(_ptu(u) - _ptu(v)) <= 1e-8 * abs(_ptu(v))


def test_helmholtz_filter(pf):
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    conn, X, x = creator.create_helmhotz_problem()
    q = pf.QuadratureBilinear2D()
    r0 = 0.05
    model = pf.Helmholtz(r0, X, conn, q, pf.BasisBilinear2D(q))
    rho = model.apply(x)
    K, R = orc.assemble_helmholtz(X, conn, r0)
    rho_ref = np.linalg.solve(np.array(K.todense()), R @ x)
    assert abs(_ptu(rho) - _ptu(rho_ref)) <= 1e-8 * abs(_ptu(rho_ref))
    g = np.random.default_rng(2).random(len(x))
    assert np.allclose(model.apply_gradient(g), R.T @ np.linalg.solve(np.array(K.todense()), g), rtol=1e-9, atol=1e-12)


# ---- boundary conditions + the solve on the device CSR (K never leaves HBM) -------------------
def test_device_cg_poisson_and_elasticity(pf):
    """Assembler.solve(method="cg", device=True): pfg_apply_dirichlet + pfg_cg against the oracle's dense solve, on the
    reference tests' 32 x 32-node mesh (tests/test_linear_poisson.py:18-40, test_elasticity.py:22-51)."""
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    conn, X, dof_fixed = creator.create_poisson_problem()
    q = pf.QuadratureBilinear2D()
    model = pf.LinearPoisson(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), gfunc)
    asm = pf.Assembler(model)
    u = asm.solve(method="cg", device=True)
    u_ref = _dense_solve(orc.assemble_poisson(X, conn), orc.assemble_poisson_rhs(X, conn, gfunc), np.asarray(dof_fixed))
    assert abs(_ptu(u) - _ptu(u_ref)) <= 1e-8 * abs(_ptu(u_ref))
    assert 0 < asm.last_iterations < 2000
    conn, X, dof_fixed, nodal_force = creator.create_linear_elasticity_problem()
    model = pf.LinearElasticity(X, conn, dof_fixed, None, nodal_force, q, pf.BasisBilinear2D(q))
    u = pf.Assembler(model).solve(method="cg", device=True)
    rhs = orc.elasticity_point_loads(2 * X.shape[0], 2, nodal_force)
    u_ref = _dense_solve(orc.assemble_elasticity(X, conn), rhs, np.asarray(dof_fixed))
    assert abs(_ptu(u) - _ptu(u_ref)) <= 1e-8 * abs(_ptu(u_ref))


def test_device_cg_nonzero_dirichlet_values(pf):
    """Prescribed non-zero values: the device rhs gets rhs[free] -= K_free,fixed u0 (pyfem.py:830-834)."""
    creator = pf.ProblemCreator(nnodes_x=24, nnodes_y=17)
    conn, X, dof_fixed = creator.create_poisson_problem()
    fixed = np.asarray(dof_fixed)
    vals = np.random.default_rng(4).random(len(fixed))
    q = pf.QuadratureBilinear2D()
    model = pf.LinearPoisson(X, conn, fixed, vals, q, pf.BasisBilinear2D(q), gfunc)
    u = pf.Assembler(model).solve(method="cg", device=True)
    u_ref = _dense_solve(orc.assemble_poisson(X, conn), orc.assemble_poisson_rhs(X, conn, gfunc), fixed, vals)
    assert np.max(np.abs(u - u_ref)) <= 1e-7 * np.max(np.abs(u_ref))
    assert np.allclose(u[fixed], vals, rtol=0, atol=1e-12)


def test_compliance_default_solver_and_device(pf):
    """model.compliance(rho) with the reference's default solver='cg' (pyfem.py:1033, 1796) must work without pyamg;
    the device variant agrees with the direct solve."""
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    conn, X, dof_fixed, nodal_force = creator.create_linear_elasticity_problem()
    q = pf.QuadratureBilinear2D()
    rho = 0.2 + 0.8 * np.random.default_rng(0).random(X.shape[0])
    model = pf.LinearElasticity(X, conn, dof_fixed, None, nodal_force, q, pf.BasisBilinear2D(q), p=3.0)
    c_direct, u_direct = model.compliance(rho, solver="direct")
    c_default, _ = model.compliance(rho)
    c_dev, u_dev = model.compliance(rho, solver="cg", device=True)
    assert abs(c_default - c_direct) <= 1e-6 * abs(c_direct)
    assert abs(c_dev - c_direct) <= 1e-6 * abs(c_direct)
    assert np.max(np.abs(u_dev - u_direct)) <= 1e-6 * np.max(np.abs(u_direct))
    conn, X, dof_fixed = creator.create_poisson_problem()
    pm = pf.LinearPoisson(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), gfunc, p=3.0)
    c_direct, _ = pm.compliance(rho, solver="direct")
    c_default, _ = pm.compliance(rho)
    c_dev, _ = pm.compliance(rho, solver="cg", device=True)
    assert abs(c_default - c_direct) <= 1e-6 * abs(c_direct)
    assert abs(c_dev - c_direct) <= 1e-6 * abs(c_direct)


def test_helmholtz_device_filter_and_transpose(pf):
    """R^T x without forming the transpose (pfg_spmv_t) and the device filter solves (pyfem.py:2102-2115)."""
    import torch
    creator = pf.ProblemCreator(nnodes_x=29, nnodes_y=23)
    conn, X, x = creator.create_helmhotz_problem()
    X = X + np.random.default_rng(1).uniform(-0.004, 0.004, size=X.shape)
    q = pf.QuadratureBilinear2D()
    model = pf.Helmholtz(0.05, X, conn, q, pf.BasisBilinear2D(q))
    K, R = orc.assemble_helmholtz(X, conn, 0.05)
    g = np.random.default_rng(2).random(len(x))
    yt = model.mesh.spmv_t(model.R_device, g).cpu().numpy()
    assert np.max(np.abs(yt - R.T @ g)) <= 1e-13 * np.max(np.abs(R.T @ g))
    # a non-symmetric matrix in the same pattern: the transposed product must pick the mirrored entries
    vals = torch.rand(model.mesh.nnz, dtype=torch.float64, device=model.mesh.device, generator=torch.Generator(model.mesh.device).manual_seed(7))
    A = model.mesh.to_scipy(vals)
    assert np.max(np.abs(model.mesh.spmv_t(vals, g).cpu().numpy() - A.T @ g)) <= 1e-13 * np.max(np.abs(A.T @ g))
    assert np.max(np.abs(model.mesh.spmv(vals, g).cpu().numpy() - A @ g)) <= 1e-13 * np.max(np.abs(A @ g))
    Kd = np.array(K.todense())
    rho = model.apply_device(x, rtol=1e-12).cpu().numpy()
    assert np.max(np.abs(rho - np.linalg.solve(Kd, R @ x))) <= 1e-9
    grad = model.apply_gradient_device(g, rtol=1e-12).cpu().numpy()
    assert np.allclose(grad, R.T @ np.linalg.solve(Kd, g), rtol=1e-8, atol=1e-11)


def test_hex_spmv_transpose_three_dofs(pf):
    """pfg_spmv / pfg_spmv_t on a 3-dof hex handle (m = 3 block layout)."""
    import torch
    X, conn = orc.structured_mesh(6, 5, 4)
    mesh = pf.DeviceMesh(X, conn, 3)
    vals = torch.rand(mesh.nnz, dtype=torch.float64, device=mesh.device, generator=torch.Generator(mesh.device).manual_seed(3))
    A = mesh.to_scipy(vals)
    g = np.random.default_rng(5).random(mesh.nrows)
    assert np.max(np.abs(mesh.spmv(vals, g).cpu().numpy() - A @ g)) <= 1e-13 * np.max(np.abs(A @ g))
    assert np.max(np.abs(mesh.spmv_t(vals, g).cpu().numpy() - A.T @ g)) <= 1e-13 * np.max(np.abs(A.T @ g))


def test_newton_loop_on_the_device(pf, capsys):
    """Assembler.solve_nonlinear(device=True): Jacobian + residual from one fused assembly, boundary conditions on the
    device CSR, the step by Jacobi-preconditioned BiCGStab (pfg_bicgstab) -- against the host Newton loop with a direct
    solve (tests/test_nonlinear_poisson.py:12-42 compares at 1e-8)."""
    creator = pf.ProblemCreator(nnodes_x=32, nnodes_y=32)
    conn, X, dof_fixed = creator.create_poisson_problem()
    X = X / X.max(axis=0)
    q = pf.QuadratureBilinear2D()
    model = pf.NonlinearPoisson2D(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q))
    xdv = np.ones(10) / 10.0
    asm = pf.Assembler(model)
    u_dev = asm.solve_nonlinear(xdv=xdv, device=True)
    assert "pyfem" in capsys.readouterr().out and len(asm.last_iterations) >= 2
    u_host = pf.Assembler(model).solve_nonlinear(method="direct", xdv=xdv)
    assert abs(_ptu(u_dev) - _ptu(u_host)) <= 1e-8 * abs(_ptu(u_host))
    assert np.max(np.abs(u_dev - u_host)) <= 1e-7 * np.max(np.abs(u_host))


def test_bicgstab_nonsymmetric_system(pf):
    """pfg_bicgstab on a diagonally dominant non-symmetric matrix in the mesh pattern, against scipy's direct solve."""
    import torch
    from scipy.sparse.linalg import spsolve
    X, conn = orc.structured_mesh(23, 19)
    mesh = pf.DeviceMesh(X, conn, 2)
    g = torch.Generator(mesh.device).manual_seed(5)
    vals = torch.rand(mesh.nnz, dtype=torch.float64, device=mesh.device, generator=g) - 0.5
    A = mesh.to_scipy(vals)
    A.setdiag(np.abs(A).sum(axis=1).A1 + 1.0)  # diagonally dominant
    vals = torch.as_tensor(A.data, device=mesh.device)
    b = np.random.default_rng(2).random(mesh.nrows) - 0.5
    x, iters, resid = mesh.bicgstab(vals, b, rtol=1e-12)
    ref = spsolve(A.tocsc(), b)
    assert np.max(np.abs(x.cpu().numpy() - ref)) <= 1e-9 * np.max(np.abs(ref)) and iters > 0


def test_compliance_on_the_device_at_1024(pf):
    """SURVEY 8f #4 at a size where the host path would move 300 MB per call: compliance(rho, device=True) on 1024^2
    quads -- assembly, Dirichlet conditions and CG in HBM -- against the residual of the solved system."""
    import torch
    n = 1024
    creator = pf.ProblemCreator(nnodes_x=n + 1, nnodes_y=n + 1)
    conn, X, dof_fixed = creator.create_poisson_problem()
    q = pf.QuadratureBilinear2D()
    model = pf.LinearPoisson(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), gfunc, p=3.0)
    rho = 0.3 + 0.7 * np.random.default_rng(0).random(X.shape[0])
    c, u = model.compliance(rho, solver="cg", device=True)
    assert np.isfinite(c) and c > 0
    # residual of the boundary-conditioned system, evaluated on the device
    vals = model.compute_jacobian_device(rho)
    rhs = torch.as_tensor(model.compute_rhs()).to(model.mesh.device)
    model.mesh.apply_dirichlet(vals, rhs, model.dof_fixed, None, enforce_symmetric=True)
    r = model.mesh.spmv(vals, u) - rhs
    assert float(torch.linalg.vector_norm(r)) <= 2e-8 * float(torch.linalg.vector_norm(rhs))
