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
