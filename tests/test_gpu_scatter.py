"""The reference's two-step interface on the device: element matrices without the scatter (Ke_mat, Re, rhs_e of
_compute_element_jacobian / _compute_element_rhs) and the scatter of caller-supplied element matrices / vectors
(ModelBase._assemble_jacobian, pyfem.py:920-931; _assemble_rhs, pyfem.py:860-875) -- the slot A2DWrapper uses."""
import numpy as np
import pytest

import pyfem_oracle as orc
from make_golden import jitter
from parity import VAL_TOL, assert_csr_matches, assert_values_close

pytestmark = pytest.mark.gpu
MODES = ["atomic", "gather"]


@pytest.fixture(scope="module")
def pf():
    import pyfem_gpu_testflight_b200 as pf
    return pf


def _mesh(three_d, permute=True):
    if three_d:
        X, conn = orc.structured_mesh(9, 7, 8)
        X = jitter(X, (9, 7, 8), seed=4)
    else:
        X, conn = orc.structured_mesh(37, 29)
        X = jitter(X, (37, 29), seed=4)
    if permute:
        conn = conn[np.random.default_rng(2).permutation(conn.shape[0])]
    return X, conn


@pytest.mark.parametrize("three_d", [False, True])
def test_element_matrices_match_oracle(pf, three_d):
    X, conn = _mesh(three_d)
    nn = X.shape[0]
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(nn)
    mesh = pf.DeviceMesh(X, conn, X.shape[1])
    Ke, _, _ = mesh.element_matrices("elasticity", field=rho, params=(3.0, 10.0, 0.3))
    assert_values_close(Ke.cpu().numpy(), orc.elasticity_Ke(X, conn, rho, 3.0), VAL_TOL, "elasticity Ke")
    mesh = pf.DeviceMesh(X, conn, 1)
    Ke, _, _ = mesh.element_matrices("poisson", field=rho, params=(2.0,))
    assert_values_close(Ke.cpu().numpy(), orc.poisson_Ke(X, conn, rho, 2.0), VAL_TOL, "poisson Ke")
    Ke, Re, _ = mesh.element_matrices("helmholtz", params=(0.07,), want_Ke2=True)
    Kr, Rr = orc.helmholtz_KeRe(X, conn, 0.07)
    assert_values_close(Ke.cpu().numpy(), Kr, VAL_TOL, "helmholtz Ke")
    assert_values_close(Re.cpu().numpy(), Rr, VAL_TOL, "helmholtz Re")
    if not three_d:
        Xn = (X - X.min(axis=0)) / (X.max(axis=0) - X.min(axis=0))
        mesh = pf.DeviceMesh(Xn, conn, 1)
        xdv = np.linspace(0.05, 0.3, 10)
        u = np.random.default_rng(5).random(nn) - 0.4
        Ke, _, fe = mesh.element_matrices("nlpoisson", field=u, params=xdv, want_fe=True)
        assert_values_close(Ke.cpu().numpy(), orc.nlpoisson_Ke(Xn, conn, xdv, u), VAL_TOL, "nlpoisson Ke")
        assert_values_close(fe.cpu().numpy(), orc.nlpoisson_res_e(Xn, conn, xdv, u), VAL_TOL, "nlpoisson res_e")


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", ["quad_m1", "quad_m2", "hex_m1", "hex_m3"])
def test_scatter_of_supplied_element_matrices(pf, case, mode):
    three_d = case.startswith("hex")
    m = int(case[-1])
    X, conn = _mesh(three_d)
    D = conn.shape[1] * m
    Ke = np.random.default_rng(9).standard_normal((conn.shape[0], D, D))  # arbitrary, non-symmetric
    Kr = orc.scatter_matrix(Ke, orc.conn_to_dof(conn, m))
    mesh = pf.DeviceMesh(X, conn, m)
    if case == "hex_m3" and mode == "gather":
        with pytest.raises(NotImplementedError):
            mesh.scatter_matrix(Ke, mode=mode)
        return
    K = mesh.to_scipy(mesh.scatter_matrix(Ke, mode=mode))
    assert_csr_matches(K, Kr.indptr, Kr.indices, Kr.data)
    if m == 1:
        fe = np.random.default_rng(10).standard_normal(conn.shape)
        want = orc.scatter_vector(fe, conn, X.shape[0], conn.shape[1])
        assert_values_close(mesh.scatter_vector(fe, mode=mode).cpu().numpy(), want, VAL_TOL, "scatter_vector")


def test_two_step_model_interface(pf):
    """model._compute_element_jacobian(model.Ke_mat) then model._assemble_jacobian(model.Ke_mat), as
    performance_test.py:52 and A2DWrapper.compute_jacobian (pyfem.py:2255-2277) use the reference."""
    X, conn = _mesh(False, permute=False)
    q = pf.QuadratureBilinear2D()
    b = pf.BasisBilinear2D(q)
    model = pf.LinearElasticity(X, conn, [0], None, {0: [0.0, 0.0]}, q, b)
    model._compute_element_jacobian(model.Ke_mat)
    assert model.Ke_mat.shape == (conn.shape[0], 8, 8)
    K2 = model._assemble_jacobian(model.Ke_mat)
    K1 = model.compute_jacobian()
    assert_csr_matches(K2, K1.indptr, K1.indices, K1.data)
    model = pf.LinearPoisson(X, conn, [0], None, q, b, lambda x: 1.0)
    model._compute_element_jacobian(model.Ke_mat)
    K2 = model._assemble_jacobian(model.Ke_mat)
    K1 = model.compute_jacobian()
    assert_csr_matches(K2, K1.indptr, K1.indices, K1.data)
    rhs_e = np.random.default_rng(3).random(conn.shape)
    rhs = np.empty(X.shape[0])
    model._assemble_rhs(rhs_e, rhs)
    assert_values_close(rhs, orc.scatter_vector(rhs_e, conn, X.shape[0], 4), VAL_TOL, "_assemble_rhs")
