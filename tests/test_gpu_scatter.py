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


# ---- chunk templates (pfg_internal.cuh): chunks with the same local topology share one set of plan tables ------------
def test_chunk_templates_share_tables_and_change_nothing(monkeypatch):
    """On a lattice-like mesh the interior chunks are byte-identical once their tables are relative to per-chunk
    bases: the plan read per assembly shrinks to the directory plus a few templates.  Sharing tables must not change
    one bit of the result (PFG_NO_TEMPLATES=1 builds the same plan without sharing), with or without a nodal field,
    for 2 x 2 blocks, scalar operators with a vector output, and hex8 scalar operators."""
    import torch
    import pyfem_gpu_testflight_b200 as pf
    X, conn = orc.structured_mesh(301, 187)
    # a graded tensor-product grid: still lattice-like for the chunking (nodes share coordinates along grid lines),
    # but every element has its own Jacobian.  (Random jitter would make the coordinate-based chunking irregular:
    # correct, but then hardly any two chunks share a topology.)
    X = np.stack([X[:, 0] ** 1.3, np.sin(0.5 * np.pi * X[:, 1] / X[:, 1].max())], axis=1)
    rho = 0.1 + 0.9 * np.random.default_rng(1).random(X.shape[0])
    u = np.random.default_rng(2).random(X.shape[0])
    xdv = np.ones(10) / 10.0
    shared2, shared1 = pf.DeviceMesh(X, conn, 2), pf.DeviceMesh(X, conn, 1)
    assert shared2.nchunks > 200 and shared2.ntemplates <= shared2.nchunks // 4
    assert shared1.ntemplates <= shared1.nchunks // 4
    monkeypatch.setenv("PFG_NO_TEMPLATES", "1")
    plain2, plain1 = pf.DeviceMesh(X, conn, 2), pf.DeviceMesh(X, conn, 1)
    monkeypatch.delenv("PFG_NO_TEMPLATES")
    assert plain2.ntemplates == plain2.nchunks == shared2.nchunks
    assert shared2.plan_bytes < plain2.plan_bytes // 4
    # only the templates' tables are kept: the handle itself shrinks
    from pyfem_gpu_testflight_b200 import _lib
    assert shared2.info(_lib.INFO_DEVICE_BYTES) < plain2.info(_lib.INFO_DEVICE_BYTES)
    for r, p in ((1.0, 0.0), (rho, 3.0)):
        assert torch.equal(shared2.assemble_elasticity(r, p, mode="gather"), plain2.assemble_elasticity(r, p, mode="gather"))
        assert torch.equal(shared1.assemble_poisson(r, p, mode="gather"), plain1.assemble_poisson(r, p, mode="gather"))
    Ka, ra = shared1.assemble_nlpoisson(xdv, u, mode="gather")
    Kb, rb = plain1.assemble_nlpoisson(xdv, u, mode="gather")
    assert torch.equal(Ka, Kb) and torch.equal(ra, rb)
    Ha, Ra = shared1.assemble_helmholtz(0.05, mode="gather")
    Hb, Rb = plain1.assemble_helmholtz(0.05, mode="gather")
    assert torch.equal(Ha, Hb) and torch.equal(Ra, Rb)
    Kref = orc.assemble_elasticity(X, conn, rho, 3.0)
    K = shared2.to_scipy(shared2.assemble_elasticity(rho, 3.0, mode="gather"))
    assert np.array_equal(K.indices, Kref.indices)
    assert_values_close(K.data, Kref.data, VAL_TOL)
    # hex8, scalar operator
    X3, c3 = orc.structured_mesh(23, 19, 17)
    h3 = pf.DeviceMesh(X3, c3, 1)
    assert h3.ntemplates < h3.nchunks
    rho3 = 0.1 + 0.9 * np.random.default_rng(3).random(X3.shape[0])
    K3 = h3.to_scipy(h3.assemble_poisson(rho3, 3.0, mode="gather"))
    assert_values_close(K3.data, orc.assemble_poisson(X3, c3, rho3, 3.0).data, VAL_TOL)


def test_element_mask_with_shared_templates():
    """pfg_mesh_set_element_mask on a handle whose chunks share tables: the skip flags live beside the records, not in
    the shared tables; masking every other element row and its complement adds up to the full matrix."""
    import torch
    import pyfem_gpu_testflight_b200 as pf
    nx, ny = 97, 61
    X, conn = orc.structured_mesh(nx, ny)
    mesh = pf.DeviceMesh(X, conn, 2)
    assert mesh.ntemplates < mesh.nchunks
    full = mesh.assemble_elasticity(1.0, 0.0, mode="gather").clone()
    rows = (np.arange(conn.shape[0]) // (nx - 1)) % 2
    mesh.set_element_mask(rows.astype(np.uint8))
    a = mesh.assemble_elasticity(1.0, 0.0, mode="gather").clone()
    mesh.set_element_mask((1 - rows).astype(np.uint8))
    b = mesh.assemble_elasticity(1.0, 0.0, mode="gather").clone()
    mesh.set_element_mask(None)
    assert torch.equal(mesh.assemble_elasticity(1.0, 0.0, mode="gather"), full)
    assert float((a + b - full).abs().max()) <= 1e-13 * float(full.abs().max())
    assert float(a.abs().max()) > 0 and float(b.abs().max()) > 0
