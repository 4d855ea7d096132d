"""CPU-only checks: the C-ABI library loads and exports every symbol of include/pyfem_b200.h, the host
mirror of the reference interface (quadrature / basis tables, ProblemCreator, Dirichlet conditions) matches
the oracle / the reference, and the product path refuses to run without a GPU."""
import os
import re

import numpy as np
import pytest

import pyfem_oracle as orc
import ref_import
from parity import assert_values_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pyfem_gpu_testflight_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pyfem_b200.h")).read()
    declared = set(re.findall(r"PFG_API\s+(?:const\s+)?\w+\*?\s+(pfg_\w+)\s*\(", header))
    assert declared, "no PFG_API declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes and header disagree"
    assert lib.pfg_abi_version() == 1
    assert isinstance(lib.pfg_last_error(), bytes)


def test_invalid_arguments_are_reported_without_a_gpu():
    # argument validation happens before any CUDA call
    import ctypes
    from pyfem_gpu_testflight_b200 import _lib
    lib = _lib.load()
    handle = ctypes.c_void_p()
    rc = lib.pfg_mesh_create(ctypes.byref(handle), 3, 1, 10, 10, None, None, 0, 10, None, 0, 0, None)
    assert rc == _lib.PFG_ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _lib.check(rc)
    rc = lib.pfg_mesh_create(ctypes.byref(handle), 4, 3, 10, 10, None, None, 0, 10, None, 0, 0, None)
    assert rc == _lib.PFG_ERR_INVALID
    assert lib.pfg_assemble_poisson(None, None, 1.0, 0.0, None, 0, None) == _lib.PFG_ERR_INVALID
    assert lib.pfg_mesh_destroy(None) == _lib.PFG_OK


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pyfem_gpu_testflight_b200 as pf
    c = pf.ProblemCreator(5, 5)
    conn, X, dof_fixed = c.create_poisson_problem()
    q = pf.QuadratureBilinear2D()
    with pytest.raises(RuntimeError, match="CUDA"):
        pf.LinearPoisson(X, conn, dof_fixed, None, q, pf.BasisBilinear2D(q), lambda x: 1.0)
    with pytest.raises(NotImplementedError):
        pf.QuadratureTriangle2D()
    with pytest.raises(NotImplementedError):
        pf.ProblemCreator(5, 5, element_type="tri")


def test_tables_match_oracle():
    import pyfem_gpu_testflight_b200 as pf
    q = pf.QuadratureBilinear2D()
    b = pf.BasisBilinear2D(q)
    pts, w, N, dN = orc.quad4_tables()
    assert np.array_equal(q.get_pt(), pts) and np.array_equal(q.get_weight(), w) and q.get_nquads() == 4
    assert_values_close(b.eval_shape_fun(), N, 1e-15)
    assert_values_close(b.eval_shape_fun_deriv(), dN, 1e-15)
    q = pf.QuadratureBlock3D()
    b = pf.BasisBlock3D(q)
    pts, w, N, dN = orc.hex8_tables()
    assert np.array_equal(q.get_pt(), pts) and q.get_nquads() == 8
    assert_values_close(b.eval_shape_fun(), N, 1e-15)
    assert_values_close(b.eval_shape_fun_deriv(), dN, 1e-15)
    assert b.eval_shape_fun() is b.N  # cached, like the reference


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_tables_and_problem_creator_match_reference():
    import pyfem_gpu_testflight_b200 as pf
    ref = ref_import.load()
    for mine_q, mine_b, ref_q, ref_b in [
        (pf.QuadratureBilinear2D, pf.BasisBilinear2D, ref.QuadratureBilinear2D, ref.BasisBilinear2D),
        (pf.QuadratureBlock3D, pf.BasisBlock3D, ref.QuadratureBlock3D, ref.BasisBlock3D),
    ]:
        q, rq = mine_q(), ref_q()
        b, rb = mine_b(q), ref_b(rq)
        assert np.array_equal(q.get_pt(), rq.get_pt()) and np.array_equal(q.get_weight(), rq.get_weight())
        assert_values_close(b.eval_shape_fun(), rb.eval_shape_fun(), 1e-15)
        assert_values_close(b.eval_shape_fun_deriv(), rb.eval_shape_fun_deriv(), 1e-15)
    for kw in (dict(nnodes_x=9, nnodes_y=6), dict(nnodes_x=5, nnodes_y=4, nnodes_z=6, element_type="block"),
               dict(nnodes_x=7, nnodes_y=5, Lx=3.0, Ly=2.0)):
        a, r = pf.ProblemCreator(**kw), ref.ProblemCreator(**kw)
        assert np.array_equal(a.conn, r.conn) and np.array_equal(a.X, r.X)
        assert a.conn.dtype == r.conn.dtype
        ca, cr = a.create_poisson_problem(), r.create_poisson_problem()
        assert list(ca[2]) == list(cr[2])
        ea, er = a.create_linear_elasticity_problem(), r.create_linear_elasticity_problem()
        assert list(ea[2]) == list(er[2]) and ea[3] == er[3]
        ha, hr = a.create_helmhotz_problem(), r.create_helmhotz_problem()
        assert np.array_equal(ha[2], hr[2])


def _host_model(cls, **attrs):
    """A model object with host attributes only (no device handle): exercises the host-side glue."""
    m = cls.__new__(cls)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("symmetric", [True, False])
@pytest.mark.parametrize("with_vals", [True, False])
def test_dirichlet_host_matches_reference(symmetric, with_vals):
    import pyfem_gpu_testflight_b200 as pf
    ref = ref_import.load()
    c = ref.ProblemCreator(9, 7)
    conn, X, dof_fixed, force = c.create_linear_elasticity_problem()
    vals = np.random.default_rng(0).random(len(dof_fixed)) if with_vals else None
    q = ref.QuadratureBilinear2D()
    rm = ref.LinearElasticity(X, conn, dof_fixed, vals, force, q, ref.BasisBilinear2D(q))
    K = rm.compute_jacobian()
    rhs = rm.compute_rhs().copy()
    Kr, rr = rm.apply_dirichlet_bcs(K.copy(), rhs.copy(), symmetric)
    ndof = K.shape[0]
    mask = np.ones(ndof, dtype=bool)
    mask[dof_fixed] = False
    mm = _host_model(pf.LinearElasticity, dof_fixed=np.array(dof_fixed), dof_fixed_vals=vals,
                     _dof_free=np.nonzero(mask)[0], ndof=ndof)
    Km, rm2 = mm.apply_dirichlet_bcs(K.copy(), rhs.copy(), symmetric)
    assert np.array_equal(Km.indptr, Kr.indptr) and np.array_equal(Km.indices, Kr.indices)
    assert_values_close(Km.data, Kr.data, 1e-15)
    assert_values_close(rm2, rr, 1e-14)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_oracle_complex_density_matches_reference():
    """The oracle evaluated in complex arithmetic is what the GPU's complex-rho assembly is checked against
    (tests/test_gpu_parity.py::test_complex_step_derivative_like_reference_tests): pin it to the unmodified reference's
    complex-step path (pyfem.py:1018-1020, 1289-1292, 1783-1785, 1933-1936) on quad and hex meshes."""
    ref = ref_import.load()
    for dims in ((9, 7), (5, 4, 6)):
        three_d = len(dims) == 3
        c = ref.ProblemCreator(*dims, element_type="block" if three_d else "quad")
        q = ref.QuadratureBlock3D() if three_d else ref.QuadratureBilinear2D()
        b = ref.BasisBlock3D(q) if three_d else ref.BasisBilinear2D(q)
        rng = np.random.default_rng(0)
        conn, X, dof_fixed = c.create_poisson_problem()
        rho = rng.random(X.shape[0]) + 1j * 0.25 * rng.random(X.shape[0])
        Kr = ref.LinearPoisson(X, conn, dof_fixed, None, q, b, lambda xq: 1.0, p=5.0).compute_jacobian(rho)
        Ko = orc.assemble_poisson(np.asarray(X, dtype=float), np.asarray(conn), rho, 5.0)
        assert np.array_equal(Kr.indices, Ko.indices) and Kr.dtype == Ko.dtype == np.complex128
        assert_values_close(Ko.data.real, Kr.data.real, 1e-13)
        assert_values_close(Ko.data.imag, Kr.data.imag, 1e-13)
        conn, X, dof_fixed, force = c.create_linear_elasticity_problem()
        Kr = ref.LinearElasticity(X, conn, dof_fixed, None, force, q, b, p=5.0).compute_jacobian(rho)
        Ko = orc.assemble_elasticity(np.asarray(X, dtype=float), np.asarray(conn), rho, 5.0)
        assert np.array_equal(Kr.indices, Ko.indices)
        assert_values_close(Ko.data.real, Kr.data.real, 1e-13)
        assert_values_close(Ko.data.imag, Kr.data.imag, 1e-13)


# ---- the algebraic identities the re-derived element operators rest on (csrc/pfg_elem.cuh), checked in numpy ---------
def test_hex8_gradients_through_the_trilinear_modes():
    """hex8_row_modes / hex8_modes_to_nodes: 8 grad N_b = sum over the seven non-constant sign monomials sigma_k(b) of
    M_k with M_x, M_y, M_z the rows of the adjugate and M_xy = eta A0 + xi A1, ..., M_xyz = eta zeta A0 + xi zeta A1 +
    xi eta A2; hence sum_q s_q G_a (x) G_b = (1/8) sum_k sigma_k(b) V_k with V_k = sum_q s_q G_a (x) M_k, including the
    eighth node (no constant mode: the zero row sums)."""
    pts, _, _, dN = orc.hex8_tables()
    rng = np.random.default_rng(0)
    A = rng.standard_normal((8, 3, 3))  # one adjugate per quadrature point
    s = rng.random(8) + 0.5
    sx = np.array([-1.0, 1.0, 1.0, -1.0, -1.0, 1.0, 1.0, -1.0])
    sy = np.array([-1.0, -1.0, 1.0, 1.0, -1.0, -1.0, 1.0, 1.0])
    sz = np.array([-1.0, -1.0, -1.0, -1.0, 1.0, 1.0, 1.0, 1.0])
    sigma = np.stack([sx, sy, sz, sx * sy, sx * sz, sy * sz, sx * sy * sz])  # (7, node)
    G = np.einsum("qbk,qkl->qbl", dN, A)  # G_b,l = sum_k dN_b,k A[k][l]  (hex8_geo)
    M = np.zeros((8, 7, 3))
    for q, (xi, eta, zeta) in enumerate(pts):
        M[q] = [A[q, 0], A[q, 1], A[q, 2], eta * A[q, 0] + xi * A[q, 1], zeta * A[q, 0] + xi * A[q, 2],
                zeta * A[q, 1] + eta * A[q, 2], eta * zeta * A[q, 0] + xi * zeta * A[q, 1] + xi * eta * A[q, 2]]
    assert np.allclose(8.0 * G, np.einsum("kb,qkl->qbl", sigma, M), rtol=0, atol=1e-13)
    for a in range(8):
        P = np.einsum("q,qi,qbj->bij", s, G[:, a], G)            # the row blocks of node a
        V = np.einsum("q,qi,qkj->kij", s, G[:, a], M)            # the seven running sums of a lane
        assert np.allclose(P, np.einsum("kb,kij->bij", sigma, V) / 8.0, rtol=0, atol=1e-12)
    assert np.allclose(sigma.sum(axis=1), 0.0)  # every mode sums to zero over the nodes: the rows of Ke do too


def test_nlpoisson_newton_term_from_eight_numbers():
    """NlPoissonQuad4Op: T[a][b] = sum_q c2 (G_a . grad u) N_b with G_a . grad u = dN_a/dxi P + dN_a/deta R, formed from
    (P_q, R_q) by two one-dimensional contractions (the basis tables are tensor products of 1 +- g)."""
    pts, _, N, dN = orc.quad4_tables()
    rng = np.random.default_rng(1)
    Pq, Rq = rng.standard_normal(4), rng.standard_normal(4)
    T_ref = np.einsum("qa,q,qb->ab", dN[:, :, 0], Pq, N) + np.einsum("qa,q,qb->ab", dN[:, :, 1], Rq, N)
    g = 1.0 / np.sqrt(3.0)
    fp, fm = 1.0 + g, 1.0 - g
    sxs, sys = [0, 1, 1, 0], [0, 0, 1, 1]          # sign bits of node / point 0..3
    qidx = lambda sx, sy: (2 if sx else 3) if sy else (1 if sx else 0)
    Bs, Bm, Cs, Cm = np.zeros((2, 2)), np.zeros(2), np.zeros((2, 2)), np.zeros(2)
    for sb in range(2):
        a0 = fp * Pq[qidx(sb, 0)] + fm * Pq[qidx(1 - sb, 0)]
        a1 = fp * Pq[qidx(sb, 1)] + fm * Pq[qidx(1 - sb, 1)]
        Bs[sb] = [fp * fp * a0 + fm * fm * a1, fm * fm * a0 + fp * fp * a1]
        Bm[sb] = fp * fm * (a0 + a1)
        d0 = fp * Rq[qidx(0, sb)] + fm * Rq[qidx(0, 1 - sb)]
        d1 = fp * Rq[qidx(1, sb)] + fm * Rq[qidx(1, 1 - sb)]
        Cs[sb] = [fp * fp * d0 + fm * fm * d1, fm * fm * d0 + fp * fp * d1]
        Cm[sb] = fp * fm * (d0 + d1)
    T = np.zeros((4, 4))
    for a in range(4):
        for b in range(4):
            pb = Bs[sxs[b]][sys[a]] if sys[a] == sys[b] else Bm[sxs[b]]
            rb = Cs[sys[b]][sxs[a]] if sxs[a] == sxs[b] else Cm[sys[b]]
            T[a, b] = ((pb if sxs[a] else -pb) + (rb if sys[a] else -rb)) / 16.0
    assert np.allclose(T, T_ref, rtol=0, atol=1e-14)


def test_bilinear_and_trilinear_coefficient_forms():
    """quad4_field4 / hex8_field8: nodal values -> monomial coefficients; the reference-space derivatives they give at
    the quadrature points equal sum_a dN_a f_a (the form the sensitivity kernels and the nonlinear operator use)."""
    rng = np.random.default_rng(2)
    pts, _, N, dN = orc.quad4_tables()
    f = rng.standard_normal(4)
    m, a, b, c = f.sum(), (f[1] - f[0]) + (f[2] - f[3]), (f[3] - f[0]) + (f[2] - f[1]), (f[2] - f[3]) - (f[1] - f[0])
    for q, (xi, eta) in enumerate(pts):
        assert np.allclose([0.25 * (m + a * xi + b * eta + c * xi * eta), 0.25 * (a + c * eta), 0.25 * (b + c * xi)],
                           [N[q] @ f, dN[q, :, 0] @ f, dN[q, :, 1] @ f], rtol=0, atol=1e-14)
    pts, _, N, dN = orc.hex8_tables()
    f = rng.standard_normal(8)
    s00, d00, s10, d10 = f[1] + f[0], f[1] - f[0], f[2] + f[3], f[2] - f[3]
    s01, d01, s11, d11 = f[5] + f[4], f[5] - f[4], f[6] + f[7], f[6] - f[7]
    ss0, sd0, ds0, dd0 = s10 + s00, s10 - s00, d10 + d00, d10 - d00
    ss1, sd1, ds1, dd1 = s11 + s01, s11 - s01, d11 + d01, d11 - d01
    cz, cy, cyz, cx, cxz, cxy, cxyz = ss1 - ss0, sd1 + sd0, sd1 - sd0, ds1 + ds0, ds1 - ds0, dd1 + dd0, dd1 - dd0
    for q, (xi, eta, zeta) in enumerate(pts):
        grad8 = [cx + cxy * eta + cxz * zeta + cxyz * eta * zeta, cy + cxy * xi + cyz * zeta + cxyz * xi * zeta,
                 cz + cyz * eta + cxz * xi + cxyz * xi * eta]
        assert np.allclose(np.array(grad8) / 8.0, dN[q].T @ f, rtol=0, atol=1e-14)
