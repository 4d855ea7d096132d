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
