"""GPU parity: the CUDA path (through the Python model API -> ctypes -> C ABI) against
 (1) the reference's own outputs committed under tests/golden (bit-exact pattern, values 1e-12 norm-relative),
 (2) the numpy oracle on larger seeded meshes, for both scatter strategies.
Run on the B200 box:  python -m pytest tests -m gpu
"""
import numpy as np
import pytest

import pyfem_oracle as orc
from make_golden import jitter, test_gfunc as gfunc
from parity import VAL_TOL, assert_csr_matches, assert_pattern_equal, assert_values_close, golden_files

pytestmark = pytest.mark.gpu

MODES = ["atomic", "gather"]


def _objs(pf, nne):
    q = pf.QuadratureBilinear2D() if nne == 4 else pf.QuadratureBlock3D()
    b = pf.BasisBilinear2D(q) if nne == 4 else pf.BasisBlock3D(q)
    return q, b


def _rho(g):
    return g["rho"] if g["rho"].ndim else float(g["rho"])


@pytest.fixture(scope="module")
def pf():
    import pyfem_gpu_testflight_b200 as pf
    return pf


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("path", golden_files("poisson"))
def test_poisson_golden(pf, path, mode):
    g = np.load(path)
    q, b = _objs(pf, g["conn"].shape[1])
    m = pf.LinearPoisson(g["X"], g["conn"], [0], None, q, b, gfunc, p=float(g["p"]), scatter=mode)
    K = m.compute_jacobian(_rho(g)) if "ramp" in path else m.compute_jacobian()
    assert_csr_matches(K, g["K_indptr"], g["K_indices"], g["K_data"])
    assert K.shape == tuple(g["K_shape"])
    rhs = m.compute_rhs()
    assert rhs is m.rhs
    assert_values_close(rhs, g["rhs"], VAL_TOL, "rhs")


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("path", golden_files("elasticity"))
def test_elasticity_golden(pf, path, mode):
    g = np.load(path)
    q, b = _objs(pf, g["conn"].shape[1])
    m = pf.LinearElasticity(g["X"], g["conn"], [0], None, {0: [0.0] * g["X"].shape[1]}, q, b, E=float(g["E"]),
                            nu=float(g["nu"]), p=float(g["p"]), scatter=mode)
    K = m.compute_jacobian(_rho(g))
    assert_csr_matches(K, g["K_indptr"], g["K_indices"], g["K_data"])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("path", golden_files("helmholtz"))
def test_helmholtz_golden(pf, path, mode):
    g = np.load(path)
    q, b = _objs(pf, g["conn"].shape[1])
    m = pf.Helmholtz(float(g["r0"]), g["X"], g["conn"], q, b, scatter=mode)
    assert_csr_matches(m.K, g["K_indptr"], g["K_indices"], g["K_data"])
    assert_csr_matches(m.R, g["R_indptr"], g["R_indices"], g["R_data"])
    assert m.compute_jacobian() is m.K
    assert_values_close(m.compute_rhs(g["x"]), g["rhs"], VAL_TOL, "R x")
    assert_values_close(m.compute_rhs_device(g["x"]).cpu().numpy(), g["rhs"], VAL_TOL, "device R x")


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("path", golden_files("nlpoisson"))
def test_nlpoisson_golden(pf, path, mode):
    g = np.load(path)
    q, b = _objs(pf, 4)
    m = pf.NonlinearPoisson2D(g["X"], g["conn"], [0], None, q, b, scatter=mode)
    K = m.compute_jacobian(g["xdv"], g["u"])
    assert_csr_matches(K, g["K_indptr"], g["K_indices"], g["K_data"])
    res = m.compute_rhs(g["xdv"], g["u"])
    assert_values_close(res, g["res"], VAL_TOL, "residual")
    Kd, rd = m.assemble_device(g["xdv"], g["u"])  # fused pass gives the same numbers
    assert np.array_equal(Kd.cpu().numpy(), K.data) or mode == "atomic"
    assert_values_close(rd.cpu().numpy(), g["res"], VAL_TOL, "fused residual")


# ---- larger seeded meshes against the oracle ---------------------------------------------------------
def _quad_case(nx, ny, seed=1, permute=False):
    X, conn = orc.structured_mesh(nx, ny)
    X = jitter(X, (nx, ny), seed=seed)
    if permute:
        conn = conn[np.random.default_rng(seed).permutation(conn.shape[0])]
    return X, conn


def _hex_case(nx, ny, nz, seed=1, permute=False):
    X, conn = orc.structured_mesh(nx, ny, nz)
    X = jitter(X, (nx, ny, nz), seed=seed)
    if permute:
        conn = conn[np.random.default_rng(seed).permutation(conn.shape[0])]
    return X, conn


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("permute", [False, True])
def test_quad_models_vs_oracle(pf, mode, permute):
    X, conn = _quad_case(131, 97, seed=3, permute=permute)
    q, b = _objs(pf, 4)
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0, 0.0]}, q, b, p=5.0, scatter=mode)
    K = m.compute_jacobian(rho)
    Kr = orc.assemble_elasticity(X, conn, rho, 5.0)
    assert_csr_matches(K, Kr.indptr, Kr.indices, Kr.data)
    m = pf.LinearPoisson(X, conn, [0], None, q, b, gfunc, p=2.0, scatter=mode)
    Kr = orc.assemble_poisson(X, conn, rho, 2.0)
    assert_csr_matches(m.compute_jacobian(rho), Kr.indptr, Kr.indices, Kr.data)
    assert_values_close(m.compute_rhs(), orc.assemble_poisson_rhs(X, conn, gfunc), VAL_TOL, "rhs")
    m = pf.Helmholtz(0.05, X, conn, q, b, scatter=mode)
    Kr, Rr = orc.assemble_helmholtz(X, conn, 0.05)
    assert_csr_matches(m.K, Kr.indptr, Kr.indices, Kr.data)
    assert_csr_matches(m.R, Rr.indptr, Rr.indices, Rr.data)
    Xn = (X - X.min(axis=0)) / (X.max(axis=0) - X.min(axis=0))
    m = pf.NonlinearPoisson2D(Xn, conn, [0], None, q, b, scatter=mode)
    xdv = np.ones(10) / 10.0
    u = np.random.default_rng(5).random(X.shape[0]) - 0.4
    Kr, rr = orc.assemble_nlpoisson(Xn, conn, xdv, u)
    assert_csr_matches(m.compute_jacobian(xdv, u), Kr.indptr, Kr.indices, Kr.data)
    assert_values_close(m.compute_rhs(xdv, u), rr, VAL_TOL, "residual")


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("permute", [False, True])
def test_hex_models_vs_oracle(pf, mode, permute):
    X, conn = _hex_case(19, 14, 11, seed=2, permute=permute)
    q, b = _objs(pf, 8)
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0, 0.0, 0.0]}, q, b, p=5.0, scatter=mode)
    Kr = orc.assemble_elasticity(X, conn, rho, 5.0)
    assert_csr_matches(m.compute_jacobian(rho), Kr.indptr, Kr.indices, Kr.data)
    m = pf.LinearPoisson(X, conn, [0], None, q, b, gfunc, p=2.0, scatter=mode)
    Kr = orc.assemble_poisson(X, conn, rho, 2.0)
    assert_csr_matches(m.compute_jacobian(rho), Kr.indptr, Kr.indices, Kr.data)
    assert_values_close(m.compute_rhs(), orc.assemble_poisson_rhs(X, conn, gfunc), VAL_TOL, "rhs")
    m = pf.Helmholtz(0.05, X, conn, q, b, scatter=mode)
    Kr, Rr = orc.assemble_helmholtz(X, conn, 0.05)
    assert_csr_matches(m.K, Kr.indptr, Kr.indices, Kr.data)
    assert_csr_matches(m.R, Rr.indptr, Rr.indices, Rr.data)


def test_gather_is_bitwise_reproducible_and_order_free(pf):
    # the gather path sums in plan order: two runs agree bit for bit, and chunking by node id gives the
    # same pattern as the coordinate tiling
    X, conn = _quad_case(97, 61, seed=9)
    a = pf.DeviceMesh(X, conn, 2)
    b = pf.DeviceMesh(X, conn, 2, reorder=False)
    v1 = a.assemble_elasticity(mode="gather").cpu().numpy()
    v2 = a.assemble_elasticity(mode="gather").cpu().numpy()
    v3 = b.assemble_elasticity(mode="gather").cpu().numpy()
    assert np.array_equal(v1, v2)
    assert_values_close(v3, v1, 1e-14)
    pa, pb = a.pattern_host(), b.pattern_host()
    assert_pattern_equal(pa[0], pa[1], pb[0], pb[1])


def test_dirichlet_on_device_matches_host(pf):
    c = pf.ProblemCreator(23, 17)
    conn, X, dof_fixed, force = c.create_linear_elasticity_problem()
    q, b = _objs(pf, 4)
    vals = np.random.default_rng(1).random(len(dof_fixed))
    for sym in (True, False):
        m = pf.LinearElasticity(X, conn, dof_fixed, vals, force, q, b)
        Kd = m.compute_jacobian_device()
        K = m.mesh.to_scipy(Kd)
        rhs = m.compute_rhs().copy()
        import torch
        rhs_d = torch.as_tensor(rhs).cuda()
        m.mesh.apply_dirichlet(Kd, rhs_d, dof_fixed, vals, enforce_symmetric=sym)
        K2 = m.mesh.to_scipy(Kd)
        K2.eliminate_zeros()
        Kh, rh = m.apply_dirichlet_bcs(K, rhs, enforce_symmetric_K=sym)
        assert (abs(K2 - Kh)).max() <= 1e-14 * abs(Kh).max()
        assert_values_close(rhs_d.cpu().numpy(), rh, 1e-13, "bc rhs")


def test_bad_inputs_raise(pf):
    X, conn = orc.structured_mesh(4, 4)
    q, b = _objs(pf, 4)
    with pytest.raises(AssertionError):  # reference asserts conn.max() == nnodes-1 (pyfem.py:681)
        pf.LinearPoisson(np.vstack([X, [[9.0, 9.0]]]), conn, [0], None, q, b, gfunc)
    with pytest.raises(NotImplementedError):  # sensitivities take a real density (the reference's do too)
        pf.LinearPoisson(X, conn, [0], None, q, b, gfunc)._compute_K_dv_sens(np.ones(16, dtype=complex), X[:, 0], X[:, 1])
    with pytest.raises(NotImplementedError):
        pf.LinearPoisson(X, conn[:, :3], [0], None, q, b, gfunc)


@pytest.mark.parametrize("three_d", [False, True])
@pytest.mark.parametrize("size", [2, 3])
def test_row_slabs_reproduce_global_matrix(pf, three_d, size):
    """The multi-GPU decomposition, with the ranks emulated one after another on this GPU: each rank's
    handle owns a row slab and gets its element block + ghost layer; the concatenated slabs are the global
    CSR (pattern bit-exact incl. global column ids, values to 1e-12)."""
    from pyfem_gpu_testflight_b200.partition import concat_slabs, partition_mesh
    if three_d:
        X, conn = _hex_case(9, 8, 13, seed=4)
    else:
        X, conn = _quad_case(41, 57, seed=4)
    m = X.shape[1]
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    Kg = orc.assemble_elasticity(X, conn, rho, 3.0)
    slabs = []
    for r in range(size):
        part = partition_mesh(X, conn, r, size)
        mesh = pf.DeviceMesh(part.X, part.conn, m, own_range=part.own_range, node_gid=part.node_gid,
                             ncols_nodes=part.nnodes_global)
        vals = mesh.assemble_elasticity(rho[part.node_gid], 3.0, mode="gather")
        vals_a = mesh.assemble_elasticity(rho[part.node_gid], 3.0, mode="atomic")
        assert_values_close(vals_a.cpu().numpy(), vals.cpu().numpy(), 1e-13)
        indptr, indices = mesh.pattern_host()
        slabs.append((indptr, indices, vals.cpu().numpy()))
        assert mesh.ncols == Kg.shape[1]
    K = concat_slabs(slabs, Kg.shape[1])
    assert np.array_equal(K.indptr, Kg.indptr) and np.array_equal(K.indices, Kg.indices)
    assert_values_close(K.data, Kg.data, VAL_TOL)


# ---- the model API as one rank of a row-slab partition (constructor keywords group= / partition=) ---------
@pytest.mark.parametrize("three_d", [False, True])
def test_model_slab_mode_matches_global_matrix(pf, three_d):
    """LinearElasticity / LinearPoisson / Helmholtz / NonlinearPoisson2D built with partition= (ghost-layer variant,
    ranks run one after another on this GPU): the row slabs, concatenated, are the oracle's global matrix -- pattern
    bit-exact including the dtype, values to 1e-12 -- and the Dirichlet conditions act on a slab as on its rows."""
    from pyfem_gpu_testflight_b200.partition import concat_slabs, partition_mesh, split_range
    dims = (9, 8, 11) if three_d else (41, 37, None)
    X, conn = orc.structured_mesh(*dims)
    X = X + np.random.default_rng(3).uniform(-0.004, 0.004, size=X.shape)
    if not three_d:
        X = (X - X.min(axis=0)) / (X.max(axis=0) - X.min(axis=0))
    size = 3
    ranges = split_range(X.shape[0], size)
    q = pf.QuadratureBlock3D() if three_d else pf.QuadratureBilinear2D()
    basis = pf.BasisBlock3D(q) if three_d else pf.BasisBilinear2D(q)
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    m = X.shape[1]
    fixed = np.arange(0, m * X.shape[0], 7)
    Kg = orc.assemble_elasticity(X, conn, rho, 3.0)
    Pg = orc.assemble_poisson(X, conn, rho, 3.0)
    slabs_e, slabs_p, bc_rows = [], [], []
    for r in range(size):
        part = partition_mesh(X, conn, r, size)
        me = pf.LinearElasticity(X, conn, fixed, None, {0: [0.0] * m}, q, basis, p=3.0, partition=part,
                                 node_ranges=ranges)
        K = me.compute_jacobian(rho)  # the GLOBAL nodal field, as a caller of the reference passes it
        gb, ge = part.owned_global_range
        assert K.shape == (m * (ge - gb), m * X.shape[0]) and K.indices.dtype == Kg.indices.dtype
        slabs_e.append((K.indptr.copy(), K.indices.copy(), K.data.copy()))
        Kb, rb = me.apply_dirichlet_bcs(K, me.compute_rhs(), enforce_symmetric_K=True)
        bc_rows.append(Kb)
        mp_ = pf.LinearPoisson(X, conn, [0], None, q, basis, gfunc, p=3.0, partition=part, node_ranges=ranges)
        P = mp_.compute_jacobian(rho)
        slabs_p.append((P.indptr.copy(), P.indices.copy(), P.data.copy()))
        rhs = mp_.compute_rhs()
        assert_values_close(rhs, orc.assemble_poisson_rhs(X, conn, gfunc)[gb:ge], VAL_TOL, "rhs slab")
    K = concat_slabs(slabs_e, Kg.shape[1])
    assert np.array_equal(K.indptr, Kg.indptr) and np.array_equal(K.indices, Kg.indices)
    assert_values_close(K.data, Kg.data, VAL_TOL)
    P = concat_slabs(slabs_p, Pg.shape[1])
    assert np.array_equal(P.indptr, Pg.indptr) and np.array_equal(P.indices, Pg.indices)
    assert_values_close(P.data, Pg.data, VAL_TOL)
    # boundary conditions on slabs == boundary conditions on the global matrix, row range by row range
    import scipy.sparse as sp
    Kd = np.array(Kg.todense())
    Kd[fixed, :] = 0.0
    Kd[:, fixed] = 0.0
    Kd[fixed, fixed] = 1.0
    got = sp.vstack(bc_rows).toarray()
    assert np.max(np.abs(got - Kd)) <= VAL_TOL * np.max(np.abs(Kd))
    if not three_d:
        xdv = np.ones(10) / 10.0
        u = np.random.default_rng(5).random(X.shape[0]) - 0.4
        Jg, rg = orc.assemble_nlpoisson(X, conn, xdv, u)
        Hk, Hr = orc.assemble_helmholtz(X, conn, 0.05)
        sj, sh = [], []
        for r in range(size):
            part = partition_mesh(X, conn, r, size)
            gb, ge = part.owned_global_range
            mn = pf.NonlinearPoisson2D(X, conn, [0], None, q, basis, partition=part, node_ranges=ranges)
            J = mn.compute_jacobian(xdv, u)
            sj.append((J.indptr.copy(), J.indices.copy(), J.data.copy()))
            assert_values_close(mn.compute_rhs(xdv, u), rg[gb:ge], VAL_TOL, "residual slab")
            mh = pf.Helmholtz(0.05, X, conn, q, basis, partition=part, node_ranges=ranges)
            sh.append((mh.R.indptr.copy(), mh.R.indices.copy(), mh.R.data.copy()))
            xfield = np.random.default_rng(9).random(X.shape[0])
            assert_values_close(mh.compute_rhs(xfield), (Hr @ xfield)[gb:ge], 1e-13, "R x slab")
        J = concat_slabs(sj, Jg.shape[1])
        assert np.array_equal(J.indices, Jg.indices)
        assert_values_close(J.data, Jg.data, VAL_TOL)
        R = concat_slabs(sh, Hr.shape[1])
        assert np.array_equal(R.indices, Hr.indices)
        assert_values_close(R.data, Hr.data, VAL_TOL)


def test_host_buffers_are_reused_but_never_aliased(pf):
    """compute_jacobian lands in pinned host buffers of the handle's pool: a buffer is reused only once the matrix
    built on it is gone, so two live matrices never share values; the cached pattern is shared read-only and copied
    on write by apply_dirichlet_bcs."""
    c = pf.ProblemCreator(25, 19)
    conn, X, dof_fixed, force = c.create_linear_elasticity_problem()
    q = pf.QuadratureBilinear2D()
    model = pf.LinearElasticity(X, conn, dof_fixed, None, force, q, pf.BasisBilinear2D(q))
    K1 = model.compute_jacobian(1.0)
    K2 = model.compute_jacobian(2.0)
    assert not np.shares_memory(K1.data, K2.data)
    assert np.array_equal(2.0 * K1.data, K2.data)
    addr = K1.data.__array_interface__["data"][0]
    ref = K1.data.copy()
    del K1
    K3 = model.compute_jacobian(1.0)  # K1's buffer is free again
    assert K3.data.__array_interface__["data"][0] == addr and np.array_equal(K3.data, ref)
    assert np.shares_memory(K3.indices, K2.indices) and not K3.indices.flags.writeable
    with pytest.raises(ValueError):
        K3.eliminate_zeros()  # an in-place pattern edit on the shared arrays fails loudly
    nnz = K2.nnz
    model.apply_dirichlet_bcs(K3, model.compute_rhs())  # copy-on-write, then eliminate_zeros
    assert K3.nnz < nnz and K2.nnz == nnz and K2.indices.flags.writeable is False
    K4 = model.compute_jacobian(1.0)
    assert K4.nnz == nnz and np.array_equal(K4.indices, K2.indices)


# ---- complex nodal density: the reference's complex-step checks run against the drop-in ------------------------------
@pytest.mark.parametrize("kind", ["quad", "hex"])
def test_complex_step_derivative_like_reference_tests(pf, kind):
    """tests/test_linear_poisson.py:57-89 and tests/test_elasticity.py:68-104 (run_dKdx): K(rho + 1j h p) assembled with
    a complex density, phi^T K psi read off the imaginary part, against _compute_K_dv_sens . p -- same seeds, same
    h = 1e-30, same 1e-12 bound.  The complex matrix itself is checked against the oracle evaluated in complex
    arithmetic (pattern bit-exact, both parts within 1e-12)."""
    creator = pf.ProblemCreator(nnodes_x=13, nnodes_y=11) if kind == "quad" else pf.ProblemCreator(7, 7, 7)
    q, b = _objs(pf, 4 if kind == "quad" else 8)
    h = 1e-30
    conn, X, dof_fixed = creator.create_poisson_problem()
    model = pf.LinearPoisson(X, conn, dof_fixed, None, q, b, gfunc, p=5.0)
    np.random.seed(0)
    nn = X.shape[0]
    phi, psi = np.random.rand(nn), np.random.rand(nn)
    rho, pert = np.random.rand(nn), np.random.rand(nn)
    dfdrho = pert.dot(model._compute_K_dv_sens(rho, phi, psi))
    K = model.compute_jacobian(rho + 1j * pert * h)
    assert K.dtype == np.complex128
    dfdrho_cs = phi.dot(K.dot(psi)).imag / h
    assert abs((dfdrho - dfdrho_cs) / dfdrho) <= 1e-12
    Kref = orc.assemble_poisson(X, conn, rho + 1j * pert * 0.25, 5.0)  # a finite imaginary part for the value check
    K = model.compute_jacobian(rho + 1j * pert * 0.25)
    assert_pattern_equal(K.indptr, K.indices, Kref.indptr, Kref.indices)
    assert_values_close(K.data.real, Kref.data.real, VAL_TOL, "Re K")
    assert_values_close(K.data.imag, Kref.data.imag, VAL_TOL, "Im K")
    # a real density through the same call stays real and takes the gather path
    assert model.compute_jacobian(rho).dtype == np.float64

    conn, X, dof_fixed, nodal_force = creator.create_linear_elasticity_problem()
    model = pf.LinearElasticity(X, conn, dof_fixed, None, nodal_force, q, b, p=5.0)
    np.random.seed(0)
    ndof = X.shape[0] * X.shape[1]
    phi, psi = np.random.rand(ndof), np.random.rand(ndof)
    rho, pert = np.random.rand(nn), np.random.rand(nn)
    dfdrho = pert.dot(model._compute_K_dv_sens(rho, phi, psi))
    K = model.compute_jacobian(rho + 1j * pert * h)
    dfdrho_cs = phi.dot(K.dot(psi)).imag / h
    assert abs((dfdrho - dfdrho_cs) / dfdrho) <= 1e-12
    Kref = orc.assemble_elasticity(X, conn, rho + 1j * pert * 0.25, 5.0)
    K = model.compute_jacobian(rho + 1j * pert * 0.25)
    assert_pattern_equal(K.indptr, K.indices, Kref.indptr, Kref.indices)
    assert_values_close(K.data.real, Kref.data.real, VAL_TOL, "Re K")
    assert_values_close(K.data.imag, Kref.data.imag, VAL_TOL, "Im K")
    # a complex scalar is a constant complex field (pyfem.py:1015-1016)
    Kc = model.compute_jacobian(0.7 + 0.1j)
    Kcr = orc.assemble_elasticity(X, conn, np.full(nn, 0.7 + 0.1j), 5.0)
    assert_values_close(Kc.data.imag, Kcr.data.imag, VAL_TOL, "constant complex rho")
