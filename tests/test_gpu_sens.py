"""GPU parity of the fused sensitivity kernel d(phi^T K psi)/d rho (the reference's _compute_K_dv_sens, pyfem.py:1239-1276
and 1872-1920) and of the compliance functions built on it: against the reference's own outputs (tests/golden/sens_*,
written by oracle/make_golden.py) and against the numpy oracle on larger seeded meshes."""
import numpy as np
import pytest

import pyfem_oracle as orc
from make_golden import jitter, test_gfunc as gfunc
from parity import VAL_TOL, assert_values_close, golden_files

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pf():
    import pyfem_gpu_testflight_b200 as pf
    return pf


def _objs(pf, nne):
    q = pf.QuadratureBilinear2D() if nne == 4 else pf.QuadratureBlock3D()
    return q, (pf.BasisBilinear2D(q) if nne == 4 else pf.BasisBlock3D(q))


@pytest.mark.parametrize("path", golden_files("sens"))
def test_sensitivities_golden(pf, path):
    g = np.load(path)
    X, conn = g["X"], g["conn"]
    q, b = _objs(pf, conn.shape[1])
    m = pf.LinearPoisson(X, conn, [0], None, q, b, gfunc, p=float(g["p_poisson"]))
    assert_values_close(m._compute_K_dv_sens(g["rho"], g["phi"], g["psi"]), g["g_poisson"], VAL_TOL, "poisson")
    m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0] * X.shape[1]}, q, b, E=float(g["E"]), nu=float(g["nu"]),
                            p=float(g["p_elast"]))
    assert_values_close(m._compute_K_dv_sens(g["rho"], g["phi_v"], g["psi_v"]), g["g_elast"], VAL_TOL, "elasticity")


@pytest.mark.parametrize("three_d", [False, True])
def test_sensitivities_vs_oracle(pf, three_d):
    if three_d:
        X, conn = orc.structured_mesh(13, 9, 11)
        X = jitter(X, (13, 9, 11), seed=6)
    else:
        X, conn = orc.structured_mesh(83, 61)
        X = jitter(X, (83, 61), seed=6)
    conn = conn[np.random.default_rng(2).permutation(conn.shape[0])]
    nn, d = X.shape
    rng = np.random.default_rng(8)
    rho = 0.05 + 0.95 * rng.random(nn)
    q, b = _objs(pf, conn.shape[1])
    phi, psi = rng.random(nn) - 0.5, rng.random(nn) - 0.5
    m = pf.LinearPoisson(X, conn, [0], None, q, b, gfunc, p=2.0)
    assert_values_close(m._compute_K_dv_sens(rho, phi, psi), orc.poisson_K_dv_sens(X, conn, rho, 2.0, phi, psi), VAL_TOL)
    assert_values_close(m._compute_K_dv_sens(0.7, phi, psi),
                        orc.poisson_K_dv_sens(X, conn, np.full(nn, 0.7), 2.0, phi, psi), VAL_TOL, "constant rho")
    phi, psi = rng.random(nn * d) - 0.5, rng.random(nn * d) - 0.5
    m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0] * d}, q, b, E=3.0, nu=0.27, p=4.0)
    assert_values_close(m._compute_K_dv_sens(rho, phi, psi),
                        orc.elasticity_K_dv_sens(X, conn, rho, 4.0, phi, psi, 3.0, 0.27), VAL_TOL)


def test_compliance_gradient_matches_finite_difference(pf):
    """compliance / compliance_grad as the reference's tests use them (tests/test_elasticity.py:68-104 checks the
    same derivative by complex step; without a complex device path a central difference stands in)."""
    c = pf.ProblemCreator(17, 13)
    conn, X, dof_fixed, force = c.create_linear_elasticity_problem()
    q, b = _objs(pf, 4)
    m = pf.LinearElasticity(X, conn, dof_fixed, None, force, q, b, p=5.0)
    rng = np.random.default_rng(0)
    rho = 0.3 + 0.6 * rng.random(X.shape[0])
    comp, u = m.compliance(rho, solver="direct")
    grad = m.compliance_grad(rho, u)
    pert = rng.random(X.shape[0]) - 0.5
    h = 1e-6
    cp, _ = m.compliance(rho + h * pert, solver="direct")
    cm, _ = m.compliance(rho - h * pert, solver="direct")
    fd = (cp - cm) / (2 * h)
    assert abs(fd - grad.dot(pert)) <= 1e-6 * abs(fd)
    assert m.volume(rho) == pytest.approx(rho.sum() / X.shape[0])
    assert np.allclose(m.volume_grad(rho), 1.0 / X.shape[0])
    # thermal compliance, weighted form (pyfem.py:1033-1101)
    conn, X, dof_fixed = c.create_poisson_problem()[:3]
    mp = pf.LinearPoisson(X, conn, dof_fixed, None, q, b, gfunc, p=3.0)
    comp, u = mp.compliance(rho, solver="direct")
    grad = mp.compliance_grad(rho, u)
    cp, _ = mp.compliance(rho + h * pert, solver="direct")
    cm, _ = mp.compliance(rho - h * pert, solver="direct")
    fd = (cp - cm) / (2 * h)
    assert abs(fd - grad.dot(pert)) <= 1e-6 * abs(fd)


@pytest.mark.parametrize("three_d", [False, True])
def test_sensitivities_on_row_slabs(pf, three_d):
    """Multi-GPU decomposition with the ranks emulated one after another: each rank's handle (element block + ghost
    layer, owned node range) yields the owned nodes' sensitivities; concatenated they are the global vector."""
    from pyfem_gpu_testflight_b200.partition import partition_mesh
    if three_d:
        X, conn = orc.structured_mesh(8, 7, 12)
        X = jitter(X, (8, 7, 12), seed=9)
    else:
        X, conn = orc.structured_mesh(37, 45)
        X = jitter(X, (37, 45), seed=9)
    nn, d = X.shape
    rng = np.random.default_rng(4)
    rho = 0.05 + 0.95 * rng.random(nn)
    phi, psi = rng.random(nn * d) - 0.5, rng.random(nn * d) - 0.5
    ref = orc.elasticity_K_dv_sens(X, conn, rho, 3.0, phi, psi, 8.0, 0.31)
    out = []
    for r in range(3):
        part = partition_mesh(X, conn, r, 3)
        mesh = pf.DeviceMesh(part.X, part.conn, d, own_range=part.own_range, node_gid=part.node_gid,
                             ncols_nodes=part.nnodes_global)
        dof = (part.node_gid[:, None] * d + np.arange(d)[None, :]).ravel()
        out.append(mesh.k_dv_sens("elasticity", rho[part.node_gid], 3.0, phi[dof], psi[dof], E=8.0, nu=0.31).cpu().numpy())
    assert_values_close(np.concatenate(out), ref, VAL_TOL)


@pytest.mark.parametrize("three_d", [False, True])
def test_sensitivity_paths_agree_and_tile_path_is_reproducible(pf, three_d):
    """Two device paths: the element-per-thread pass with atomic nodal adds (a handle with ndims dofs per node) and
    the tile-plan pass on a scalar handle (node-window staging of rho / phi / psi, plan-ordered nodal sums).  They
    agree to rounding; the tile pass repeats bit for bit; a row slab of the scalar handle yields the owned nodes."""
    import torch
    from pyfem_gpu_testflight_b200.partition import partition_mesh
    if three_d:
        X, conn = orc.structured_mesh(14, 11, 9)
    else:
        X, conn = orc.structured_mesh(131, 77)
    nn, d = X.shape
    rng = np.random.default_rng(11)
    rho = 0.05 + 0.95 * rng.random(nn)
    phi, psi = rng.random(nn * d) - 0.5, rng.random(nn * d) - 0.5
    ref = orc.elasticity_K_dv_sens(X, conn, rho, 3.0, phi, psi, 8.0, 0.31)
    vec, sca = pf.DeviceMesh(X, conn, d), pf.DeviceMesh(X, conn, 1)
    ga = vec.k_dv_sens("elasticity", rho, 3.0, phi, psi, E=8.0, nu=0.31)
    gt = sca.k_dv_sens("elasticity", rho, 3.0, phi, psi, E=8.0, nu=0.31, deterministic=True)
    assert torch.equal(gt, sca.k_dv_sens("elasticity", rho, 3.0, phi, psi, E=8.0, nu=0.31, deterministic=True))
    with pytest.raises(NotImplementedError):
        vec.k_dv_sens("elasticity", rho, 3.0, phi, psi, deterministic=True)  # needs the scalar handle's plan
    assert_values_close(ga.cpu().numpy(), ref, VAL_TOL, "atomic pass")
    assert_values_close(gt.cpu().numpy(), ref, VAL_TOL, "tile pass")
    gc = sca.k_dv_sens("elasticity", 0.6, 3.0, phi, psi, E=8.0, nu=0.31, deterministic=True)  # constant density
    assert_values_close(gc.cpu().numpy(), orc.elasticity_K_dv_sens(X, conn, np.full(nn, 0.6), 3.0, phi, psi, 8.0, 0.31),
                        VAL_TOL, "constant rho")
    out = []
    for r in range(3):
        part = partition_mesh(X, conn, r, 3)
        mesh = pf.DeviceMesh(part.X, part.conn, 1, own_range=part.own_range, node_gid=part.node_gid,
                             ncols_nodes=part.nnodes_global)
        dof = (part.node_gid[:, None] * d + np.arange(d)[None, :]).ravel()
        out.append(mesh.k_dv_sens("elasticity", rho[part.node_gid], 3.0, phi[dof], psi[dof], E=8.0, nu=0.31,
                                  deterministic=True).cpu().numpy())
    assert_values_close(np.concatenate(out), ref, VAL_TOL, "slabs")
    # through the model: deterministic_sens switches compliance_grad's kernel to the plan-ordered pass
    q, b = _objs(pf, conn.shape[1])
    model = pf.LinearElasticity(X, conn, [0], None, {0: [0.0] * d}, q, b, E=8.0, nu=0.31, p=3.0)
    g_atomic = model._compute_K_dv_sens(rho, phi, psi)
    model.deterministic_sens = True
    g_ordered = model._compute_K_dv_sens(rho, phi, psi)
    assert np.array_equal(g_ordered, model._compute_K_dv_sens(rho, phi, psi))
    assert_values_close(g_atomic, ref, VAL_TOL)
    assert_values_close(g_ordered, ref, VAL_TOL)
