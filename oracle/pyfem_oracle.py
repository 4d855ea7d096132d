"""
CPU oracle for the pyfem assembly hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a numpy/scipy restatement of the one path BASELINE.json names:
per-element quadrature -> element matrices / residuals -> scatter-add into a
global CSR matrix and RHS vector.  It exists only so that tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs have an independent checker and
a CPU arm to time.  Nothing under pyfem_gpu_testflight_b200/ may import it.

Parity status: PINNED.  `oracle/make_golden.py` imports the unmodified
reference (/root/reference/pyfem.py + utils.py, with matplotlib/pyamg stubbed)
in the build container, runs every physics on small jittered meshes, and
commits inputs + the reference's own outputs under tests/golden/*.npz;
tests/test_oracle_golden.py checks this restatement against those files
(pattern bit-exact, values <= 1e-13 norm-relative).

Each function cites the reference lines (relative to /root/reference) that it
restates.  The third-party arithmetic on the path (numpy einsum / linalg.det /
add.at and scipy.sparse coo->csr; unpinned in the reference, numpy 2.3.5 and
scipy 1.18.1 here) is used as-is, exactly as the reference uses it.
"""
import numpy as np
from scipy import sparse, special

INV_SQRT3 = 1.0 / np.sqrt(3.0)


# --------------------------------------------------------------------------
# Quadrature + basis tables  (pyfem.py:83-112 quadrature, :253-338 basis)
# --------------------------------------------------------------------------
def quad4_tables():
    """2x2 Gauss rule in CCW order, unit weights (pyfem.py:87-92); bilinear
    shape functions and their reference-space derivatives (pyfem.py:263-283).
    Returns pts (4,2), w (4,), N (4,4) [q,node], Nderiv (4,4,2) [q,node,ref-axis]."""
    g = INV_SQRT3
    pts = np.array([[-g, -g], [g, -g], [g, g], [-g, g]])
    w = np.ones(4)
    sx = np.array([-1.0, 1.0, 1.0, -1.0])  # xi sign of local nodes 0..3
    sy = np.array([-1.0, -1.0, 1.0, 1.0])
    N = np.zeros((4, 4))
    dN = np.zeros((4, 4, 2))
    for q, (xi, eta) in enumerate(pts):
        for a in range(4):
            fx = (1.0 + xi) if sx[a] > 0 else (1.0 - xi)
            fy = (1.0 + eta) if sy[a] > 0 else (1.0 - eta)
            N[q, a] = 0.25 * fx * fy
            dN[q, a, 0] = sx[a] * 0.25 * fy
            dN[q, a, 1] = sy[a] * 0.25 * fx
    return pts, w, N, dN


def hex8_tables():
    """2x2x2 Gauss rule, x slowest / z fastest, unit weights (pyfem.py:101-110);
    trilinear shape functions and derivatives (pyfem.py:297-337)."""
    g = INV_SQRT3
    pts = np.array([[sxq * g, syq * g, szq * g]
                    for sxq in (-1, 1) for syq in (-1, 1) for szq in (-1, 1)], dtype=float)
    w = np.ones(8)
    sx = np.array([-1.0, 1.0, 1.0, -1.0, -1.0, 1.0, 1.0, -1.0])
    sy = np.array([-1.0, -1.0, 1.0, 1.0, -1.0, -1.0, 1.0, 1.0])
    sz = np.array([-1.0, -1.0, -1.0, -1.0, 1.0, 1.0, 1.0, 1.0])
    N = np.zeros((8, 8))
    dN = np.zeros((8, 8, 3))
    for q, (xi, eta, zeta) in enumerate(pts):
        for a in range(8):
            fx = (1.0 + xi) if sx[a] > 0 else (1.0 - xi)
            fy = (1.0 + eta) if sy[a] > 0 else (1.0 - eta)
            fz = (1.0 + zeta) if sz[a] > 0 else (1.0 - zeta)
            N[q, a] = 0.125 * fx * fy * fz
            dN[q, a, 0] = sx[a] * 0.125 * fy * fz
            dN[q, a, 1] = sy[a] * 0.125 * fx * fz
            dN[q, a, 2] = sz[a] * 0.125 * fx * fy
    return pts, w, N, dN


def tables_for(nnodes_per_elem):
    if nnodes_per_elem == 4:
        return quad4_tables()
    if nnodes_per_elem == 8:
        return hex8_tables()
    raise ValueError("oracle covers quad4 and hex8 only")


# --------------------------------------------------------------------------
# DOF numbering and COO pattern  (utils.py:267-298, pyfem.py:837-858)
# --------------------------------------------------------------------------
def conn_to_dof(conn, m):
    """conn_dof[e, a*m + axis] = m*conn[e,a] + axis (utils.py:293-296);
    for m == 1 it is conn itself (utils.py:286-289)."""
    conn = np.asarray(conn, dtype=np.int64)
    if m == 1:
        return conn
    E, n = conn.shape
    out = np.empty((E, n * m), dtype=np.int64)
    for axis in range(m):
        out[:, axis::m] = m * conn + axis
    return out


def coo_pattern(conn_dof):
    """nz_i / nz_j: every (row, col) pair of an element's D dofs, element-major,
    row-major inside the element (pyfem.py:846-858)."""
    E, D = conn_dof.shape
    rows = np.repeat(conn_dof, D, axis=1).reshape(-1)
    cols = np.tile(conn_dof, (1, D)).reshape(-1)
    return rows, cols


# --------------------------------------------------------------------------
# Geometry  (utils.py:154-264)
# --------------------------------------------------------------------------
def geometry(X, conn, dN):
    """Xe = X[conn] (utils.py:167); J[i,q,j,k] = sum_l dN[q,l,k] Xe[i,l,j]
    (utils.py:184); detJ by LAPACK det (utils.py:199); closed-form inverse with
    entry-wise division by detJ (utils.py:243-260); Ngrad = dN . invJ (utils.py:263)."""
    Xe = np.asarray(X, dtype=float)[conn]
    J = np.einsum("qlk,ilj->iqjk", dN, Xe)
    detJ = np.linalg.det(J)
    d = J.shape[-1]
    inv = np.empty_like(J)
    if d == 2:
        inv[..., 0, 0] = J[..., 1, 1] / detJ
        inv[..., 0, 1] = -J[..., 0, 1] / detJ
        inv[..., 1, 0] = -J[..., 1, 0] / detJ
        inv[..., 1, 1] = J[..., 0, 0] / detJ
    else:
        def cof(r0, c0, r1, c1):
            return J[..., r0, c0] * J[..., r1, c1]
        inv[..., 0, 0] = (cof(1, 1, 2, 2) - cof(1, 2, 2, 1)) / detJ
        inv[..., 0, 1] = -(cof(0, 1, 2, 2) - cof(0, 2, 2, 1)) / detJ
        inv[..., 0, 2] = (cof(0, 1, 1, 2) - cof(0, 2, 1, 1)) / detJ
        inv[..., 1, 0] = -(cof(1, 0, 2, 2) - cof(1, 2, 2, 0)) / detJ
        inv[..., 1, 1] = (cof(0, 0, 2, 2) - cof(0, 2, 2, 0)) / detJ
        inv[..., 1, 2] = -(cof(0, 0, 1, 2) - cof(0, 2, 1, 0)) / detJ
        inv[..., 2, 0] = (cof(1, 0, 2, 1) - cof(1, 1, 2, 0)) / detJ
        inv[..., 2, 1] = -(cof(0, 0, 2, 1) - cof(0, 1, 2, 0)) / detJ
        inv[..., 2, 2] = (cof(0, 0, 1, 1) - cof(0, 1, 1, 0)) / detJ
    Ngrad = np.einsum("jkm,ijml->ijkl", dN, inv)
    return Xe, J, detJ, Ngrad


def to_quad(N, data_e):
    """node -> quadrature interpolation (utils.py:218-220)."""
    if data_e.ndim == 2:
        return np.einsum("jl,il->ij", N, data_e)
    return np.einsum("jl,ilk->ijk", N, data_e)


def ramp(rho, conn, N, p, nnodes):
    """rho scalar -> constant nodal field (pyfem.py:1015-1016, 1780-1781);
    rho_q = N rho[conn]; c_q = rho_q / (1 + p (1 - rho_q)) (pyfem.py:1294-1300,
    1938-1944)."""
    if not hasattr(rho, "__len__"):
        rho = np.ones(nnodes) * rho
    rho_q = to_quad(N, np.asarray(rho)[conn])
    return rho_q / (1.0 + p * (1.0 - rho_q))


def ramp_deriv(rho, conn, N, p):
    """d/d rho_q of the RAMP factor, (1 + p) / (1 + p (1 - rho_q))^2 (pyfem.py:1319-1325)."""
    rho_q = to_quad(N, np.asarray(rho)[conn])
    return (1.0 + p) / (1.0 + p * (1.0 - rho_q)) ** 2


# --------------------------------------------------------------------------
# Element matrices / vectors
# --------------------------------------------------------------------------
def poisson_Ke(X, conn, rho=1.0, p=0.0):
    """pyfem.py:1188-1217 with the einsum of :1177-1185."""
    _, w, N, dN = tables_for(conn.shape[1])
    _, _, detJ, Ngrad = geometry(X, conn, dN)
    kq = ramp(rho, conn, N, p, X.shape[0])
    return np.einsum("iq,iq,q,iqjl,iqkl->ijk", kq, detJ, w, Ngrad, Ngrad, optimize=True)


def poisson_rhs_e(X, conn, gfunc):
    """pyfem.py:1137-1173: fe[i,j] = sum_q detJ w N[q,j] g(x_q) (einsum :1132-1134)."""
    _, w, N, dN = tables_for(conn.shape[1])
    Xe, _, detJ, _ = geometry(X, conn, dN)
    Xq = to_quad(N, Xe)
    g = np.zeros(Xq.shape[:-1])
    g[...] = gfunc(Xq)
    return np.einsum("ik,k,jk,ik->ij", detJ, w, N, g, optimize=True)


def elasticity_C0(ndims, E=10.0, nu=0.3):
    """pyfem.py:1746-1757."""
    if ndims == 2:
        C0 = E * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, 0.5 * (1.0 - nu)]])
        C0 *= 1.0 / (1.0 - nu ** 2)
        return C0
    C0 = np.zeros((6, 6))
    for i in range(3):
        for j in range(3):
            C0[i, j] = (1 - nu) if i == j else nu
        C0[3 + i, 3 + i] = 0.5 - nu
    C0 *= E / ((1 + nu) * (1 - 2 * nu))
    return C0


def strain_displacement(Ngrad):
    """B matrix (pyfem.py:1988-2011): 2-D rows [ex, ey, gxy]; 3-D rows
    [ex, ey, ez, gxy, gyz, gxz]; dof order interleaved (node, axis)."""
    E, Q, n, d = Ngrad.shape
    s = 3 if d == 2 else 6
    B = np.zeros((E, Q, s, n * d))
    gx, gy = Ngrad[..., 0], Ngrad[..., 1]
    if d == 2:
        B[:, :, 0, 0::2] = gx
        B[:, :, 1, 1::2] = gy
        B[:, :, 2, 0::2] = gy
        B[:, :, 2, 1::2] = gx
    else:
        gz = Ngrad[..., 2]
        B[:, :, 0, 0::3] = gx
        B[:, :, 1, 1::3] = gy
        B[:, :, 2, 2::3] = gz
        B[:, :, 3, 0::3] = gy
        B[:, :, 3, 1::3] = gx
        B[:, :, 4, 1::3] = gz
        B[:, :, 4, 2::3] = gy
        B[:, :, 5, 0::3] = gz
        B[:, :, 5, 2::3] = gx
    return B


def elasticity_Ke(X, conn, rho=1.0, p=0.0, E=10.0, nu=0.3):
    """pyfem.py:2029-2068 with the einsum of :2017-2026."""
    _, w, N, dN = tables_for(conn.shape[1])
    _, _, detJ, Ngrad = geometry(X, conn, dN)
    Cq = ramp(rho, conn, N, p, X.shape[0])
    B = strain_displacement(Ngrad)
    C0 = elasticity_C0(X.shape[1], E, nu)
    return np.einsum("iq,q,iqnj,iq,nm,iqmk->ijk", detJ, w, B, Cq, C0, B, optimize=True)


def helmholtz_KeRe(X, conn, r0):
    """pyfem.py:2138-2177: Re = sum detJ w N N^T (:2134-2136);
    Ke = r0^2 sum detJ w gradN.gradN (:2127-2130) + Re (:2176)."""
    _, w, N, dN = tables_for(conn.shape[1])
    _, _, detJ, Ngrad = geometry(X, conn, dN)
    Re = np.einsum("iq,q,qj,qk->ijk", detJ, w, N, N, optimize=True)
    Ke = np.einsum("iq,q,iqjl,iqkl->ijk", detJ * r0 ** 2, w, Ngrad, Ngrad, optimize=True)
    Ke += Re
    return Ke, Re


def nlpoisson_h(xdv, Xq):
    """Bernstein-weighted coefficient field (pyfem.py:1450-1472)."""
    x, y = Xq[..., 0], Xq[..., 1]
    M = np.shape(xdv)[0]
    h = np.zeros(x.shape)
    for k in range(M):
        coef = special.binom(M - 1, k)
        h += xdv[k] * (coef * (1.0 - x) ** (M - 1 - k) * x ** k) * (4.0 * y * (1.0 - y))
    return h + 1.0


def nlpoisson_g(Xq):
    """Source term (pyfem.py:1427-1448)."""
    x, y = Xq[..., 0], Xq[..., 1]
    return 1e4 * x * (1.0 - x) * (1.0 - 2.0 * x) * y * (1.0 - y) * (1.0 - 2.0 * y)


def nlpoisson_Ke(X, conn, xdv, u):
    """Newton Jacobian, non-symmetric (pyfem.py:1541-1610; einsums :1595-1609)."""
    _, w, N, dN = quad4_tables()
    Xe, _, detJ, Ngrad = geometry(X, conn, dN)
    Xq = to_quad(N, Xe)
    ue = np.asarray(u, dtype=float)[conn]
    uq = to_quad(N, ue)
    h = nlpoisson_h(xdv, Xq)
    Ke = np.einsum("nq,q,nqjl,nqkl->njk", detJ * h * (1.0 + uq ** 2), w, Ngrad, Ngrad)
    Ke += np.einsum("nq,nqjl,nqkl,nk,qi->nji", 2.0 * detJ * h * uq * w, Ngrad, Ngrad, ue, N)
    return Ke


def nlpoisson_res_e(X, conn, xdv, u):
    """Element residual (pyfem.py:1474-1539; einsum :1530-1537)."""
    _, w, N, dN = quad4_tables()
    Xe, _, detJ, Ngrad = geometry(X, conn, dN)
    Xq = to_quad(N, Xe)
    ue = np.asarray(u, dtype=float)[conn]
    uq = to_quad(N, ue)
    h = nlpoisson_h(xdv, Xq)
    g = nlpoisson_g(Xq)
    r = np.einsum("nq,nqjl,nqkl,nk->nj", detJ * h * (1.0 + uq ** 2) * w, Ngrad, Ngrad, ue)
    r -= np.dot(detJ * w * g, N)
    return r


# --------------------------------------------------------------------------
# Global scatter  (pyfem.py:860-875 vector, :920-931 matrix)
# --------------------------------------------------------------------------
def scatter_matrix(Ke, conn_dof):
    """coo_matrix((Ke.flatten(), (nz_i, nz_j))).tocsr() -- duplicates summed,
    explicit zeros kept, shape inferred from max index (pyfem.py:930-931)."""
    rows, cols = coo_pattern(conn_dof)
    return sparse.coo_matrix((Ke.reshape(-1), (rows, cols))).tocsr()


def scatter_vector(fe, conn_dof, ndof, nquads):
    """np.add.at per local column, looping over range(nquads) exactly as the
    reference does (pyfem.py:872-874; survey trap T3: only right when Q == D)."""
    out = np.zeros(ndof)
    for c in range(nquads):
        np.add.at(out, conn_dof[:, c], fe[:, c])
    return out


# --------------------------------------------------------------------------
# Whole-path entry points (what compute_jacobian / compute_rhs return)
# --------------------------------------------------------------------------
def assemble_poisson(X, conn, rho=1.0, p=0.0):
    conn = np.asarray(conn, dtype=np.int64)
    return scatter_matrix(poisson_Ke(X, conn, rho, p), conn)


def assemble_poisson_rhs(X, conn, gfunc):
    conn = np.asarray(conn, dtype=np.int64)
    return scatter_vector(poisson_rhs_e(X, conn, gfunc), conn, X.shape[0], conn.shape[1])


def assemble_elasticity(X, conn, rho=1.0, p=0.0, E=10.0, nu=0.3):
    conn = np.asarray(conn, dtype=np.int64)
    return scatter_matrix(elasticity_Ke(X, conn, rho, p, E, nu), conn_to_dof(conn, X.shape[1]))


def assemble_helmholtz(X, conn, r0):
    conn = np.asarray(conn, dtype=np.int64)
    Ke, Re = helmholtz_KeRe(X, conn, r0)
    return scatter_matrix(Ke, conn), scatter_matrix(Re, conn)


def assemble_nlpoisson(X, conn, xdv, u):
    conn = np.asarray(conn, dtype=np.int64)
    K = scatter_matrix(nlpoisson_Ke(X, conn, xdv, u), conn)
    res = scatter_vector(nlpoisson_res_e(X, conn, xdv, u), conn, X.shape[0], 4)
    return K, res


# --------------------------------------------------------------------------
# Sensitivities d(phi^T K psi) / d rho  (the step after assembly in the topology-optimisation loop)
# --------------------------------------------------------------------------
def _scatter_nodal(inner, conn, nnodes):
    """np.add.at over the element's local nodes (pyfem.py:1272-1275, 1916-1919)."""
    out = np.zeros(nnodes)
    for i in range(conn.shape[1]):
        np.add.at(out, conn[:, i], inner[:, i])
    return out


def poisson_K_dv_sens(X, conn, rho, p, phi, psi):
    """LinearPoisson._compute_K_dv_sens (pyfem.py:1239-1276): Ke_deriv by the einsum of :1220-1230 with
    kappa_q_deriv = N[q,o] ramp'(rho_q) (:1325-1328), inner product :1234-1236, nodal scatter :1272-1275."""
    _, w, N, dN = tables_for(conn.shape[1])
    _, _, detJ, Ngrad = geometry(X, conn, dN)
    dk = np.einsum("ql,iq->iql", N, ramp_deriv(rho, conn, N, p))
    Ke_deriv = np.einsum("iqo,iq,q,iqjl,iqkl->ijko", dk, detJ, w, Ngrad, Ngrad, optimize=True)
    inner = np.einsum("ij,ik,ijko->io", np.asarray(phi)[conn], np.asarray(psi)[conn], Ke_deriv)
    return _scatter_nodal(inner, conn, X.shape[0])


def elasticity_K_dv_sens(X, conn, rho, p, phi, psi, E=10.0, nu=0.3):
    """LinearElasticity._compute_K_dv_sens (pyfem.py:1872-1920): einsum :1900-1909, inner product :1912-1914."""
    _, w, N, dN = tables_for(conn.shape[1])
    _, _, detJ, Ngrad = geometry(X, conn, dN)
    B = strain_displacement(Ngrad)
    C0 = elasticity_C0(X.shape[1], E, nu)
    dC = np.einsum("ql,iq->iql", N, ramp_deriv(rho, conn, N, p))
    Ke_deriv = np.einsum("iq,q,iqnj,iqo,nm,iqmk->ijko", detJ, w, B, dC, C0, B, optimize=True)
    cd = conn_to_dof(conn, X.shape[1])
    inner = np.einsum("ij,ik,ijko->io", np.asarray(phi)[cd], np.asarray(psi)[cd], Ke_deriv)
    return _scatter_nodal(inner, conn, X.shape[0])


def elasticity_point_loads(ndof, ndims, nodal_force):
    """rhs[m*node + axis] = force (assignment; pyfem.py:1765-1767)."""
    rhs = np.zeros(ndof)
    nodes = np.array(list(nodal_force.keys()), dtype=np.int64)
    vals = np.array(list(nodal_force.values()), dtype=float)
    dofs = (ndims * nodes[:, None] + np.arange(ndims)[None, :]).reshape(-1)
    rhs[dofs] = vals.reshape(-1)
    return rhs


# --------------------------------------------------------------------------
# Synthetic structured meshes identical to ProblemCreator (pyfem.py:2469-2535)
# --------------------------------------------------------------------------
def structured_mesh(nx, ny, nz=None, Lx=None, Ly=None, Lz=None):
    """Vectorised equivalent of ProblemCreator's loops: node id = i + j*nx + k*nx*ny,
    quad conn (pyfem.py:2499-2502), hex conn (pyfem.py:2527-2534)."""
    three_d = nz is not None
    nzz = nz if three_d else 1
    Lx = (nx - 1) / (ny - 1) if Lx is None else Lx
    Ly = 1.0 if Ly is None else Ly
    Lz = (nzz - 1) / (ny - 1) if Lz is None else Lz
    x = np.linspace(0, Lx, nx)
    y = np.linspace(0, Ly, ny)
    z = np.linspace(0, Lz, nzz)
    Z, Y, Xg = np.meshgrid(z, y, x, indexing="ij")
    X = np.stack([Xg.ravel(), Y.ravel(), Z.ravel()], axis=1)
    ids = np.arange(nx * ny * nzz, dtype=np.int64).reshape(nzz, ny, nx)
    if not three_d:
        c = ids[0]
        conn = np.stack([c[:-1, :-1], c[:-1, 1:], c[1:, 1:], c[1:, :-1]], axis=-1).reshape(-1, 4)
        return np.ascontiguousarray(X[:, :2]), conn
    lo, hi = ids[:-1], ids[1:]
    conn = np.stack([lo[:, :-1, :-1], lo[:, :-1, 1:], lo[:, 1:, 1:], lo[:, 1:, :-1],
                     hi[:, :-1, :-1], hi[:, :-1, 1:], hi[:, 1:, 1:], hi[:, 1:, :-1]],
                    axis=-1).reshape(-1, 8)
    return X, conn
