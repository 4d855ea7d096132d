"""GPU parity on irregular / degenerate meshes, and size-independent properties at BASELINE's full sizes.

The oracle cannot assemble 16 M elements on the host, so the full-size checks use properties the
domain offers: rigid-body null space, symmetry, exact linearity in rho for p = 0, translation invariance of the
stencil, bitwise reproducibility, agreement of the two scatter strategies, and the pattern's closed-form size.
"""
import numpy as np
import pytest

import pyfem_oracle as orc
from make_golden import test_gfunc as gfunc
from parity import VAL_TOL, assert_csr_matches, assert_values_close

pytestmark = pytest.mark.gpu
MODES = ["atomic", "gather"]


@pytest.fixture(scope="module")
def pf():
    import pyfem_gpu_testflight_b200 as pf
    return pf


def _objs(pf):
    q = pf.QuadratureBilinear2D()
    return q, pf.BasisBilinear2D(q)


# ---- irregular quad meshes -------------------------------------------------------------------------
def three_patch_mesh(n, seed=0):
    """Three n x n quad patches glued around a centre node (a 'Y' block mesh): the centre has valence 3, the patch
    seams valence 4 with irregular neighbour sets, element numbering follows the patches (no global structure)."""
    ang = [np.pi / 2, np.pi / 2 + 2 * np.pi / 3, np.pi / 2 + 4 * np.pi / 3]
    d = [np.array([np.cos(a), np.sin(a)]) for a in ang]
    node_id, X, conn = {}, [], []

    def node(key, xy):
        if key not in node_id:
            node_id[key] = len(X)
            X.append(xy)
        return node_id[key]

    for p in range(3):
        e1, e2 = d[p], d[(p + 1) % 3]

        def key(i, j):
            # nodes on the seam (j == 0 of patch p) coincide with (i == 0) nodes of the previous patch
            if i == 0 and j == 0:
                return ("c",)
            if j == 0:
                return ("s", p, i)
            if i == 0:
                return ("s", (p + 1) % 3, j)
            return ("p", p, i, j)

        for j in range(n):
            for i in range(n):
                ids = []
                for (a, b) in ((i, j), (i + 1, j), (i + 1, j + 1), (i, j + 1)):
                    ids.append(node(key(a, b), (a * e1 + b * e2) / n))
                conn.append(ids)
    X = np.array(X)
    rng = np.random.default_rng(seed)
    X = X + rng.uniform(-0.15, 0.15, size=X.shape) / n
    X = (X - X.min(axis=0)) / (X.max(axis=0) - X.min(axis=0))  # nonlinear Poisson wants [0,1]^2
    conn = np.array(conn, dtype=np.int64)
    # make every element counter-clockwise
    x = X[conn]
    area = 0.5 * np.sum(x[:, :, 0] * np.roll(x[:, :, 1], -1, axis=1) - np.roll(x[:, :, 0], -1, axis=1) * x[:, :, 1], axis=1)
    conn[area < 0] = conn[area < 0][:, ::-1]
    return X, conn


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n", [3, 17])
def test_three_patch_mesh_all_physics(pf, mode, n):
    X, conn = three_patch_mesh(n, seed=n)
    conn = conn[np.random.default_rng(1).permutation(conn.shape[0])]
    q, b = _objs(pf)
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0, 0.0]}, q, b, p=4.0, scatter=mode)
    Kr = orc.assemble_elasticity(X, conn, rho, 4.0)
    assert_csr_matches(m.compute_jacobian(rho), Kr.indptr, Kr.indices, Kr.data)
    m = pf.LinearPoisson(X, conn, [0], None, q, b, gfunc, p=1.5, scatter=mode)
    Kr = orc.assemble_poisson(X, conn, rho, 1.5)
    assert_csr_matches(m.compute_jacobian(rho), Kr.indptr, Kr.indices, Kr.data)
    assert_values_close(m.compute_rhs(), orc.assemble_poisson_rhs(X, conn, gfunc), VAL_TOL, "rhs")
    m = pf.Helmholtz(0.1, X, conn, q, b, scatter=mode)
    Kr, Rr = orc.assemble_helmholtz(X, conn, 0.1)
    assert_csr_matches(m.K, Kr.indptr, Kr.indices, Kr.data)
    assert_csr_matches(m.R, Rr.indptr, Rr.indices, Rr.data)
    m = pf.NonlinearPoisson2D(X, conn, [0], None, q, b, scatter=mode)
    xdv = np.linspace(0.02, 0.2, 10)
    u = np.random.default_rng(5).random(X.shape[0]) - 0.4
    Kr, rr = orc.assemble_nlpoisson(X, conn, xdv, u)
    assert_csr_matches(m.compute_jacobian(xdv, u), Kr.indptr, Kr.indices, Kr.data)
    assert_values_close(m.compute_rhs(xdv, u), rr, VAL_TOL, "residual")


@pytest.mark.parametrize("mode", MODES)
def test_fan_mesh_high_valence(pf, mode):
    """Eight quads sharing one node (valence 8): more contributions per block than the unrolled fast path holds."""
    k = 8
    X = [[0.0, 0.0]]
    for i in range(2 * k):
        a = 2 * np.pi * i / (2 * k)
        r = 1.0 if i % 2 == 0 else 1.3
        X.append([r * np.cos(a), r * np.sin(a)])
    conn = [[0, 1 + 2 * i, 1 + (2 * i + 1) % (2 * k), 1 + (2 * i + 2) % (2 * k)] for i in range(k)]
    X, conn = np.array(X), np.array(conn, dtype=np.int64)
    q, b = _objs(pf)
    Kr = orc.assemble_elasticity(X, conn, 0.7, 2.0)
    m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0, 0.0]}, q, b, p=2.0, scatter=mode)
    assert_csr_matches(m.compute_jacobian(0.7), Kr.indptr, Kr.indices, Kr.data)
    Kr = orc.assemble_poisson(X, conn)
    m = pf.LinearPoisson(X, conn, [0], None, q, b, gfunc, scatter=mode)
    assert_csr_matches(m.compute_jacobian(), Kr.indptr, Kr.indices, Kr.data)


@pytest.mark.parametrize("mode", MODES)
def test_degenerate_sizes(pf, mode):
    """One element; a strip of elements; a mesh with a node no element references (empty CSR rows, as scipy gives)."""
    q, b = _objs(pf)
    for nx, ny in ((2, 2), (2, 9), (33, 2)):
        X, conn = orc.structured_mesh(nx, ny)
        Kr = orc.assemble_elasticity(X, conn)
        m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0, 0.0]}, q, b, scatter=mode)
        assert_csr_matches(m.compute_jacobian(), Kr.indptr, Kr.indices, Kr.data)
    X, conn = orc.structured_mesh(6, 5)
    # add an unreferenced node in the middle of the numbering: renumber nodes >= 7 up by one
    X2 = np.insert(X, 7, [[0.33, 0.77]], axis=0)
    conn2 = np.where(conn >= 7, conn + 1, conn)
    Kr = orc.assemble_poisson(X2, conn2)
    m = pf.LinearPoisson(X2, conn2, [0], None, q, b, gfunc, scatter=mode)
    K = m.compute_jacobian()
    assert_csr_matches(K, Kr.indptr, Kr.indices, Kr.data)
    assert K.indptr[7] == K.indptr[8]  # the unreferenced node's row is empty
    Kr = orc.assemble_elasticity(X2, conn2)
    m = pf.LinearElasticity(X2, conn2, [0], None, {0: [0.0, 0.0]}, q, b, scatter=mode)
    assert_csr_matches(m.compute_jacobian(), Kr.indptr, Kr.indices, Kr.data)


# ---- full-size properties (BASELINE configs[1] and [3]) ---------------------------------------------
@pytest.fixture(scope="module")
def big_quad(pf):
    import torch
    n = 4096
    c = pf.ProblemCreator(n + 1, n + 1)
    mesh = pf.DeviceMesh(c.X, c.conn, 2)
    torch.cuda.synchronize()
    return n, c, mesh


def test_full_size_elasticity_properties(pf, big_quad):
    import torch
    n, c, mesh = big_quad
    nn = (n + 1) ** 2
    assert mesh.nelems == n * n == 16777216
    assert mesh.nnz == 4 * (3 * (n + 1) - 2) ** 2 == 604078084  # closed form m^2 * prod(3 nn_k - 2)
    assert mesh.idx_bytes == 4                                   # scipy's rule: COO nnz = 2^30 fits int32
    K = mesh.assemble_elasticity(1.0, 0.0, mode="gather")
    # bitwise reproducible; exactly linear in a constant rho when p = 0 (scaling by a power of two is exact)
    assert torch.equal(K, mesh.assemble_elasticity(1.0, 0.0, mode="gather"))
    assert torch.equal(2.0 * K, mesh.assemble_elasticity(2.0, 0.0, mode="gather"))
    scale = float(K.abs().max())
    # the two scatter strategies agree
    Ka = mesh.assemble_elasticity(1.0, 0.0, mode="atomic")
    assert float((Ka - K).abs().max()) <= 1e-13 * scale
    del Ka
    # rigid-body translations are in the null space: K [1,0,1,0,...] = K [0,1,0,1,...] = 0
    g = torch.Generator(device="cuda").manual_seed(1)
    for axis in (0, 1):
        t = torch.zeros(2 * nn, dtype=torch.float64, device="cuda")
        t[axis::2] = 1.0
        assert float(mesh.spmv(K, t).abs().max()) <= 1e-12 * scale
    # symmetry: x' K y == y' K x
    x = torch.rand(2 * nn, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.rand(2 * nn, dtype=torch.float64, device="cuda", generator=g) - 0.5
    a, b = float(x @ mesh.spmv(K, y)), float(y @ mesh.spmv(K, x))
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b), scale)
    # translation invariance: interior nodes of the uniform mesh carry the same 2 x 18 row values
    indptr, _ = mesh.pattern()
    rows = [2 * (j * (n + 1) + i) for (i, j) in ((7, 9), (2048, 2048), (4000, 123))]
    ref = K[int(indptr[rows[0]]):int(indptr[rows[0] + 2])]
    for r in rows[1:]:
        got = K[int(indptr[r]):int(indptr[r + 2])]
        assert float((got - ref).abs().max()) <= 1e-12 * scale
    # the sub-mesh the oracle can afford: rows of a 33 x 33-node corner patch equal the oracle's rows
    Xs, cs = orc.structured_mesh(40, 40, Lx=39.0 / n * c.X[:, 0].max(), Ly=39.0 / n * c.X[:, 1].max())
    Ks = orc.assemble_elasticity(Xs, cs)
    for (i, j) in ((0, 0), (5, 0), (17, 30)):
        rs, rb = 2 * (j * 40 + i), 2 * (j * (n + 1) + i)
        got = K[int(indptr[rb]):int(indptr[rb + 1])].cpu().numpy()
        want = Ks.data[Ks.indptr[rs]:Ks.indptr[rs + 1]]
        assert_values_close(got, want, VAL_TOL, f"row of node ({i},{j})")


def test_full_size_nlpoisson_properties(pf):
    import torch
    n = 4096
    c = pf.ProblemCreator(n + 1, n + 1)
    X = c.X / c.X.max(axis=0)
    mesh = pf.DeviceMesh(X, c.conn, 1)
    nn = (n + 1) ** 2
    assert mesh.nnz == (3 * (n + 1) - 2) ** 2 == 151019521
    xdv = np.ones(10) / 10.0
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.rand(nn, dtype=torch.float64, device="cuda", generator=g)
    K, res = mesh.assemble_nlpoisson(xdv, u, mode="gather")
    K2, res2 = mesh.assemble_nlpoisson(xdv, u, mode="gather")
    assert torch.equal(K, K2) and torch.equal(res, res2)
    Ka, resa = mesh.assemble_nlpoisson(xdv, u, mode="atomic")
    assert float((Ka - K).abs().max()) <= 1e-13 * float(K.abs().max())
    assert float((resa - res).abs().max()) <= 1e-13 * float(res.abs().max())
    # the Jacobian is the derivative of the residual: res(u + eps v) - res(u - eps v) ~ 2 eps K v
    v = torch.rand(nn, dtype=torch.float64, device="cuda", generator=g) - 0.5
    eps = 1e-6
    _, rp = mesh.assemble_nlpoisson(xdv, u + eps * v, want_K=False, mode="gather")
    rp = rp.clone()
    _, rm = mesh.assemble_nlpoisson(xdv, u - eps * v, want_K=False, mode="gather")
    fd = (rp - rm) / (2 * eps)
    Kv = mesh.spmv(K, v)
    assert float((fd - Kv).abs().max()) <= 1e-6 * float(Kv.abs().max())


# ---- irregular hex meshes: the owner-computes chunk-row pass of hex8 elasticity and its fall-backs ------------------
def _hex_jittered(nx, ny, nz, seed):
    X, conn = orc.structured_mesh(nx, ny, nz)
    h = 1.0 / (max(nx, ny, nz) - 1)
    return X + np.random.default_rng(seed).uniform(-0.15 * h, 0.15 * h, size=X.shape), conn


def _rotate_local_numbering(conn, seed):
    """Renumber a random half of the elements by a quarter turn about the local zeta axis (still right-handed): two
    elements around a node may then hold it and a common neighbour at the same local positions, which the chunk-row
    pass does not cover (MeshDev::hex_rows_ok) -- AUTO must fall back to the atomic scatter, GATHER to the
    first-format gather kernel."""
    conn = conn.copy()
    pick = np.random.default_rng(seed).random(conn.shape[0]) < 0.5
    conn[pick] = conn[pick][:, [1, 2, 3, 0, 5, 6, 7, 4]]
    return conn


def extruded_three_patch_mesh(n, nz, seed=0):
    """The 'Y' block quad mesh extruded in z: the centre column has six elements per node instead of eight."""
    X2, quads = three_patch_mesh(n, seed=seed)
    nn2 = X2.shape[0]
    X = np.vstack([np.column_stack([X2, np.full(nn2, k / nz)]) for k in range(nz + 1)])
    conn = np.vstack([np.hstack([quads + k * nn2, quads + (k + 1) * nn2]) for k in range(nz)])
    X[:, 2] += np.random.default_rng(seed + 1).uniform(-0.1, 0.1, size=X.shape[0]) / nz
    return X, conn.astype(np.int64)


@pytest.mark.parametrize("mode", ["auto", "atomic", "gather"])
@pytest.mark.parametrize("case", ["rotated", "extruded_y", "lattice"])
def test_hex_elasticity_irregular_meshes(pf, mode, case):
    from pyfem_gpu_testflight_b200 import _lib
    if case == "rotated":
        X, conn = _hex_jittered(9, 8, 7, seed=3)
        conn = _rotate_local_numbering(conn, seed=4)
    elif case == "extruded_y":
        X, conn = extruded_three_patch_mesh(5, 4, seed=2)
    else:
        X, conn = _hex_jittered(12, 7, 9, seed=5)
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    Kr = orc.assemble_elasticity(X, conn, rho, 4.0, 6.0, 0.25)
    q = pf.QuadratureBlock3D()
    m = pf.LinearElasticity(X, conn, [0], None, {0: [0.0, 0.0, 0.0]}, q, pf.BasisBlock3D(q), E=6.0, nu=0.25, p=4.0,
                            scatter=mode)
    assert_csr_matches(m.compute_jacobian(rho), Kr.indptr, Kr.indices, Kr.data)
    if case == "lattice":
        assert m.mesh.info(_lib.INFO_HEX_ROWS) == 1  # the chunk-row pass is what ran for auto / gather


def test_hex_chunk_rows_bitwise_reproducible(pf):
    X, conn = _hex_jittered(11, 10, 9, seed=7)
    mesh = pf.DeviceMesh(X, conn, 3)
    rho = 0.2 + 0.8 * np.random.default_rng(1).random(X.shape[0])
    v1 = mesh.assemble_elasticity(rho, 3.0, mode="gather").cpu().numpy()
    v2 = mesh.assemble_elasticity(rho, 3.0, mode="gather").cpu().numpy()
    va = mesh.assemble_elasticity(rho, 3.0, mode="atomic").cpu().numpy()
    assert np.array_equal(v1, v2)
    assert_values_close(v1, va, 1e-13)


def test_hex_chunk_rows_large_chunks_second_round(pf, monkeypatch):
    """Chunks with more nodes than the consumer warps take in one round (7 warps x 4 nodes): the second round reads
    its node / incidence tables directly instead of from the prefetch slots."""
    from pyfem_gpu_testflight_b200 import _lib
    monkeypatch.setenv("PFG_HEX_CHUNK_NODES", "45")
    X, conn = _hex_jittered(10, 9, 8, seed=11)
    rho = 0.05 + 0.95 * np.random.default_rng(3).random(X.shape[0])
    mesh = pf.DeviceMesh(X, conn, 3)
    assert mesh.info(_lib.INFO_HEX_ROWS) == 1
    Kr = orc.assemble_elasticity(X, conn, rho, 2.0)
    v = mesh.assemble_elasticity(rho, 2.0, mode="gather").cpu().numpy()
    assert_values_close(v, Kr.data, VAL_TOL)


def test_hex_chunk_rows_fall_back_when_chunks_exceed_shared_memory(pf, monkeypatch):
    """Chunks whose element records do not fit the geometry ring: the handle drops the chunk-row pass when it is
    built; AUTO assembles with the atomic scatter, GATHER says so."""
    from pyfem_gpu_testflight_b200 import _lib
    monkeypatch.setenv("PFG_HEX_CHUNK_NODES", "150")
    X, conn = _hex_jittered(11, 10, 9, seed=13)
    mesh = pf.DeviceMesh(X, conn, 3)
    assert mesh.info(_lib.INFO_HEX_ROWS) == 0
    Kr = orc.assemble_elasticity(X, conn, 1.0, 0.0)
    assert_values_close(mesh.assemble_elasticity(mode="auto").cpu().numpy(), Kr.data, VAL_TOL)
    with pytest.raises(NotImplementedError):
        mesh.assemble_elasticity(mode="gather")


# ---- full-size properties for BASELINE configs[2] (Helmholtz 4096 x 2048) and configs[4] (hex8 256^3) ---------------
def test_full_size_helmholtz_properties(pf):
    """C3: closed-form nnz, bitwise repeatability, atomic-vs-gather agreement, sum(R) = area of the domain
    (R_ij = int N_i N_j, a partition of unity), K - R = r0^2 * stiffness annihilates constants, symmetry."""
    import torch
    nx, ny = 4096, 2048
    c = pf.ProblemCreator(nx + 1, ny + 1)
    mesh = pf.DeviceMesh(c.X, c.conn, 1)
    nn = (nx + 1) * (ny + 1)
    assert mesh.nelems == nx * ny == 8388608
    assert mesh.nnz == (3 * (nx + 1) - 2) * (3 * (ny + 1) - 2) == 75515905
    assert mesh.idx_bytes == 4
    r0 = 0.05
    K, R = mesh.assemble_helmholtz(r0, mode="gather")
    K2, R2 = mesh.assemble_helmholtz(r0, mode="gather")
    assert torch.equal(K, K2) and torch.equal(R, R2)
    del K2, R2
    Ka, Ra = mesh.assemble_helmholtz(r0, mode="atomic")
    assert float((Ka - K).abs().max()) <= 1e-13 * float(K.abs().max())
    assert float((Ra - R).abs().max()) <= 1e-13 * float(R.abs().max())
    del Ka, Ra
    area = float(c.X[:, 0].max() * c.X[:, 1].max())
    assert abs(float(R.sum()) - area) <= 1e-11 * area
    ones = torch.ones(nn, dtype=torch.float64, device="cuda")
    row_sums = mesh.spmv(R, ones)
    assert abs(float(row_sums.sum()) - area) <= 1e-11 * area
    # (K - R) 1 = r0^2 * (stiffness . 1) = 0
    assert float((mesh.spmv(K, ones) - row_sums).abs().max()) <= 1e-12 * float(K.abs().max())
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(nn, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.rand(nn, dtype=torch.float64, device="cuda", generator=g) - 0.5
    a, b = float(x @ mesh.spmv(K, y)), float(y @ mesh.spmv(K, x))
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
    # R^T x through the transposed product equals R x for the symmetric R
    assert float((mesh.spmv_t(R, x) - mesh.spmv(R, x)).abs().max()) <= 1e-13 * float(R.abs().max())


def _lattice_row_columns(dims, node_ijk, m):
    """Expected CSR columns of the m dof rows of a lattice node (sorted neighbour ids, interleaved dofs)."""
    nbrs = []
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                p = [node_ijk[0] + di, node_ijk[1] + dj, node_ijk[2] + dk]
                if all(0 <= p[a] < dims[a] for a in range(3)):
                    nbrs.append(p[0] + dims[0] * (p[1] + dims[1] * p[2]))
    nbrs = np.sort(np.array(nbrs, dtype=np.int64))
    return (m * nbrs[:, None] + np.arange(m)[None, :]).ravel()


def test_int64_index_pattern_hex(pf):
    """SURVEY trap T2 at the size where it bites: 156^3 hexes with 3 dofs per node have 156^3 * 576 > 2^31 - 1 COO
    entries, so scipy's coo -> csr (and this engine) switch indptr AND indices to int64.  k_write_indices<long> runs."""
    import torch
    n = 156
    c = pf.ProblemCreator(n + 1, n + 1, n + 1)
    mesh = pf.DeviceMesh(c.X, c.conn, 3)
    assert mesh.nelems * 576 > 2 ** 31 - 1
    assert mesh.idx_bytes == 8 and mesh.index_dtype == np.int64
    assert mesh.nnz == 9 * (3 * (n + 1) - 2) ** 3
    indptr, indices = mesh.pattern()
    assert indptr.dtype == torch.int64 and indices.dtype == torch.int64
    assert int(indptr[0]) == 0 and int(indptr[-1]) == mesh.nnz
    dims = (n + 1, n + 1, n + 1)
    for ijk in ((0, 0, 0), (1, 0, 0), (5, 7, 3), (n, n, n), (n, 3, 77)):
        node = ijk[0] + dims[0] * (ijk[1] + dims[1] * ijk[2])
        want = _lattice_row_columns(dims, ijk, 3)
        for alpha in range(3):
            a, b = int(indptr[3 * node + alpha]), int(indptr[3 * node + alpha + 1])
            assert np.array_equal(indices[a:b].cpu().numpy(), want)
    # values of a corner patch equal the oracle's rows on a small mesh with the same element size
    K = mesh.assemble_elasticity(1.0, 0.0)
    Xs, cs = orc.structured_mesh(6, 6, 6, Lx=5.0 / n * c.X[:, 0].max(), Ly=5.0 / n * c.X[:, 1].max(),
                                 Lz=5.0 / n * c.X[:, 2].max())
    Ks = orc.assemble_elasticity(Xs, cs)
    for (i, j, k) in ((0, 0, 0), (2, 1, 3), (4, 4, 4)):
        rs, rb = 3 * (i + 6 * (j + 6 * k)), 3 * (i + dims[0] * (j + dims[1] * k))
        got = K[int(indptr[rb]):int(indptr[rb + 1])].cpu().numpy()
        assert_values_close(got, Ks.data[Ks.indptr[rs]:Ks.indptr[rs + 1]], VAL_TOL, f"row of node ({i},{j},{k})")


def test_slab_index_dtype_follows_global_rule(pf):
    """A rank's slab reports the index dtype of the reference's GLOBAL matrix (engine.index_bytes_rule): small
    slabs of a mesh whose global COO count exceeds 2^31 - 1 are int64, like the matrix they are rows of."""
    from pyfem_gpu_testflight_b200.engine import index_bytes_rule
    from pyfem_gpu_testflight_b200.partition import structured_slab
    assert index_bytes_rule(16777216, 8, 33570818) == 4      # C2: COO nnz = 2^30
    assert index_bytes_rule(16777216, 24, 50923779) == 8     # C5
    assert index_bytes_rule(2097152, 24, 50923779) == 4      # one eighth of C5 judged by its own count
    part = structured_slab(9, 8, 11, 1, 3)
    local = pf.DeviceMesh(part.X, part.conn, 3, own_range=part.own_range, node_gid=part.node_gid,
                          ncols_nodes=part.nnodes_global)
    assert local.idx_bytes == 4 and local.pattern_host()[1].dtype == np.int32
    as_c5 = pf.DeviceMesh(part.X, part.conn, 3, own_range=part.own_range, node_gid=part.node_gid,
                          ncols_nodes=part.nnodes_global, nelems_global=16777216)
    assert as_c5.idx_bytes == 8
    ip, ix = as_c5.pattern_host()
    assert ip.dtype == np.int64 and ix.dtype == np.int64
    assert np.array_equal(ix, local.pattern_host()[1]) and np.array_equal(ip, local.pattern_host()[0])


def test_full_size_hex_elasticity_properties(pf):
    """C5 on one GPU (hex8 256^3, 4.09 G CSR values, int64 pattern): closed-form nnz, bitwise repeatability,
    agreement with the atomic scatter, the six rigid-body modes (three translations, three rotations) in the null
    space, symmetry, translation invariance of interior rows and oracle rows of a corner patch."""
    import torch
    n = 256
    c = pf.ProblemCreator(n + 1, n + 1, n + 1)
    mesh = pf.DeviceMesh(c.X, c.conn, 3)
    nn = (n + 1) ** 3
    assert mesh.nelems == n ** 3 == 16777216
    assert mesh.nnz == 9 * (3 * (n + 1) - 2) ** 3 == 4092809481
    assert mesh.idx_bytes == 8
    K = mesh.assemble_elasticity(1.0, 0.0)
    K2 = mesh.assemble_elasticity(1.0, 0.0)
    assert torch.equal(K, K2)
    scale = float(K.abs().max())
    mesh.assemble_elasticity(1.0, 0.0, out=K2, mode="atomic")
    K2 -= K
    assert float(K2.abs().max()) <= 1e-13 * scale
    del K2
    X = torch.as_tensor(c.X, device="cuda")
    for axis in range(3):  # translations
        t = torch.zeros(3 * nn, dtype=torch.float64, device="cuda")
        t[axis::3] = 1.0
        assert float(mesh.spmv(K, t).abs().max()) <= 1e-12 * scale
    for axis in range(3):  # infinitesimal rotations u = e_axis x (x - centre)
        xc = X - 0.5
        e = torch.zeros(3, dtype=torch.float64, device="cuda")
        e[axis] = 1.0
        u = torch.cross(e.expand_as(xc), xc, dim=1).reshape(-1).contiguous()
        assert float(mesh.spmv(K, u).abs().max()) <= 1e-11 * scale
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(3 * nn, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.rand(3 * nn, dtype=torch.float64, device="cuda", generator=g) - 0.5
    a, b = float(x @ mesh.spmv(K, y)), float(y @ mesh.spmv(K, x))
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b), scale)
    d = n + 1
    # interior rows: 3 x 81 values each, identical on the uniform mesh (the device pattern is not materialised
    # on the host for 4 G entries: row offsets follow from the closed-form row lengths of interior planes)
    indptr, _ = mesh.pattern()
    rows = [3 * (i + d * (j + d * k)) for (i, j, k) in ((7, 9, 11), (128, 128, 128), (250, 3, 77))]
    ref = K[int(indptr[rows[0]]):int(indptr[rows[0] + 3])]
    assert ref.numel() == 3 * 81
    for r in rows[1:]:
        assert float((K[int(indptr[r]):int(indptr[r + 3])] - ref).abs().max()) <= 1e-12 * scale
    Xs, cs = orc.structured_mesh(6, 6, 6, Lx=5.0 / n, Ly=5.0 / n, Lz=5.0 / n)
    Ks = orc.assemble_elasticity(Xs, cs)
    for (i, j, k) in ((0, 0, 0), (2, 1, 3), (4, 4, 4)):
        rs, rb = 3 * (i + 6 * (j + 6 * k)), 3 * (i + d * (j + d * k))
        got = K[int(indptr[rb]):int(indptr[rb + 1])].cpu().numpy()
        assert_values_close(got, Ks.data[Ks.indptr[rs]:Ks.indptr[rs + 1]], VAL_TOL, f"row of node ({i},{j},{k})")
