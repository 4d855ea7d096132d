"""
Generate tests/golden/*.npz by running the UNMODIFIED reference in the build
container -- TEST INFRASTRUCTURE.  Run from the repo root:

    python oracle/make_golden.py

Each file holds the inputs of one case and the reference's own outputs
(CSR indptr / indices / data of compute_jacobian, and compute_rhs where the
physics has one).  The fixtures are small (<= 961 elements) and are what pins
oracle/pyfem_oracle.py and the CUDA path to the reference on the GPU box, where
/root/reference does not exist.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402
from pyfem_oracle import structured_mesh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def test_gfunc(x):
    # the source term of tests/test_linear_poisson.py:11-14
    _x = x[..., 0]
    _y = x[..., 1]
    return _x * (_x - 5.0) * (_x - 10.0) * _y * (_y - 4.0)


def jitter(X, shape, seed=1, frac=0.2):
    rng = np.random.default_rng(seed)
    ext = X.max(axis=0) - X.min(axis=0)
    h = ext / (np.array(shape[: X.shape[1]]) - 1)
    return X + rng.uniform(-frac, frac, size=X.shape) * h


def csr_fields(K, prefix="K"):
    assert K.has_sorted_indices
    return {f"{prefix}_indptr": K.indptr, f"{prefix}_indices": K.indices, f"{prefix}_data": K.data,
            f"{prefix}_shape": np.array(K.shape)}


def main():
    os.makedirs(OUT, exist_ok=True)
    os.chdir(tempfile.mkdtemp())  # keep any stray profiler.log out of the repo
    ref = ref_import.load()

    def quad_objs():
        q = ref.QuadratureBilinear2D()
        return q, ref.BasisBilinear2D(q)

    def hex_objs():
        q = ref.QuadratureBlock3D()
        return q, ref.BasisBlock3D(q)

    meshes = {}
    X, conn = structured_mesh(7, 6)
    meshes["quad_jit"] = (jitter(X, (7, 6)), conn, quad_objs)
    X, conn = structured_mesh(9, 5)
    perm = np.random.default_rng(7).permutation(conn.shape[0])
    meshes["quad_perm"] = (jitter(X, (9, 5), seed=3), conn[perm], quad_objs)
    X, conn = structured_mesh(32, 32)
    meshes["quad_32"] = (X, conn, quad_objs)  # the mesh of the reference's own tests
    X, conn = structured_mesh(4, 3, 3)
    meshes["hex_jit"] = (jitter(X, (4, 3, 3)), conn, hex_objs)
    X, conn = structured_mesh(5, 4, 6)
    perm = np.random.default_rng(11).permutation(conn.shape[0])
    meshes["hex_perm"] = (jitter(X, (5, 4, 6), seed=5), conn[perm], hex_objs)

    for mname, (X, conn, objs) in meshes.items():
        nn = X.shape[0]
        rng = np.random.default_rng(0)
        rho = 0.05 + 0.95 * rng.random(nn)
        quadrature, basis = objs()

        # ---- linear Poisson: K(rho, p) and rhs(gfunc)
        for tag, (r, p) in {"default": (1.0, 0.0), "ramp": (rho, 3.0)}.items():
            m = ref.LinearPoisson(X, conn, [0], None, quadrature, basis, test_gfunc, p=p)
            K = m.compute_jacobian(r) if tag == "ramp" else m.compute_jacobian()
            rhs = m.compute_rhs().copy()
            np.savez_compressed(os.path.join(OUT, f"poisson_{mname}_{tag}.npz"), X=X, conn=conn,
                                rho=np.asarray(r, dtype=float), p=p, rhs=rhs, **csr_fields(K))

        # ---- linear elasticity (plane stress in 2-D)
        for tag, (r, p) in {"default": (1.0, 0.0), "ramp": (rho, 5.0)}.items():
            m = ref.LinearElasticity(X, conn, [0], None, {0: [0.0] * X.shape[1]}, quadrature, basis,
                                     E=10.0 if tag == "default" else 7.5,
                                     nu=0.3 if tag == "default" else 0.22, p=p)
            K = m.compute_jacobian(r) if tag == "ramp" else m.compute_jacobian()
            np.savez_compressed(os.path.join(OUT, f"elasticity_{mname}_{tag}.npz"), X=X, conn=conn,
                                rho=np.asarray(r, dtype=float), p=p,
                                E=10.0 if tag == "default" else 7.5,
                                nu=0.3 if tag == "default" else 0.22, **csr_fields(K))

        # ---- Helmholtz filter: K and R, one pattern
        r0 = 0.1
        m = ref.Helmholtz(r0, X, conn, quadrature, basis)
        xin = rng.random(nn)
        np.savez_compressed(os.path.join(OUT, f"helmholtz_{mname}.npz"), X=X, conn=conn, r0=r0,
                            x=xin, rhs=m.compute_rhs(xin).copy(),
                            **csr_fields(m.K, "K"), **csr_fields(m.R, "R"))

        # ---- nonlinear Poisson (2-D only): Jacobian + residual at a non-trivial state
        if X.shape[1] == 2:
            Xn = (X - X.min(axis=0)) / (X.max(axis=0) - X.min(axis=0))  # unit square
            m = ref.NonlinearPoisson2D(Xn, conn, [0], None, quadrature, basis)
            xdv = np.ones(10) / 10.0 if mname == "quad_32" else rng.random(6)
            u = rng.random(nn) - 0.3
            K = m.compute_jacobian(xdv, u)
            res = m.compute_rhs(xdv, u).copy()
            np.savez_compressed(os.path.join(OUT, f"nlpoisson_{mname}.npz"), X=Xn, conn=conn, xdv=xdv,
                                u=u, res=res, **csr_fields(K))

        # ---- sensitivities d(phi^T K psi)/d rho (drawn after everything above: earlier files keep their inputs)
        d = X.shape[1]
        phi, psi = rng.random(nn), rng.random(nn)
        m = ref.LinearPoisson(X, conn, [0], None, quadrature, basis, test_gfunc, p=3.0)
        g_poisson = m._compute_K_dv_sens(rho, phi, psi).copy()
        phi_v, psi_v = rng.random(nn * d) - 0.5, rng.random(nn * d) - 0.5
        m = ref.LinearElasticity(X, conn, [0], None, {0: [0.0] * d}, quadrature, basis, E=7.5, nu=0.22, p=5.0)
        g_elast = m._compute_K_dv_sens(rho, phi_v, psi_v).copy()
        np.savez_compressed(os.path.join(OUT, f"sens_{mname}.npz"), X=X, conn=conn, rho=rho, phi=phi, psi=psi,
                            p_poisson=3.0, g_poisson=g_poisson, phi_v=phi_v, psi_v=psi_v, p_elast=5.0, E=7.5,
                            nu=0.22, g_elast=g_elast)

    print("wrote", len(os.listdir(OUT)), "files to", OUT)


if __name__ == "__main__":
    main()
