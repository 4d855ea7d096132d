"""Times the UNMODIFIED reference (imported from /root/reference, build container only) beside the oracle port on the
same sample, to show that bench.py's CPU arm (`kind: "port"`; the reference's sources do not travel to the GPU box) is
representative of the reference's own cost:  python oracle/compare_cpu_arms.py [n_side] > profiles/r02_cpu_reference_vs_port.json
Test infrastructure (not imported by the product)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyfem_oracle as orc  # noqa: E402
import ref_import  # noqa: E402


def best_of(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    ref = ref_import.load()
    out = {"sample_elements": n * n, "numpy": np.__version__, "host_cores": os.cpu_count(), "cases": {}}
    c = ref.ProblemCreator(n + 1, n + 1)
    conn, X, dof_fixed, force = c.create_linear_elasticity_problem()
    q = ref.QuadratureBilinear2D()
    model = ref.LinearElasticity(X, conn, dof_fixed, None, force, q, ref.BasisBilinear2D(q))
    t_ref = best_of(lambda: model.compute_jacobian())
    t_port = best_of(lambda: orc.assemble_elasticity(np.asarray(X, dtype=float), np.asarray(conn)))
    out["cases"]["c2_elasticity_quad"] = {"reference_elem_per_s": n * n / t_ref, "port_elem_per_s": n * n / t_port}
    conn, X, dof_fixed = c.create_poisson_problem()
    Xn = np.asarray(X, dtype=float) / np.max(X, axis=0)
    nl = ref.NonlinearPoisson2D(Xn, conn, dof_fixed, None, q, ref.BasisBilinear2D(q))
    xdv, u = np.ones(10) / 10.0, np.random.default_rng(0).random(Xn.shape[0])
    t_ref = best_of(lambda: (nl.compute_jacobian(xdv, u), nl.compute_rhs(xdv, u)))
    t_port = best_of(lambda: orc.assemble_nlpoisson(Xn, np.asarray(conn), xdv, u))
    out["cases"]["c4_nlpoisson_quad"] = {"reference_elem_per_s": n * n / t_ref, "port_elem_per_s": n * n / t_port}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
