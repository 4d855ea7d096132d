"""Pin oracle/pyfem_oracle.py to the reference's own outputs (tests/golden, written by
oracle/make_golden.py from the unmodified reference).  CPU only."""
import numpy as np
import pytest

import pyfem_oracle as orc
from make_golden import test_gfunc as gfunc
from parity import assert_csr_matches, assert_values_close, golden_files

ORACLE_TOL = 1e-13  # same arithmetic as the reference; only einsum contraction order may differ


@pytest.mark.parametrize("path", golden_files("poisson"))
def test_poisson(path):
    g = np.load(path)
    rho = g["rho"] if g["rho"].ndim else float(g["rho"])
    K = orc.assemble_poisson(g["X"], g["conn"], rho, float(g["p"]))
    assert_csr_matches(K, g["K_indptr"], g["K_indices"], g["K_data"], ORACLE_TOL)
    rhs = orc.assemble_poisson_rhs(g["X"], g["conn"], gfunc)
    assert_values_close(rhs, g["rhs"], ORACLE_TOL, "rhs")


@pytest.mark.parametrize("path", golden_files("elasticity"))
def test_elasticity(path):
    g = np.load(path)
    rho = g["rho"] if g["rho"].ndim else float(g["rho"])
    K = orc.assemble_elasticity(g["X"], g["conn"], rho, float(g["p"]), float(g["E"]), float(g["nu"]))
    assert_csr_matches(K, g["K_indptr"], g["K_indices"], g["K_data"], ORACLE_TOL)


@pytest.mark.parametrize("path", golden_files("helmholtz"))
def test_helmholtz(path):
    g = np.load(path)
    K, R = orc.assemble_helmholtz(g["X"], g["conn"], float(g["r0"]))
    assert_csr_matches(K, g["K_indptr"], g["K_indices"], g["K_data"], ORACLE_TOL)
    assert_csr_matches(R, g["R_indptr"], g["R_indices"], g["R_data"], ORACLE_TOL)
    assert_values_close(R.dot(g["x"]), g["rhs"], ORACLE_TOL, "R x")


@pytest.mark.parametrize("path", golden_files("nlpoisson"))
def test_nlpoisson(path):
    g = np.load(path)
    K, res = orc.assemble_nlpoisson(g["X"], g["conn"], g["xdv"], g["u"])
    assert_csr_matches(K, g["K_indptr"], g["K_indices"], g["K_data"], ORACLE_TOL)
    assert_values_close(res, g["res"], ORACLE_TOL, "residual")


@pytest.mark.parametrize("path", golden_files("sens"))
def test_sensitivities(path):
    g = np.load(path)
    d = orc.poisson_K_dv_sens(g["X"], g["conn"], g["rho"], float(g["p_poisson"]), g["phi"], g["psi"])
    assert_values_close(d, g["g_poisson"], ORACLE_TOL, "poisson dK/drho")
    d = orc.elasticity_K_dv_sens(g["X"], g["conn"], g["rho"], float(g["p_elast"]), g["phi_v"], g["psi_v"],
                                 float(g["E"]), float(g["nu"]))
    assert_values_close(d, g["g_elast"], ORACLE_TOL, "elasticity dK/drho")


def test_structured_mesh_matches_closed_form_nnz():
    # nnz = m^2 * prod(3 nn_k - 2)
    X, conn = orc.structured_mesh(9, 7)
    K = orc.assemble_elasticity(X, conn)
    assert K.nnz == 4 * (3 * 9 - 2) * (3 * 7 - 2)
    X, conn = orc.structured_mesh(4, 5, 3)
    K = orc.assemble_poisson(X, conn)
    assert K.nnz == (3 * 4 - 2) * (3 * 5 - 2) * (3 * 3 - 2)
