"""Shared parity helpers: the contract of SURVEY.md section 8(c).

Pattern: indptr / indices equal as integer arrays AND same dtype / shape (bit-exact).
Values:  max|a - b| <= tol * max|b|  (max-abs norms; tol = 1e-12 per BASELINE.json's north_star;
         the same norm-relative form the reference uses in examples/SciTech2023/verification/verify.py:37-38).
"""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VAL_TOL = 1e-12


def golden_files(prefix):
    files = sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))
    assert files, f"no golden fixtures for {prefix}"
    return files


def assert_pattern_equal(indptr, indices, ref_indptr, ref_indices):
    assert indptr.dtype == ref_indptr.dtype, (indptr.dtype, ref_indptr.dtype)
    assert indices.dtype == ref_indices.dtype, (indices.dtype, ref_indices.dtype)
    assert indptr.shape == ref_indptr.shape
    assert indices.shape == ref_indices.shape
    assert np.array_equal(indptr, ref_indptr)
    assert np.array_equal(indices, ref_indices)


def norm_rel(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.max(np.abs(b)) if b.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / scale)


def assert_values_close(a, b, tol=VAL_TOL, what="values"):
    err = norm_rel(a, b)
    assert err <= tol, f"{what}: max|d|/max|ref| = {err:.3e} > {tol:.1e}"


def assert_csr_matches(K, ref_indptr, ref_indices, ref_data, tol=VAL_TOL):
    assert_pattern_equal(K.indptr, K.indices, ref_indptr, ref_indices)
    assert_values_close(K.data, ref_data, tol, "csr data")
