"""ctypes binding of include/pyfem_b200.h (libpyfem_b200.so).

There is no CPU fallback: if the library is missing and cannot be built, or a CUDA call fails, this
raises.  Status codes map to the reference's exception classes (ValueError for bad input, AssertionError
for the conn sanity asserts of pyfem.py:680-681, RuntimeError otherwise).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_void_p

from . import _build

PFG_OK = 0
PFG_ERR_INVALID = -1
PFG_ERR_CUDA = -2
PFG_ERR_UNSUPPORTED = -3
PFG_ERR_MESH = -4
PFG_ERR_NOMEM = -5
PFG_ERR_NOCONV = -6

PFG_QUAD4 = 4
PFG_HEX8 = 8

MODE_AUTO = 0
MODE_ATOMIC = 1
MODE_GATHER = 2
MODES = {"auto": MODE_AUTO, "atomic": MODE_ATOMIC, "gather": MODE_GATHER}

INFO_NNZ, INFO_NROWS, INFO_NCOLS, INFO_IDX_BYTES, INFO_NCHUNKS, INFO_CHUNK_ELEMS, INFO_PLAN_BYTES, \
    INFO_DEVICE_BYTES, INFO_MAX_ROW_BLOCKS, INFO_MAX_VALENCE, INFO_HEX_ROWS, INFO_TEMPLATES = range(12)

PHYS_POISSON, PHYS_ELASTICITY, PHYS_HELMHOLTZ, PHYS_NLPOISSON = 1, 2, 3, 4

CREATE_NO_GATHER_PLAN = 1
CREATE_NO_REORDER = 2

REDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_int)  # pfg_reduce_fn(user, offset, count)
HALO_FN = ctypes.CFUNCTYPE(c_int, c_void_p)                 # pfg_halo_fn(user)
HALO2_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int)         # pfg_halo2_fn(user, which)

# name -> (restype, argtypes); must list every PFG_API symbol of include/pyfem_b200.h
PROTOTYPES = {
    "pfg_abi_version": (c_int, []),
    "pfg_last_error": (c_char_p, []),
    "pfg_mesh_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "pfg_mesh_destroy": (c_int, [c_void_p]),
    "pfg_mesh_get": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "pfg_mesh_pattern": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pfg_assemble_poisson": (c_int, [c_void_p, c_void_p, c_double, c_double, c_void_p, c_int, c_void_p]),
    "pfg_assemble_elasticity": (c_int, [c_void_p, c_void_p, c_double, c_double, c_double, c_double, c_void_p,
                                        c_int, c_void_p]),
    "pfg_assemble_helmholtz": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "pfg_assemble_nlpoisson": (c_int, [c_void_p, POINTER(c_double), c_int, c_void_p, c_void_p, c_void_p, c_int,
                                       c_void_p]),
    "pfg_assemble_poisson_complex": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "pfg_assemble_elasticity_complex": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_void_p,
                                                c_void_p, c_void_p]),
    "pfg_quad_points": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pfg_poisson_rhs": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pfg_apply_dirichlet": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "pfg_spmv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pfg_spmv_t": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pfg_cg": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double, c_double, c_int, c_int,
                       POINTER(c_int), POINTER(c_double), c_void_p]),
    "pfg_cg_dist": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_double,
                            c_double, c_int, c_int, REDUCE_FN, HALO_FN, c_void_p, POINTER(c_int), POINTER(c_double),
                            c_void_p]),
    "pfg_cg_dist_begin": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, REDUCE_FN,
                                  HALO_FN, c_void_p, c_void_p]),
    "pfg_cg_dist_steps": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, REDUCE_FN,
                                  HALO_FN, c_void_p, c_void_p]),
    "pfg_bicgstab_dist": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double,
                                  c_int, c_int, REDUCE_FN, HALO2_FN, c_void_p, POINTER(c_int), POINTER(c_double),
                                  c_void_p]),
    "pfg_scatter_matrix": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pfg_scatter_vector": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pfg_element_matrices": (c_int, [c_void_p, c_int, c_void_p, c_double, POINTER(c_double), c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "pfg_k_dv_sens": (c_int, [c_void_p, c_int, c_void_p, c_double, c_double, POINTER(c_double), c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p]),
    "pfg_bicgstab": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_int, POINTER(c_int),
                             POINTER(c_double), c_void_p]),
    "pfg_k_dv_sens_ordered": (c_int, [c_void_p, c_int, c_void_p, c_double, c_double, POINTER(c_double), c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "pfg_probe_fp64": (c_int, [c_int, c_int, POINTER(c_double), POINTER(c_double)]),
    "pfg_mesh_set_element_mask": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pfg_add_indexed": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the sources are newer) and return the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    # build() returns at once unless a source / header is newer than the library (or the library is missing); where
    # there is no compiler (a box that received the prebuilt .so) a present library is loaded as it is
    try:
        _build.build(force=os.environ.get("PFG_REBUILD") == "1")
    except RuntimeError:
        if not os.path.isfile(path):
            raise
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.pfg_abi_version() != 1:
        raise RuntimeError("libpyfem_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status):
    """Raise the reference-style exception for a non-zero status."""
    if status == PFG_OK:
        return
    msg = load().pfg_last_error().decode("utf-8", "replace")
    if status == PFG_ERR_INVALID:
        raise ValueError(msg)
    if status == PFG_ERR_MESH:
        raise AssertionError(msg)
    if status == PFG_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == PFG_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
