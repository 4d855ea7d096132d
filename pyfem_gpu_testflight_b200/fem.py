"""Quadrature rules and basis functions with the reference's interface (pyfem.py:19-65, 175-250).

On the device these tables are compile-time constants (csrc/pfg_elem.cuh, struct Elem<NNE>); the host
objects exist because the model constructors take them (`LinearPoisson(X, conn, ..., quadrature, basis)`)
and callers read `get_nquads()`, `get_weight()`, `eval_shape_fun()` and `eval_shape_fun_deriv()`.
Only the element families BASELINE.json's configs use have a device path: bilinear quad4 and
trilinear hex8 with 2-point Gauss rules.
"""
import numpy as np

_G = 1.0 / np.sqrt(3.0)

# local node signs, counter-clockwise bottom face first (pyfem.py:263-268, 297-306)
_QUAD_SIGNS = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=float)
_HEX_SIGNS = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                       [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)


class QuadratureBase:
    """Points `pts` (nquads, ndims) and `weights` (nquads,) (pyfem.py:19-65)."""

    device_elem = None  # nnodes_per_elem of the matching device element, None = no device path

    def __init__(self, pts, weights):
        pts = np.asarray(pts, dtype=float)
        weights = np.asarray(weights, dtype=float)
        assert len(pts) == len(weights)
        self.pts = pts
        self.weights = weights
        self.nquads = pts.shape[0]

    def get_nquads(self):
        return self.nquads

    def get_pt(self, idx=None):
        # like the reference, a falsy idx (None or 0) returns every point (pyfem.py:51; survey trap T5)
        return self.pts[idx] if idx else self.pts

    def get_weight(self, idx=None):
        return self.weights[idx] if idx else self.weights


class QuadratureBilinear2D(QuadratureBase):
    """2x2 Gauss points in counter-clockwise order, unit weights (pyfem.py:83-95)."""

    device_elem = 4

    def __init__(self):
        super().__init__(_QUAD_SIGNS * _G, np.ones(4))


class QuadratureBlock3D(QuadratureBase):
    """2x2x2 Gauss points, x slowest and z fastest, unit weights (pyfem.py:97-112)."""

    device_elem = 8

    def __init__(self):
        grid = np.array([[sx, sy, sz] for sx in (-1.0, 1.0) for sy in (-1.0, 1.0) for sz in (-1.0, 1.0)])
        super().__init__(grid * _G, np.ones(8))


class BasisBase:
    """Shape functions N (nquads, nnodes_per_elem) and reference-space derivatives
    Nderiv (nquads, nnodes_per_elem, ndims), evaluated once and cached (pyfem.py:175-250)."""

    device_elem = None
    _signs = None

    def __init__(self, ndims, nnodes_per_elem, quadrature):
        self.ndims = ndims
        self.nnodes_per_elem = nnodes_per_elem
        self.quadrature = quadrature
        self.nquads = quadrature.get_nquads()
        self.N = None
        self.Nderiv = None

    def _factors(self):
        # f[q, a, k] = 1 + sign[a, k] * xi[q, k]
        pts = np.asarray(self.quadrature.get_pt())
        return 1.0 + self._signs[None, :, :] * pts[:, None, :]

    def eval_shape_fun(self):
        if self.N is None:
            self.N = np.prod(self._factors(), axis=2) / float(2 ** self.ndims)
        return self.N

    def eval_shape_fun_deriv(self):
        if self.Nderiv is None:
            f = self._factors()
            d = np.empty((self.nquads, self.nnodes_per_elem, self.ndims))
            for k in range(self.ndims):
                others = [j for j in range(self.ndims) if j != k]
                d[:, :, k] = self._signs[None, :, k] * np.prod(f[:, :, others], axis=2) / float(2 ** self.ndims)
            self.Nderiv = d
        return self.Nderiv


class BasisBilinear2D(BasisBase):
    """4-node bilinear quadrilateral (pyfem.py:253-284)."""

    device_elem = 4
    _signs = _QUAD_SIGNS

    def __init__(self, quadrature):
        super().__init__(2, 4, quadrature)


class BasisBlock3D(BasisBase):
    """8-node trilinear hexahedron (pyfem.py:287-338)."""

    device_elem = 8
    _signs = _HEX_SIGNS

    def __init__(self, quadrature):
        super().__init__(3, 8, quadrature)


def _no_device_path(name, where):
    def ctor(*args, **kwargs):
        raise NotImplementedError(
            f"{name} ({where}) has no B200 device path: this engine covers the quad4 / hex8 assembly "
            "configurations of BASELINE.json and deliberately has no CPU fallback")
    ctor.__name__ = name
    return ctor


# element families of the reference that are outside the hot-path scope
QuadratureTriangle2D = _no_device_path("QuadratureTriangle2D", "pyfem.py:68-80")
QuadratureTetrahedron5Point = _no_device_path("QuadratureTetrahedron5Point", "pyfem.py:115-134")
QuadratureBrick333Point = _no_device_path("QuadratureBrick333Point", "pyfem.py:137-172")
BasisTriangle2D = _no_device_path("BasisTriangle2D", "pyfem.py:341-377")
BasisTetrahedron10node = _no_device_path("BasisTetrahedron10node", "pyfem.py:380-445")
BasisBrick20Nodes = _no_device_path("BasisBrick20Nodes", "pyfem.py:448-631")
