"""Vectorised structured-mesh / problem generator with ProblemCreator's interface (pyfem.py:2426-2773).

Produces the same `conn`, `X`, `dof_fixed`, `nodal_force` and `x` arrays as the reference's Python loops
(node id = i + j*nx + k*nx*ny, pyfem.py:2479-2486; quad connectivity :2499-2502; hex :2527-2534) without
the loops, so 16 M-element inputs take a second instead of minutes.  quad and block elements only.
"""
import numpy as np


class ProblemCreator:
    def __init__(self, nnodes_x, nnodes_y, nnodes_z=None, Lx=None, Ly=None, Lz=None, element_type=None):
        if nnodes_z is None:
            self.ndims = 2
            nnodes_z = 1
            element_type = "quad" if element_type is None else element_type
            if element_type not in ("quad", "tri"):
                raise AssertionError("2-D meshes are 'quad' or 'tri'")
        else:
            self.ndims = 3
            element_type = "block" if element_type is None else element_type
            if element_type not in ("block", "tet", "brick20"):
                raise AssertionError("3-D meshes are 'block', 'tet' or 'brick20'")
        if element_type not in ("quad", "block"):
            raise NotImplementedError(f"element_type {element_type!r} has no device path (quad and block only)")
        Lx = (nnodes_x - 1) / (nnodes_y - 1) if Lx is None else Lx
        Ly = 1.0 if Ly is None else Ly
        Lz = (nnodes_z - 1) / (nnodes_y - 1) if Lz is None else Lz
        x = np.linspace(0, Lx, nnodes_x)
        y = np.linspace(0, Ly, nnodes_y)
        z = np.linspace(0, Lz, nnodes_z)
        nodes3d = np.arange(nnodes_x * nnodes_y * nnodes_z).reshape(nnodes_z, nnodes_y, nnodes_x)
        X = np.empty((nodes3d.size, 3))
        X[:, 0] = np.broadcast_to(x[None, None, :], nodes3d.shape).ravel()
        X[:, 1] = np.broadcast_to(y[None, :, None], nodes3d.shape).ravel()
        X[:, 2] = np.broadcast_to(z[:, None, None], nodes3d.shape).ravel()
        if element_type == "quad":
            n = nodes3d[0]
            corners = [n[:-1, :-1], n[:-1, 1:], n[1:, 1:], n[1:, :-1]]
        else:
            lo, hi = nodes3d[:-1], nodes3d[1:]
            corners = [lo[:, :-1, :-1], lo[:, :-1, 1:], lo[:, 1:, 1:], lo[:, 1:, :-1],
                       hi[:, :-1, :-1], hi[:, :-1, 1:], hi[:, 1:, 1:], hi[:, 1:, :-1]]
        conn = np.stack([c.ravel() for c in corners], axis=1).astype(int)
        self.nnodes_x, self.nnodes_y, self.nnodes_z = nnodes_x, nnodes_y, nnodes_z
        self.nnodes = nodes3d.size
        self.nodes3d = nodes3d
        self.conn = conn
        self.X = X[:, : self.ndims]

    def create_poisson_problem(self):
        """Fix the x = 0 face (pyfem.py:2727-2734)."""
        dof_fixed = [int(v) for v in self.nodes3d[:, :, 0].ravel()]
        return self.conn, self.X, dof_fixed

    def create_linear_elasticity_problem(self):
        """Clamp the x = 0 face, unit downward load on the y = 0, x = Lx edge (pyfem.py:2736-2755)."""
        face = self.nodes3d[:, :, 0].ravel()
        dof_fixed = [int(v) for v in (self.ndims * face[:, None] + np.arange(self.ndims)[None, :]).ravel()]
        nodal_force = {int(self.nodes3d[k, 0, -1]): [0.0, -1.0, 0.0][: self.ndims] for k in range(self.nnodes_z)}
        return self.conn, self.X, dof_fixed, nodal_force

    def create_helmhotz_problem(self):
        """Raw design field: 0.95 in the low corner, 1e-3 elsewhere (pyfem.py:2757-2773)."""
        k, j, i = np.meshgrid(np.arange(self.nnodes_z), np.arange(self.nnodes_y), np.arange(self.nnodes_x),
                              indexing="ij")
        low = (i < self.nnodes_x / 2) & (j < self.nnodes_y / 2) & (k < self.nnodes_z / 2)
        x = np.where(low, 0.95, 1e-3).ravel()
        return self.conn, self.X, x
