"""Assembler: the callers of the assembly path (pyfem.py:2286-2423) -- linear solve and Newton loop.

By default the assembled scipy CSR goes to scipy's direct solver, or to cg / gmres (preconditioned with pyamg
smoothed aggregation when pyamg is installed, as in the reference, else unpreconditioned).  `solve(method="cg",
device=True)` is the GPU consumer of SURVEY.md section 8f #4: boundary conditions and conjugate gradients run on
the device CSR, so the matrix never crosses PCIe.
"""
import numpy as np
from scipy.sparse.linalg import cg, gmres, spsolve


class Assembler:
    def __init__(self, model):
        self.model = model

    def solve(self, method="gmres", device=False):
        """Static analysis (pyfem.py:2299-2317).  device=True (with method="cg"): assembly, Dirichlet conditions and
        a Jacobi-preconditioned CG all run on the device CSR -- K is never copied to the host."""
        assert method in ("direct", "cg", "gmres")
        if device:
            if method != "cg":
                raise NotImplementedError("the device solve is conjugate gradients: use method='cg'")
            vals = self.model.compute_jacobian_device()
            u, _, self.last_iterations = self.model.solve_device(vals, self.model.compute_rhs())
            return u.cpu().numpy()
        K = self.model.compute_jacobian()
        rhs = self.model.compute_rhs()
        K, rhs = self.model.apply_dirichlet_bcs(K, rhs, enforce_symmetric_K=True)
        return self._solve_linear_system(K, rhs, method)

    def solve_nonlinear(self, method="gmres", xdv=None, u0=None, tol=1e-10, atol=1e-12, max_iter=10, device=False):
        """Newton iteration with Jacobian + residual re-assembled every step (pyfem.py:2319-2355).  device=True: the
        whole loop stays in HBM -- one fused assembly of Jacobian and residual, boundary conditions on the device CSR
        (pattern kept), the step solved by Jacobi-preconditioned BiCGStab; only the residual norm crosses PCIe."""
        assert method in ("direct", "cg", "gmres")
        if device:
            return self._solve_nonlinear_device(xdv, u0, tol, atol, max_iter)
        u = np.zeros(self.model.nnodes) if u0 is None else u0
        res_norm_init = None
        for k in range(max_iter):
            K = self.model.compute_jacobian(xdv, u)
            res = self.model.compute_rhs(xdv, u)
            self.model.apply_dirichlet_bcs(K, res, enforce_symmetric_K=False)
            res_norm = np.sqrt(np.dot(res, res))
            print("pyfem", "{0:5d} {1:25.15e}".format(k, res_norm))
            if k == 0:
                res_norm_init = res_norm
            elif res_norm < tol * res_norm_init or res_norm < atol:
                break
            u -= self._solve_linear_system(K, res, method)
        return u

    def _solve_nonlinear_device(self, xdv, u0, tol, atol, max_iter):
        import torch
        model, mesh = self.model, self.model.mesh
        # one rank of a row-slab partition: u is kept on the global numbering (the assembly reads the rank's local
        # nodes from it), the rank updates its own rows and fetches the ghost entries from their owners; the step
        # comes from the BiCGStab over all ranks and the residual norm is the global one.  Returns the rank's rows.
        solver = model._slab_solver() if (model.slab is not None and model.slab.size > 1) else None
        u = torch.zeros(model.nnodes, dtype=torch.float64, device=mesh.device) if u0 is None else \
            torch.as_tensor(u0, dtype=torch.float64).to(mesh.device).clone()
        if u.numel() != model.nnodes:
            raise ValueError(f"u0 must have {model.nnodes} entries (the global mesh)")
        own = slice(solver.row0, solver.row0 + mesh.nrows) if solver is not None else slice(None)
        res_norm_init = None
        self.last_iterations = []
        for k in range(max_iter):
            K, res = model.assemble_device(xdv, u)  # Jacobian and residual from one pass over the elements
            mesh.apply_dirichlet(K, res, model.dof_fixed, None, enforce_symmetric=False)
            rr = torch.dot(res, res).reshape(1)
            if solver is not None:
                solver.exchange.all_reduce(rr)
            res_norm = float(torch.sqrt(rr))
            print("pyfem", "{0:5d} {1:25.15e}".format(k, res_norm))
            if k == 0:
                res_norm_init = res_norm
            elif res_norm < tol * res_norm_init or res_norm < atol:
                break
            du, iters, _ = (solver if solver is not None else mesh).bicgstab(K, res, rtol=1e-8)
            self.last_iterations.append(iters)
            u[own] -= du
            if solver is not None:
                solver.exchange.refresh(u)
        return u[own].cpu().numpy()

    def _setup_amg(self, K):
        try:
            import pyamg
        except ImportError:
            return None
        return pyamg.smoothed_aggregation_solver(K).aspreconditioner()

    def _solve_linear_system(self, K, rhs, method):
        if method == "direct":
            return spsolve(K, rhs)
        M = self._setup_amg(K)
        solver = cg if method == "cg" else gmres
        u, fail = solver(K, rhs, rtol=1e-8, M=M, atol=0.0)
        if fail:
            raise RuntimeError(f"{method} failed with code {fail}")
        return u
