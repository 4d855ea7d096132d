"""Physics models with the reference's constructor / method surface (pyfem.py:634-2177), assembling on
the GPU.  `compute_jacobian` returns the same scipy CSR matrix (pattern bit-exact, values to 1e-12 of
max|K|) and `compute_rhs` the same persistent rhs vector as the reference; `*_device` variants keep the
result in HBM for GPU consumers.

What stays on the host is glue the reference also does in Python: argument casting, the persistent
`rhs` array, point loads, Dirichlet conditions on a host CSR.  There is no CPU assembly path: element
families / dtypes without a kernel raise NotImplementedError.
"""
import numpy as np
from scipy import sparse

from .engine import DeviceMesh, _torch
from .fem import BasisBase, QuadratureBase


class ModelBase:
    """Mesh arrays, dof maps, device handle and the scatter (pyfem.py:634-931)."""

    def __init__(self, ndof_per_node, X, conn, dof_fixed, dof_fixed_vals, quadrature: QuadratureBase,
                 basis: BasisBase, device=None, scatter="auto", group=None, partition=None, node_ranges=None,
                 halo="ghost"):
        """Reference arguments (pyfem.py:640-657) plus, keyword-only in practice:
          device   torch device of the handle (default: the current CUDA device)
          scatter  "auto" | "gather" | "atomic"
          group    a torch.distributed process group: the model becomes ONE RANK of a row-slab partition of the
                   (global) mesh passed in X / conn -- compute_jacobian returns the rank's row slab (scipy CSR,
                   global columns, the reference's index dtype), gather() rebuilds the global matrix on rank 0
          partition / node_ranges   a ready-made partition.LocalMesh (e.g. structured_slab) and every rank's global
                   node range, instead of splitting X / conn here
          halo     "ghost": ghost-element layer, no exchange | "nccl" / "p2p": every element integrated once,
                   interface rows summed over NCCL send/recv or through NVLink peer stores (halo.py)
        """
        self.ndof_per_node = int(ndof_per_node)
        self.X = np.array(X, dtype=float)
        self.conn = np.array(conn, dtype=int)
        self.dof_fixed = np.array(dof_fixed, dtype=int)
        self.dof_fixed_vals = None if dof_fixed_vals is None else np.array(dof_fixed_vals, dtype=float)
        self.quadrature = quadrature
        self.basis = basis
        self.scatter = scatter

        self.nelems, self.nnodes_per_elem = self.conn.shape
        self.nnodes, self.ndims = self.X.shape
        self.nquads = quadrature.get_nquads()
        if getattr(quadrature, "device_elem", None) != self.nnodes_per_elem or \
                getattr(basis, "device_elem", None) != self.nnodes_per_elem:
            raise NotImplementedError("quadrature / basis / connectivity combination has no device path "
                                      "(bilinear quad4 and trilinear hex8 with their 2-point Gauss rules only)")
        self.nodes = np.arange(self.nnodes)
        self.ndof = self.nnodes * self.ndof_per_node

        # device: H2D of X / conn, the conn.min()/max() asserts of pyfem.py:680-681, CSR pattern, plans
        self.slab = self._reducer = self._slab_cg = None
        if group is not None or partition is not None:
            self._init_slab(group, partition, node_ranges, halo, device)
        else:
            self.mesh = DeviceMesh(self.X, self.conn, self.ndof_per_node, device=device)
        self._asm = self._reducer if self._reducer is not None else self.mesh

        # persistent right-hand side (pyfem.py:756; survey trap T7)
        self.rhs = np.zeros(self.ndof)
        self._dof = self._dof_each_node = self._conn_dof = self._dof_free = None

    # ---- one rank of a row-slab partition --------------------------------------------------------------
    def _init_slab(self, group, partition, node_ranges, halo, device):
        from .partition import SlabContext
        if halo not in ("ghost", "nccl", "p2p"):
            raise ValueError(f"unknown halo variant {halo!r}")
        if partition is not None and partition.nnodes_global != self.nnodes:
            # X / conn may be the rank's local arrays when the global mesh was never built (bench.py at full size)
            self.nnodes, self.ndof = partition.nnodes_global, partition.nnodes_global * self.ndof_per_node
            self.nodes = np.arange(self.nnodes)
        self.slab = SlabContext(self.X, self.conn, group=group, partition=partition, ranges=node_ranges)
        part = self.slab.part
        nel_global = int(part.nelems_global) if getattr(part, "nelems_global", None) else None
        if self.slab.size == 1:  # a partition of one: the plain single-GPU handle
            self.mesh = DeviceMesh(part.X, part.conn, self.ndof_per_node, device=device)
        elif halo == "ghost":
            self.mesh = DeviceMesh(part.X, part.conn, self.ndof_per_node, device=device, own_range=part.own_range,
                                   node_gid=part.node_gid, ncols_nodes=part.nnodes_global, nelems_global=nel_global)
        else:
            from .halo import ReduceAssembler
            self._reducer = ReduceAssembler(part, self.ndof_per_node, self.slab.ranges, device=device, group=group,
                                            transport=halo, nelems_global=nel_global)
            self.mesh = self._reducer.mesh
        self._row0 = self.slab.owned_nodes[0] * self.ndof_per_node  # global dof of the slab's first row

    def _local(self, f):
        """Nodal field as the handle wants it: global -> this rank's local nodes in slab mode."""
        return f if self.slab is None else self.slab.local_field(f)

    def gather(self, K_slab, dst=0):
        """Slab mode: the global CSR (the reference's matrix) on rank `dst`, None elsewhere."""
        return K_slab if self.slab is None else self.slab.gather_matrix(K_slab, dst)

    def gather_vector(self, v_owned, dst=0):
        return v_owned if self.slab is None else self.slab.gather_vector(v_owned, dst)

    def _owned(self, vec_global):
        """The owned rows of a global dof vector (slab mode), the vector itself otherwise."""
        if self.slab is None:
            return vec_global
        m = self.ndof_per_node
        b, e = self.slab.owned_nodes
        return vec_global[b * m: e * m]

    # ---- dof maps (utils.create_dof, utils.py:267-298), built on first use ---------------------------
    def _make_dof(self):
        m = self.ndof_per_node
        if m == 1:
            self._dof, self._dof_each_node, self._conn_dof = self.nodes, self.nodes, self.conn
        else:
            self._dof = np.arange(self.ndof)
            self._dof_each_node = self._dof.reshape(self.nnodes, m)
            self._conn_dof = (m * self.conn[:, :, None] + np.arange(m)[None, None, :]).reshape(self.nelems, -1)

    @property
    def dof(self):
        if self._dof is None:
            self._make_dof()
        return self._dof

    @property
    def dof_each_node(self):
        if self._dof_each_node is None:
            self._make_dof()
        return self._dof_each_node

    @property
    def conn_dof(self):
        if self._conn_dof is None:
            self._make_dof()
        return self._conn_dof

    @property
    def dof_free(self):
        if self._dof_free is None:
            mask = np.ones(self.ndof, dtype=bool)
            mask[self.dof_fixed] = False
            self._dof_free = np.nonzero(mask)[0]  # == np.setdiff1d(dof, dof_fixed), pyfem.py:699
        return self._dof_free

    # ---- abstract interface --------------------------------------------------------------------------
    def compute_rhs(self):
        return self.rhs

    def compute_jacobian(self):
        raise NotImplementedError

    # ---- the reference's internal two-step interface (element matrices, then scatter) ---------------------
    @property
    def Ke_mat(self):
        """(nelems, D, D) element-matrix buffer of the reference (pyfem.py:721), allocated on first use: the fused
        path never needs it, but scripts such as performance_test.py:52 and plugin back-ends in the style of
        A2DWrapper (pyfem.py:2255-2277) fill it and hand it to _assemble_jacobian."""
        if getattr(self, "_Ke_mat", None) is None:
            D = self.nnodes_per_elem * self.ndof_per_node
            self._Ke_mat = np.zeros((self.nelems, D, D))
        return self._Ke_mat

    def _assemble_jacobian(self, Ke_mat):
        """Scatter caller-supplied element matrices into the global CSR (pyfem.py:920-931)."""
        return self._to_scipy(self.mesh.scatter_matrix(Ke_mat, mode=self.scatter))

    def _assemble_rhs(self, rhs_e, rhs):
        """rhs[conn] += rhs_e, after zeroing rhs (pyfem.py:860-875); scalar models."""
        rhs[:] = self.mesh.scatter_vector(rhs_e, mode=self.scatter).cpu().numpy()
        return rhs

    def _element_matrices(self, physics, out, **kw):
        Ke, _, _ = self.mesh.element_matrices(physics, **kw)
        if out is not None:
            out[...] = Ke.cpu().numpy()
            return out
        return Ke

    # ---- helpers -------------------------------------------------------------------------------------
    def _to_scipy(self, vals):
        # values land in a pinned buffer of the handle's pool (reused once the previous matrix is gone); the cached
        # indptr / indices are shared read-only and copied on write by apply_dirichlet_bcs
        return self.mesh.to_scipy(vals, copy_pattern=False, reuse_host_buffers=True)

    def _vec_to_host(self, vec, out):
        self._owned(out)[:] = vec.cpu().numpy()
        return self._owned(out)

    def solve_device(self, vals, rhs, rtol=1e-8, atol=0.0, max_iter=None):
        """Dirichlet conditions + conjugate gradients entirely in HBM: `vals` are the device
        CSR values of compute_jacobian_device (edited in place, pattern kept), `rhs` a host or device vector.
        Returns (u, rhs with the conditions applied, CG iterations), u and rhs as device tensors.  The matrix is
        never copied to the host."""
        torch = _torch()
        rhs_d = torch.as_tensor(rhs).to(device=self.mesh.device, dtype=torch.float64).clone()
        self.mesh.apply_dirichlet(vals, rhs_d, self.dof_fixed, self.dof_fixed_vals, enforce_symmetric=True)
        if self.slab is not None and self.slab.size > 1:
            # one rank of a row-slab partition: `rhs` / `u` are the rank's rows, the solve runs over all ranks
            # (slab_solve.py: halo exchange of the search direction + three all-reduced scalars per iteration)
            u, iters, _ = self._slab_solver().cg(vals, rhs_d, rtol=rtol, atol=atol, max_iter=max_iter)
        else:
            u, iters, _ = self.mesh.cg(vals, rhs_d, rtol=rtol, atol=atol, max_iter=max_iter)
        return u, rhs_d, iters

    def _slab_solver(self):
        if self._slab_cg is None:
            from .slab_solve import SlabKrylov
            self._slab_cg = SlabKrylov(self.mesh, self.slab.part, self.slab.ranges, self.slab.rank, self.slab.group)
        return self._slab_cg

    def _local_dofs(self, v):
        """A dof vector on this rank's LOCAL nodes (owned + ghost) from the global vector or from the rank's own rows
        (then the ghost entries come from their owners: one halo exchange, slab_solve.HaloExchange)."""
        torch = _torch()
        m = self.ndof_per_node
        part = self.slab.part
        dof = torch.as_tensor((np.asarray(part.node_gid)[:, None] * m + np.arange(m)).ravel(), device=self.mesh.device)
        v = torch.as_tensor(np.asarray(v) if not hasattr(v, "device") else v).to(device=self.mesh.device, dtype=torch.float64)
        if v.numel() == self.ndof:
            return v[dof]
        if v.numel() != self.mesh.nrows:
            raise ValueError(f"dof vector has {v.numel()} entries, expected {self.ndof} (global) or {self.mesh.nrows} "
                             "(this rank's rows)")
        return self._global_dofs(v)[dof]

    def _global_dofs(self, v_owned):
        """The rank's rows placed in a global-length device vector with the ghost entries fetched from their owners
        (what a slab product reads); other entries are zero."""
        torch = _torch()
        solver = self._slab_solver()
        v = torch.as_tensor(v_owned).to(device=self.mesh.device, dtype=torch.float64)
        full = torch.zeros(self.ndof, dtype=torch.float64, device=self.mesh.device)
        full[solver.row0: solver.row0 + v.numel()] = v
        return solver.exchange.refresh(full)

    def _global_sum(self, v):
        """Sum of a per-rank scalar over the ranks of a slab partition (the value itself otherwise)."""
        if self.slab is None or self.slab.size == 1:
            return v
        import torch.distributed as dist
        on_host = dist.get_backend(self.slab.group) == "gloo"  # gloo reduces host memory
        t = _torch().tensor([float(v)], dtype=_torch().float64, device="cpu" if on_host else self.mesh.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.slab.group)
        return float(t.item())

    def apply_dirichlet_bcs(self, K, rhs, enforce_symmetric_K=True):
        """Host-side Dirichlet conditions with the reference's semantics (pyfem.py:780-835): rows (and, if
        asked, columns) of fixed dofs zeroed, unit diagonal, explicit zeros eliminated, rhs updated.
        K and rhs are edited in place (Assembler.solve_nonlinear relies on that, pyfem.py:2343) and returned.
        """
        if getattr(K, "_pfg_shared_pattern", False):  # copy-on-write: eliminate_zeros below edits the pattern
            K.indices, K.indptr = K.indices.copy(), K.indptr.copy()
            K._pfg_shared_pattern = False
        row0 = getattr(self, "_row0", 0) if getattr(self, "slab", None) is not None else 0  # slab: rows row0 ...
        nrows = K.shape[0]
        is_fixed = np.zeros(K.shape[1], dtype=bool)
        is_fixed[self.dof_fixed] = True
        row_fixed = is_fixed[row0: row0 + nrows]
        free_rows = np.nonzero(~row_fixed)[0]
        Krb_u0 = None
        u0 = None
        if self.dof_fixed_vals is not None:
            u0 = np.zeros(K.shape[1])
            u0[self.dof_fixed] = self.dof_fixed_vals
            if enforce_symmetric_K:
                Krb_u0 = (K @ u0)[free_rows] - 0.0  # only fixed columns carry non-zero u0
        diag = K.diagonal(k=row0)
        row_of = np.repeat(np.arange(nrows), np.diff(K.indptr))
        K.data[row_fixed[row_of]] = 0.0
        if enforce_symmetric_K:
            K.data[is_fixed[K.indices]] = 0.0
        diag[row_fixed] = 1.0
        K.setdiag(diag, k=row0)
        K.eliminate_zeros()
        if rhs.shape[0] != nrows:  # slab mode with the global persistent rhs: edit the owned rows
            rhs = rhs[row0: row0 + nrows]
        if self.dof_fixed_vals is None:
            rhs[row_fixed] = 0.0
        else:
            rhs[row_fixed] = u0[row0: row0 + nrows][row_fixed]
            if enforce_symmetric_K:
                rhs[free_rows] -= Krb_u0
        return K, rhs


class _DensityFunctions:
    """compliance / volume and their gradients (LinearPoisson pyfem.py:1033-1123, LinearElasticity :1796-1870).
    The solves stay on the host, as in the reference (the device CG of Assembler.solve(device=True) is the
    GPU consumer of SURVEY 8f #4)."""

    def _solve_compliance_system(self, rho, solver, device=False):
        assert solver == "direct" or solver == "cg" or solver == "gmres"
        if device:
            # assembly, boundary conditions and the solve stay in HBM; only rhs and u cross PCIe
            if solver != "cg":
                raise NotImplementedError("the device solve is conjugate gradients: use solver='cg'")
            rhs = self.compute_rhs()
            u, rhs_d, _ = self.solve_device(self.compute_jacobian_device(rho), rhs)
            rhs[:] = rhs_d.cpu().numpy()  # the persistent rhs is edited in place, as apply_dirichlet_bcs does
            return None, rhs, u.cpu().numpy()
        K = self.compute_jacobian(rho)
        rhs = self.compute_rhs()
        K, rhs = self.apply_dirichlet_bcs(K, rhs, enforce_symmetric_K=True)
        if solver == "direct":
            from scipy.sparse.linalg import spsolve
            return K, rhs, spsolve(K.tocsc(), rhs)
        # cg / gmres as in the reference (pyfem.py:1056-1068): pyamg smoothed aggregation as the preconditioner when
        # pyamg is installed, unpreconditioned otherwise (same rule as Assembler._solve_linear_system)
        from scipy.sparse.linalg import cg, gmres
        try:
            import pyamg
            M = pyamg.smoothed_aggregation_solver(K).aspreconditioner()
        except ImportError:
            M = None
        u, fail = (cg if solver == "cg" else gmres)(K, rhs, rtol=1e-8, atol=0.0, M=M)
        if fail:
            raise RuntimeError(f"{solver} failed with code {fail}")
        return K, rhs, u

    def volume(self, rho):
        return np.asarray(rho).sum() / self.nnodes

    def volume_grad(self, rho):
        return np.ones(self.nnodes) / self.nnodes

    deterministic_sens = False  # True: plan-ordered nodal sums (bitwise reproducible) instead of atomic adds

    def _sens_mesh(self):
        """The handle of the deterministic sensitivity pass: a scalar handle (one dof row per node), whose tile plan
        carries the nodal-vector codes.  For elasticity it is a second handle of the same mesh, built on first use."""
        if self.ndof_per_node == 1:
            return self.mesh
        if getattr(self, "_scalar_mesh", None) is None:
            self._scalar_mesh = DeviceMesh(self.X, self.conn, 1, device=self.mesh.device)
        return self._scalar_mesh

    def _k_dv_sens_slab(self, physics, rho, phi, psi, **kw):
        """Sensitivities at this rank's owned nodes: every element touching an owned node is local (ghost layer), so
        no value crosses ranks; phi / psi may be global vectors or the rank's rows (e.g. the u of solve_device)."""
        if self.deterministic_sens:
            raise NotImplementedError("the plan-ordered sensitivity pass is single-GPU; slab models use the atomic pass")
        rho_l = self._local(np.ones(self.nnodes) * rho if not hasattr(rho, "__len__") else rho)
        phi_l, psi_l = self._local_dofs(phi), self._local_dofs(psi)
        if self._reducer is not None:  # the reduce variant masks the ghost elements: integrate them here too
            self.mesh.set_element_mask(None)
        try:
            return self.mesh.k_dv_sens(physics, rho_l, self.p, phi_l, psi_l, **kw).cpu().numpy()
        finally:
            if self._reducer is not None:
                self.mesh.set_element_mask(self._reducer.plan.skip_mask)

    def _k_dv_sens(self, physics, rho, phi, psi, **kw):
        _check_real(rho)
        if self.slab is not None and self.slab.size > 1:
            return self._k_dv_sens_slab(physics, rho, phi, psi, **kw)
        rho = np.ones(self.nnodes) * rho if not hasattr(rho, "__len__") else rho
        mesh = self._sens_mesh() if self.deterministic_sens else self.mesh
        return mesh.k_dv_sens(physics, rho, self.p, phi, psi, deterministic=self.deterministic_sens, **kw).cpu().numpy()


def _is_complex(rho):
    return bool(rho.is_complex() if hasattr(rho, "is_complex") else np.iscomplexobj(rho))  # torch tensor or numpy


def _check_real(rho):
    if _is_complex(rho):
        raise NotImplementedError("complex rho (complex-step verification, pyfem.py:1019-1020) has no device "
                                  "path and this engine has no CPU fallback")


class LinearPoisson(_DensityFunctions, ModelBase):
    """-k lap(u) = g with RAMP-penalised conductivity (pyfem.py:934-1329)."""

    def __init__(self, X, conn, dof_fixed, dof_fixed_vals, quadrature, basis, gfunc, kappa0=1.0, p=0.0, **kw):
        super().__init__(1, X, conn, dof_fixed, dof_fixed_vals, quadrature, basis, **kw)
        self.gfunc = gfunc
        self.kappa0 = kappa0  # stored and, as in the reference, never used (pyfem.py:977; survey trap T5)
        self.p = p

    def compute_jacobian_device(self, rho=1.0, out=None):
        """CSR values on the device.  A complex rho (the reference's complex-step checks, pyfem.py:1018-1020) yields
        complex128 values: Re K and Im K from two real assemblies (single-GPU handles)."""
        if _is_complex(rho) and self._reducer is not None:
            raise NotImplementedError("complex rho with the reduce variant of a slab partition")
        return self._asm.assemble_poisson(self._local(rho), self.p, out=out, mode=self.scatter)

    def compute_jacobian(self, rho=1.0):
        return self._to_scipy(self.compute_jacobian_device(rho))

    def _compute_element_jacobian(self, Ke_mat, rho=1.0):
        """Element matrices only (pyfem.py:1188-1217); the reference reads rho from the last material update."""
        rho_t, rho_c = self.mesh._rho(rho)
        return self._element_matrices("poisson", Ke_mat, field=rho_t, field_const=rho_c, params=(self.p,))

    def _source_at_quads(self):
        """g(x_q): the user callable runs on a CUDA tensor of quadrature coordinates; callables that need
        numpy get a host array (the callable is user code, not part of the assembly path)."""
        torch = _torch()
        Xq = self.mesh.quad_points()
        try:
            g = self.gfunc(Xq)
        except (TypeError, RuntimeError, AttributeError):
            g = self.gfunc(Xq.cpu().numpy())
        g = torch.as_tensor(g, dtype=torch.float64, device=self.mesh.device)
        return g.expand(Xq.shape[:-1]).contiguous() if g.dim() == 0 or g.shape != Xq.shape[:-1] else g.contiguous()

    def compute_rhs_device(self, out=None):
        if self._reducer is not None:  # the slab's ghost elements complete the owned rows: integrate them here too
            self.mesh.set_element_mask(None)
        try:
            return self.mesh.poisson_rhs(self._source_at_quads(), out=out, mode=self.scatter)
        finally:
            if self._reducer is not None:
                self.mesh.set_element_mask(self._reducer.plan.skip_mask)

    def compute_rhs(self):
        return self._vec_to_host(self.compute_rhs_device(), self.rhs)


    def _compute_K_dv_sens(self, rho, phi, psi):
        """d(phi^T K psi)/d rho (pyfem.py:1239-1276)."""
        return self._k_dv_sens("poisson", rho, phi, psi)

    def compliance(self, rho, solver="cg", weighted=True, device=False):
        """Thermal compliance and the solution (pyfem.py:1033-1073).  device=True: Dirichlet conditions and a
        Jacobi-preconditioned CG run on the device CSR (homogeneous or symmetric-eliminated conditions)."""
        _, rhs, u = self._solve_compliance_system(rho, solver, device)
        if weighted:
            return self._global_sum(rhs.dot(u)), u  # slab mode: u and rhs are the rank's rows
        return self._global_sum(np.sum(u)) / self.ndof, u

    def compliance_grad(self, rho, u, weighted=True):
        """pyfem.py:1075-1101."""
        if weighted:
            psi = u
        else:
            from scipy.sparse.linalg import spsolve
            K = self.compute_jacobian(rho)
            K, rhs = self.apply_dirichlet_bcs(K, np.ones(len(u)), enforce_symmetric_K=True)
            psi = spsolve(K.tocsc(), rhs) / len(u)
        return -self._compute_K_dv_sens(rho, psi, u)


class NonlinearPoisson2D(ModelBase):
    """-div(h(x)(1+u^2) grad u) = g on quad4 meshes (pyfem.py:1332-1664)."""

    def __init__(self, X, conn, dof_fixed, dof_fixed_vals, quadrature, basis, **kw):
        super().__init__(1, X, conn, dof_fixed, dof_fixed_vals, quadrature, basis, **kw)
        if self.nnodes_per_elem != 4:
            raise NotImplementedError("NonlinearPoisson2D is a quad4 model")

    def assemble_device(self, xdv, u, want_K=True, want_res=True):
        """Jacobian values and residual from ONE pass over the elements (device tensors)."""
        if self._reducer is not None:  # the reduce variant ships Jacobian and residual shares together
            return self._reducer.assemble_nlpoisson(xdv, self._local(u), mode=self.scatter)
        return self.mesh.assemble_nlpoisson(xdv, self._local(u), want_K=want_K, want_res=want_res, mode=self.scatter)

    def compute_jacobian(self, xdv, u):
        K, _ = self.assemble_device(xdv, u, want_K=True, want_res=False)
        return self._to_scipy(K)

    def compute_rhs(self, xdv, u):
        """The residual, as in the reference (pyfem.py:1375-1388)."""
        _, res = self.assemble_device(xdv, u, want_K=False, want_res=True)
        return self._vec_to_host(res, self.rhs)


class LinearElasticity(_DensityFunctions, ModelBase):
    """Linear elasticity, plane stress in 2-D (pyfem.py:1667-2068)."""

    def __init__(self, X, conn, dof_fixed, dof_fixed_vals, nodal_force, quadrature, basis, E=10.0, nu=0.3, p=0.0,
                 **kw):
        super().__init__(np.asarray(X).shape[1], X, conn, dof_fixed, dof_fixed_vals, quadrature, basis, **kw)
        self.nodal_force = nodal_force
        self.E, self.nu, self.p = E, nu, p
        if self.ndims == 2:  # pyfem.py:1746-1750
            self.C0 = E / (1.0 - nu ** 2) * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, 0.5 * (1.0 - nu)]])
        else:  # pyfem.py:1752-1757
            self.C0 = np.zeros((6, 6))
            self.C0[:3, :3] = nu
            self.C0[np.arange(3), np.arange(3)] = 1.0 - nu
            self.C0[np.arange(3, 6), np.arange(3, 6)] = 0.5 - nu
            self.C0 *= E / ((1.0 + nu) * (1.0 - 2.0 * nu))

    def compute_rhs(self):
        """Point loads assigned (not added) into the persistent rhs (pyfem.py:1760-1768)."""
        nodes = np.array(list(self.nodal_force.keys()), dtype=int)
        self.rhs[self.dof_each_node[nodes].flatten()] = np.array(list(self.nodal_force.values())).flatten()
        return self._owned(self.rhs)

    def compute_jacobian_device(self, rho=1.0, out=None):
        """CSR values on the device; complex rho (complex-step checks, pyfem.py:1783-1785) yields complex128 values."""
        if _is_complex(rho) and self._reducer is not None:
            raise NotImplementedError("complex rho with the reduce variant of a slab partition")
        return self._asm.assemble_elasticity(self._local(rho), self.p, self.E, self.nu, out=out, mode=self.scatter)

    def _compute_element_jacobian(self, Ke_mat, rho=1.0):
        """Element matrices only (pyfem.py:2029-2068)."""
        rho_t, rho_c = self.mesh._rho(rho)
        return self._element_matrices("elasticity", Ke_mat, field=rho_t, field_const=rho_c,
                                      params=(self.p, self.E, self.nu))

    def compute_jacobian(self, rho=1.0):
        return self._to_scipy(self.compute_jacobian_device(rho))


    def _compute_K_dv_sens(self, rho, phi, psi):
        """d(phi^T K psi)/d rho (pyfem.py:1872-1920)."""
        return self._k_dv_sens("elasticity", rho, phi, psi, E=self.E, nu=self.nu)

    def compliance(self, rho, solver="cg", device=False):
        """Compliance and the solution (pyfem.py:1796-1833); device=True keeps the system and the CG solve in HBM."""
        _, rhs, u = self._solve_compliance_system(rho, solver, device)
        return self._global_sum(rhs.dot(u)), u  # slab mode: u and rhs are the rank's rows

    def compliance_grad(self, rho, u):
        """pyfem.py:1835-1847."""
        return -self._compute_K_dv_sens(rho, u, u)


class Helmholtz(ModelBase):
    """Helmholtz filter -r^2 lap(rho) + rho = x (pyfem.py:2071-2177): K and R assembled once per mesh."""

    def __init__(self, r0, X, conn, quadrature, basis, **kw):
        super().__init__(1, X, conn, [], None, quadrature, basis, **kw)
        self.r0 = r0
        self.K_device, self.R_device = self._asm.assemble_helmholtz(r0, mode=self.scatter)
        self.R = self._to_scipy(self.R_device)
        self.RT = self.R.transpose()
        self.K = self._to_scipy(self.K_device)
        self._Ksolve = None

    @property
    def Ksolve(self):
        # The reference builds a pyamg Ruge-Stuben hierarchy here (pyfem.py:2098); the solver is outside
        # the assembly path, so a sparse factorisation with the same .solve(b, tol) call stands in.
        if self._Ksolve is None:
            from scipy.sparse.linalg import splu
            lu = splu(self.K.tocsc())

            class _Solve:
                def solve(self, b, tol=1e-8):
                    return lu.solve(np.asarray(b, dtype=float))
            self._Ksolve = _Solve()
        return self._Ksolve

    def apply(self, x):
        return self.Ksolve.solve(self.compute_rhs(x), tol=1e-8)

    def apply_gradient(self, gradrho):
        return self.RT.dot(self.Ksolve.solve(gradrho, tol=1e-8))

    def compute_rhs(self, x):
        self._owned(self.rhs)[:] = self.R.dot(x)  # slab mode: x is the global field, the result the owned rows
        return self._owned(self.rhs)

    def compute_rhs_device(self, x, out=None):
        return self.mesh.spmv(self.R_device, x, out=out)

    def _cg_device(self, b, rtol):
        if self.slab is not None and self.slab.size > 1:
            return self._slab_solver().cg(self.K_device, b, rtol=rtol)[0]
        return self.mesh.cg(self.K_device, b, rtol=rtol)[0]

    def apply_device(self, x, rtol=1e-8):
        """Filtered field K^-1 R x with R.x and the CG solve on the device (pyfem.py:2102-2107).  Slab mode: x is the
        global field, the result the rank's rows (the solve runs over all ranks)."""
        return self._cg_device(self.compute_rhs_device(x), rtol)

    def apply_gradient_device(self, gradrho, rtol=1e-8):
        """R^T K^-1 g on the device (pyfem.py:2109-2115): CG solve, then the transposed product.  Slab mode: g and the
        result are the rank's rows; a slab does not hold the transposed entries, so the product uses R itself with the
        ghost entries of K^-1 g fetched from their owners (R is the mass matrix: its (i, j) and (j, i) entries sum the
        same element integrals, in possibly different order)."""
        y = self._cg_device(gradrho, rtol)
        if self.slab is not None and self.slab.size > 1:
            return self.mesh.spmv(self.R_device, self._global_dofs(y))
        return self.mesh.spmv_t(self.R_device, y)

    def compute_jacobian(self):
        return self.K
