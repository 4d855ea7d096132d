"""DeviceMesh: the once-per-mesh device handle (pfg_mesh) with torch tensors as buffer plumbing.

torch supplies device memory, the current stream and host<->device copies; all arithmetic happens in
libpyfem_b200.so.  One DeviceMesh belongs to one GPU / rank.
"""
import ctypes
from ctypes import byref, c_double, c_int64, c_void_p

import numpy as np

from . import _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pyfem_gpu_testflight_b200 needs a CUDA device: the assembly path is "
                           "hand-written sm_100a CUDA and has no CPU fallback")
    return torch


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def index_bytes_rule(nelems, ndof_per_elem, ncols):
    """scipy's index width for coo -> csr (scipy/sparse/_coo.py:59-61, 419; SURVEY trap T2): int32 iff
    max(COO nnz = nelems * D^2, ncols) <= 2^31 - 1, else int64, for indptr AND indices.  For a rank's row slab the
    rule is evaluated with the GLOBAL element count, so that concatenated slabs carry the dtype of the reference's
    global matrix (256^3 hexes on 8 ranks: every slab is int64, as the global matrix is)."""
    return 4 if max(int(nelems) * int(ndof_per_elem) ** 2, int(ncols)) <= 2 ** 31 - 1 else 8


class DeviceMesh:
    """Device copies of the mesh, CSR pattern, element->slot map and gather plan (pfg_mesh_create).

    X (nnodes, ndims) float64; conn (nelems, nnodes_per_elem) integer; numpy arrays or torch tensors.
    own_range = (begin, end) node rows this handle assembles (default all); node_gid = strictly increasing
    local->global node ids for the reported column indices (multi-GPU slabs); nelems_global = element count of the
    whole mesh when this handle is one rank's slab (index_bytes_rule).
    """

    def __init__(self, X, conn, ndof_per_node, device=None, own_range=None, node_gid=None, ncols_nodes=None,
                 build_gather_plan=True, reorder=True, nelems_global=None):
        torch = _torch()
        self._lib = _lib.load()
        self._handle = None
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        Xd = torch.as_tensor(X).to(device=self.device, dtype=torch.float64).contiguous()
        cd = torch.as_tensor(conn).to(device=self.device, dtype=torch.int64).contiguous()
        if Xd.dim() != 2 or cd.dim() != 2:
            raise ValueError("X must be (nnodes, ndims) and conn (nelems, nnodes_per_elem)")
        self.nnodes, self.ndims = int(Xd.shape[0]), int(Xd.shape[1])
        self.nelems, self.nnodes_per_elem = int(cd.shape[0]), int(cd.shape[1])
        self.ndof_per_node = int(ndof_per_node)
        if (self.nnodes_per_elem, self.ndims) not in ((4, 2), (8, 3)):
            raise NotImplementedError(
                f"no device path for {self.nnodes_per_elem}-node elements in {self.ndims}-D "
                "(quad4 and hex8 only; no CPU fallback)")
        begin, end = (0, self.nnodes) if own_range is None else (int(own_range[0]), int(own_range[1]))
        gid = None
        if node_gid is not None:
            gid = torch.as_tensor(node_gid).to(device=self.device, dtype=torch.int64).contiguous()
            if ncols_nodes is None:
                raise ValueError("ncols_nodes (global node count) is required with node_gid")
        flags = (0 if build_gather_plan else _lib.CREATE_NO_GATHER_PLAN) | (0 if reorder else _lib.CREATE_NO_REORDER)
        handle = c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_mesh_create(
                byref(handle), self.nnodes_per_elem, self.ndof_per_node, self.nnodes, self.nelems, _ptr(Xd), _ptr(cd),
                begin, end, _ptr(gid), int(ncols_nodes or 0), flags, self._stream()))
        self._handle = handle
        self.own_begin, self.own_end = begin, end
        self.nnz = self.info(_lib.INFO_NNZ)
        self.nrows = self.info(_lib.INFO_NROWS)
        self.ncols = self.info(_lib.INFO_NCOLS)
        self.idx_bytes = self.info(_lib.INFO_IDX_BYTES)  # the rule applied to this handle's own element count
        if nelems_global is not None:
            self.idx_bytes = max(self.idx_bytes, index_bytes_rule(nelems_global, self.ndof_per_elem, self.ncols))
        self.nchunks = self.info(_lib.INFO_NCHUNKS)
        self.chunk_elems = self.info(_lib.INFO_CHUNK_ELEMS)
        self.plan_bytes = self.info(_lib.INFO_PLAN_BYTES)
        self.ntemplates = self.info(_lib.INFO_TEMPLATES)
        self._pattern = None
        self._pattern_host = None
        self._host_pool = []  # pinned host buffers for to_scipy(reuse_host_buffers=True)

    # ---- plumbing -------------------------------------------------------------------------------
    def _stream(self):
        torch = _torch()
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def info(self, what):
        v = c_int64()
        _lib.check(self._lib.pfg_mesh_get(self._handle, what, byref(v)))
        return int(v.value)

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.pfg_mesh_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def index_dtype(self):
        return np.int32 if self.idx_bytes == 4 else np.int64

    def new_values(self):
        torch = _torch()
        return torch.empty(self.nnz, dtype=torch.float64, device=self.device)

    def new_vector(self):
        torch = _torch()
        return torch.empty(self.nrows, dtype=torch.float64, device=self.device)

    def _dev_f64(self, a, n, what):
        torch = _torch()
        t = torch.as_tensor(a)
        if t.is_complex():
            raise NotImplementedError(f"complex {what} (complex-step verification) has no device path")
        t = t.to(device=self.device, dtype=torch.float64, non_blocking=True).contiguous()
        if t.numel() != n:
            raise ValueError(f"{what} must have {n} entries, got {t.numel()}")
        return t

    # ---- pattern --------------------------------------------------------------------------------
    def pattern(self, idx_bytes=None):
        """(indptr, indices) device tensors: K.indptr / K.indices of the reference's tocsr()."""
        torch = _torch()
        idx_bytes = idx_bytes or self.idx_bytes
        if self._pattern is not None and self._pattern[0] == idx_bytes:
            return self._pattern[1], self._pattern[2]
        dt = torch.int32 if idx_bytes == 4 else torch.int64
        indptr = torch.empty(self.nrows + 1, dtype=dt, device=self.device)
        indices = torch.empty(self.nnz, dtype=dt, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_mesh_pattern(self._handle, _ptr(indptr), _ptr(indices), idx_bytes, self._stream()))
        self._pattern = (idx_bytes, indptr, indices)
        return indptr, indices

    def pattern_host(self):
        if self._pattern_host is None:
            indptr, indices = self.pattern()
            self._pattern_host = (indptr.cpu().numpy(), indices.cpu().numpy())
            for a in self._pattern_host:  # shared with every matrix built by to_scipy(copy_pattern=False)
                a.flags.writeable = False
        return self._pattern_host

    # ---- assembly -------------------------------------------------------------------------------
    def _rho(self, rho):
        """nodal density tensor or (None, scalar) -- scalar rho is a constant field (pyfem.py:1015-1016)."""
        if rho is None:
            return None, 1.0
        if np.ndim(rho) == 0:  # Python / numpy scalars and 0-d arrays: the reference's `not hasattr(rho, "__len__")`
            if rho.is_complex() if hasattr(rho, "is_complex") else np.iscomplexobj(rho):
                raise NotImplementedError("complex rho (complex-step verification) has no device path")
            return None, float(rho)
        return self._dev_f64(rho, self.nnodes, "rho"), 0.0

    def _complex_parts(self, rho):
        """(real, imaginary) device tensors of a complex nodal density, or None for real input.  Complex scalars are
        constant fields (pyfem.py:1015-1016)."""
        torch = _torch()
        is_complex = rho.is_complex() if hasattr(rho, "is_complex") else np.iscomplexobj(rho)
        if not is_complex:
            return None
        if not hasattr(rho, "is_complex"):  # (a Python complex would become complex64 through torch.as_tensor)
            rho = np.asarray(rho, dtype=np.complex128)
        t = torch.as_tensor(rho).to(device=self.device, dtype=torch.complex128).reshape(-1)
        if t.numel() == 1:
            t = t.expand(self.nnodes)
        if t.numel() != self.nnodes:
            raise ValueError(f"rho must have {self.nnodes} entries, got {t.numel()}")
        return t.real.contiguous(), t.imag.contiguous()

    def _assemble_complex(self, fn, parts, *args):
        """K for complex rho (complex-step verification): Re K and Im K as two real assemblies, returned as one
        complex128 tensor of CSR values."""
        torch = _torch()
        re, im = self.new_values(), self.new_values()
        with torch.cuda.device(self.device):
            _lib.check(fn(self._handle, _ptr(parts[0]), _ptr(parts[1]), *args, _ptr(re), _ptr(im), self._stream()))
        return torch.complex(re, im)

    def assemble_poisson(self, rho=1.0, p=0.0, out=None, mode="auto"):
        torch = _torch()
        parts = self._complex_parts(rho)
        if parts is not None:
            return self._assemble_complex(self._lib.pfg_assemble_poisson_complex, parts, float(p))
        rho_t, rho_c = self._rho(rho)
        out = self.new_values() if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_assemble_poisson(self._handle, _ptr(rho_t), rho_c, float(p), _ptr(out),
                                                      _lib.MODES[mode], self._stream()))
        return out

    def assemble_elasticity(self, rho=1.0, p=0.0, E=10.0, nu=0.3, out=None, mode="auto"):
        torch = _torch()
        parts = self._complex_parts(rho)
        if parts is not None:
            return self._assemble_complex(self._lib.pfg_assemble_elasticity_complex, parts, float(p), float(E), float(nu))
        rho_t, rho_c = self._rho(rho)
        out = self.new_values() if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_assemble_elasticity(self._handle, _ptr(rho_t), rho_c, float(p), float(E),
                                                         float(nu), _ptr(out), _lib.MODES[mode], self._stream()))
        return out

    def assemble_helmholtz(self, r0, out_K=None, out_R=None, mode="auto"):
        torch = _torch()
        out_K = self.new_values() if out_K is None else out_K
        out_R = self.new_values() if out_R is None else out_R
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_assemble_helmholtz(self._handle, float(r0), _ptr(out_K), _ptr(out_R),
                                                        _lib.MODES[mode], self._stream()))
        return out_K, out_R

    def assemble_nlpoisson(self, xdv, u, want_K=True, want_res=True, out_K=None, out_res=None, mode="auto"):
        torch = _torch()
        xdv = np.ascontiguousarray(np.asarray(xdv, dtype=np.float64).reshape(-1))
        u_t = self._dev_f64(u, self.nnodes, "u")
        if want_K and out_K is None:
            out_K = self.new_values()
        if want_res and out_res is None:
            out_res = self.new_vector()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_assemble_nlpoisson(
                self._handle, xdv.ctypes.data_as(ctypes.POINTER(c_double)), int(xdv.size), _ptr(u_t),
                _ptr(out_K if want_K else None), _ptr(out_res if want_res else None), _lib.MODES[mode],
                self._stream()))
        return (out_K if want_K else None), (out_res if want_res else None)

    def quad_points(self):
        """Physical coordinates of the quadrature points, (nelems, nquads, ndims) on the device."""
        torch = _torch()
        Xq = torch.empty((self.nelems, self.nnodes_per_elem, self.ndims), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_quad_points(self._handle, _ptr(Xq), self._stream()))
        return Xq

    def poisson_rhs(self, gq, out=None, mode="auto"):
        torch = _torch()
        gq = self._dev_f64(gq, self.nelems * self.nnodes_per_elem, "g at the quadrature points")
        out = self.new_vector() if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_poisson_rhs(self._handle, _ptr(gq), _ptr(out), _lib.MODES[mode], self._stream()))
        return out

    def apply_dirichlet(self, vals, rhs, dof_fixed, dof_fixed_vals=None, enforce_symmetric=True):
        """Device-side ModelBase.apply_dirichlet_bcs with the pattern kept (explicit zeros stay)."""
        torch = _torch()
        fixed = torch.as_tensor(np.asarray(dof_fixed, dtype=np.int64)).to(self.device)
        fv = None
        if dof_fixed_vals is not None:
            fv = self._dev_f64(dof_fixed_vals, fixed.numel(), "dof_fixed_vals")
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_apply_dirichlet(self._handle, _ptr(fixed), _ptr(fv), int(fixed.numel()),
                                                     1 if enforce_symmetric else 0, _ptr(vals), _ptr(rhs),
                                                     self._stream()))
        return vals, rhs

    def spmv(self, vals, x, out=None):
        torch = _torch()
        x = self._dev_f64(x, self.ncols, "x")
        out = self.new_vector() if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_spmv(self._handle, _ptr(vals), _ptr(x), _ptr(out), self._stream()))
        return out

    def spmv_t(self, vals, x, out=None):
        """y = A^T x without forming the transpose (Helmholtz.apply_gradient's RT.dot, pyfem.py:2114)."""
        torch = _torch()
        x = self._dev_f64(x, self.nrows, "x")
        out = torch.empty(self.ncols, dtype=torch.float64, device=self.device) if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_spmv_t(self._handle, _ptr(vals), _ptr(x), _ptr(out), self._stream()))
        return out

    def cg(self, vals, b, x0=None, rtol=1e-8, atol=0.0, max_iter=None, check_every=16):
        """Jacobi-preconditioned conjugate gradients on the device CSR (the matrix never leaves HBM): the device
        stand-in for Assembler._solve_linear_system(method="cg") (pyfem.py:2403-2423).  Returns (x, iterations,
        |r|); raises RuntimeError like the reference when max_iter is reached first."""
        torch = _torch()
        b = self._dev_f64(b, self.nrows, "b")
        x = torch.empty_like(b) if x0 is None else self._dev_f64(x0, self.nrows, "x0").clone()
        iters, resid = ctypes.c_int(0), c_double(0.0)
        max_iter = 10 * self.nrows if max_iter is None else int(max_iter)  # scipy's default maxiter
        with torch.cuda.device(self.device):
            st = self._lib.pfg_cg(self._handle, _ptr(vals), _ptr(b), _ptr(x), 1 if x0 is None else 0, float(rtol),
                                  float(atol), max_iter, int(check_every), byref(iters), byref(resid), self._stream())
        if st == _lib.PFG_ERR_NOCONV:
            raise RuntimeError(f"cg failed with code {iters.value}")  # scipy reports the iteration count as the code
        _lib.check(st)
        return x, int(iters.value), float(resid.value)

    def bicgstab(self, vals, b, rtol=1e-8, atol=0.0, max_iter=None, check_every=8):
        """Jacobi-preconditioned BiCGStab on the device CSR for non-symmetric systems (the Newton Jacobian of
        NonlinearPoisson2D): the device stand-in for Assembler._solve_linear_system(method="gmres").  Returns
        (x, iterations, |r|)."""
        torch = _torch()
        b = self._dev_f64(b, self.nrows, "b")
        x = torch.empty_like(b)
        iters, resid = ctypes.c_int(0), c_double(0.0)
        max_iter = 10 * self.nrows if max_iter is None else int(max_iter)
        with torch.cuda.device(self.device):
            st = self._lib.pfg_bicgstab(self._handle, _ptr(vals), _ptr(b), _ptr(x), float(rtol), float(atol), max_iter,
                                        int(check_every), byref(iters), byref(resid), self._stream())
        if st == _lib.PFG_ERR_NOCONV:
            raise RuntimeError(f"bicgstab failed with code {iters.value}")
        _lib.check(st)
        return x, int(iters.value), float(resid.value)

    # ---- the scatter on its own, and element matrices without the scatter ------------------------------
    @property
    def ndof_per_elem(self):
        return self.nnodes_per_elem * self.ndof_per_node

    def scatter_matrix(self, Ke, out=None, mode="auto"):
        """ModelBase._assemble_jacobian(Ke_mat) (pyfem.py:920-931) for caller-supplied element matrices
        (nelems, D, D): CSR values on the device."""
        torch = _torch()
        D = self.ndof_per_elem
        Ke = self._dev_f64(Ke, self.nelems * D * D, "element matrices")
        out = self.new_values() if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_scatter_matrix(self._handle, _ptr(Ke), _ptr(out), _lib.MODES[mode], self._stream()))
        return out

    def scatter_vector(self, fe, out=None, mode="auto"):
        """ModelBase._assemble_rhs(rhs_e, rhs) (pyfem.py:860-875) for scalar handles: fe is (nelems, nnodes_per_elem)."""
        torch = _torch()
        fe = self._dev_f64(fe, self.nelems * self.nnodes_per_elem, "element vectors")
        out = self.new_vector() if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_scatter_vector(self._handle, _ptr(fe), _ptr(out), _lib.MODES[mode], self._stream()))
        return out

    def element_matrices(self, physics, field=None, field_const=1.0, params=(), want_Ke=True, want_Ke2=False,
                         want_fe=False):
        """Element matrices / vectors without the scatter (the reference's Ke_mat, Re, rhs_e): device tensors
        (nelems, D, D) and (nelems, nnodes_per_elem).  physics: "poisson" params (p,), "elasticity" params (p, E, nu),
        "helmholtz" params (r0,) [Ke2 = Re], "nlpoisson" params = xdv, field = u [fe = residual]."""
        torch = _torch()
        code = {"poisson": _lib.PHYS_POISSON, "elasticity": _lib.PHYS_ELASTICITY, "helmholtz": _lib.PHYS_HELMHOLTZ,
                "nlpoisson": _lib.PHYS_NLPOISSON}[physics]
        D = self.ndof_per_elem
        f = None if field is None else self._dev_f64(field, self.nnodes, "nodal field")
        par = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
        mk = lambda shape: torch.empty(shape, dtype=torch.float64, device=self.device)
        Ke = mk((self.nelems, D, D)) if want_Ke else None
        Ke2 = mk((self.nelems, D, D)) if want_Ke2 else None
        fe = mk((self.nelems, self.nnodes_per_elem)) if want_fe else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_element_matrices(
                self._handle, code, _ptr(f), float(field_const), par.ctypes.data_as(ctypes.POINTER(c_double)),
                int(par.size), _ptr(Ke), _ptr(Ke2), _ptr(fe), self._stream()))
        return Ke, Ke2, fe

    def k_dv_sens(self, physics, rho, p, phi, psi, E=10.0, nu=0.3, out=None, deterministic=False):
        """d(phi^T K(rho) psi)/d rho at the owned nodes (the reference's _compute_K_dv_sens, pyfem.py:1239-1276 and
        1872-1920), fused on the device.  physics: "poisson" (scalar handle) or "elasticity" (a handle with ndims dofs
        per node, or a scalar handle of the same mesh).  Default: element-per-thread pass with atomic nodal adds.
        deterministic=True (scalar handles with a gather plan): node-window staging of rho / phi / psi and
        plan-ordered nodal sums -- bitwise reproducible, measured ~1.5x slower (pfg_k_dv_sens_ordered)."""
        torch = _torch()
        code = {"poisson": _lib.PHYS_POISSON, "elasticity": _lib.PHYS_ELASTICITY}[physics]
        rho_t, rho_c = self._rho(rho)
        # phi / psi carry one entry per node (Poisson) or ndims entries (elasticity), whatever this handle's own dof
        # count: a scalar handle of the mesh runs the deterministic tile-plan pass for both physics
        ndof = self.nnodes * (1 if physics == "poisson" else self.ndims)
        phi_t, psi_t = self._dev_f64(phi, ndof, "phi"), self._dev_f64(psi, ndof, "psi")
        par = np.ascontiguousarray(np.asarray((E, nu), dtype=np.float64))
        if out is None:
            out = torch.empty(self.own_end - self.own_begin, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            fn = self._lib.pfg_k_dv_sens_ordered if deterministic else self._lib.pfg_k_dv_sens
            _lib.check(fn(self._handle, code, _ptr(rho_t), rho_c, float(p),
                          par.ctypes.data_as(ctypes.POINTER(c_double)), 2, _ptr(phi_t), _ptr(psi_t), _ptr(out),
                          self._stream()))
        return out

    # ---- multi-GPU reduce variant (halo.py) ---------------------------------------------------------
    def set_element_mask(self, skip):
        """Elements with skip != 0 stay in the pattern but are not integrated by this handle (another rank ships
        their contributions, halo.ReduceAssembler).  None clears the mask."""
        torch = _torch()
        t = None
        if skip is not None:
            t = torch.as_tensor(np.asarray(skip, dtype=np.uint8)).to(self.device).contiguous()
            if t.numel() != self.nelems:
                raise ValueError(f"element mask must have {self.nelems} entries")
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_mesh_set_element_mask(self._handle, _ptr(t), self._stream()))

    def add_indexed(self, vals, idx, src):
        """vals[idx] += src on the device (idx unique): a neighbour's interface-row contributions."""
        torch = _torch()
        n = int(idx.numel())
        if src.numel() != n:
            raise ValueError("idx and src differ in length")
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pfg_add_indexed(_ptr(vals), _ptr(idx), _ptr(src), n, self._stream()))
        return vals

    # ---- host views -------------------------------------------------------------------------------
    def _pinned_values(self):
        """A pinned host buffer of nnz doubles that no live matrix references: buffers handed out earlier are reused
        as soon as the scipy matrix built on them is gone (reference counts), so a Newton loop alternates between
        two buffers and never pays a fresh 4.8 GB allocation + page-fault pass per assembly."""
        import sys
        torch = _torch()
        for arr in self._host_pool:
            if sys.getrefcount(arr) <= 3:  # the pool's reference, the loop variable, getrefcount's argument
                return arr
        arr = torch.empty(self.nnz, dtype=torch.float64).pin_memory().numpy()
        self._host_pool.append(arr)
        if len(self._host_pool) > 4:  # matrices that stay alive keep their buffers; the pool forgets the oldest
            self._host_pool.pop(0)
        return arr

    def to_scipy(self, vals, copy_pattern=True, out=None, reuse_host_buffers=False):
        """scipy.sparse.csr_matrix on the host from device values (one D2H copy of nnz doubles).

        copy_pattern=True hands out private copies of indptr / indices, so callers may edit the matrix in
        place without touching the cached pattern.  copy_pattern=False shares the cached arrays, marked read-only:
        an in-place pattern edit (eliminate_zeros, sort_indices) then raises instead of corrupting the cache, and
        ModelBase.apply_dirichlet_bcs copies them first (copy-on-write).
        `out` may be a float64 host array (e.g. the numpy view of a pinned torch tensor) to receive the values;
        reuse_host_buffers=True takes it from the handle's pool of pinned buffers instead.
        """
        torch = _torch()
        from scipy import sparse
        indptr, indices = self.pattern_host()
        if vals.is_complex():  # complex-step verification: a fresh complex128 array
            out = np.empty(self.nnz, dtype=np.complex128)
        elif out is None:
            out = self._pinned_values() if reuse_host_buffers else np.empty(self.nnz, dtype=np.float64)
        data = out
        torch.from_numpy(data).copy_(vals)
        if copy_pattern:
            indptr, indices = indptr.copy(), indices.copy()
        K = sparse.csr_matrix((data, indices, indptr), shape=(self.nrows, self.ncols), copy=False)
        K.has_sorted_indices = True
        K.has_canonical_format = True
        K._pfg_shared_pattern = not copy_pattern
        return K
