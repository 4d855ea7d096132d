/*
 * pyfem_b200.h -- C ABI of the B200-native finite-element assembly engine.
 *
 * The reference (aaronyicongfu/pyfem_gpu_testflight) is pure Python and has no FFI; its
 * boundary for this path is the physics-model API in pyfem.py (ModelBase and subclasses).
 * Each entry point below names the reference interface it replaces (file:line relative to
 * the reference root).  The Python host mirror (pyfem_gpu_testflight_b200/) binds these
 * with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every *_dev pointer is a CUDA device pointer owned by the caller
 *     (torch tensor.data_ptr()); `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - calls enqueue work on `stream` and return without synchronising, except
 *     pfg_mesh_create (synchronises: it sizes allocations from device results).
 *   - return value: PFG_OK or a negative pfg_status; pfg_last_error() gives the message for
 *     the calling thread's most recent failure.
 *   - a handle is used by one host thread at a time (the reference's models are not
 *     re-entrant either: they mutate shared scratch, pyfem.py:705-756).
 *   - float64 throughout; node / element ids are int64 at the boundary as in the reference
 *     (pyfem.py:660, utils.py:290-292).
 */
#ifndef PYFEM_B200_H
#define PYFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFG_ABI_VERSION 1

#if defined(__GNUC__)
#define PFG_API __attribute__((visibility("default")))
#else
#define PFG_API
#endif

typedef struct pfg_mesh pfg_mesh; /* opaque: device copies of the mesh, CSR pattern, scatter/gather plans */

typedef enum pfg_status {
    PFG_OK = 0,
    PFG_ERR_INVALID = -1,     /* bad argument (reference: ValueError) */
    PFG_ERR_CUDA = -2,        /* CUDA runtime failure (reference: RuntimeError) */
    PFG_ERR_UNSUPPORTED = -3, /* element / physics combination not implemented on the device */
    PFG_ERR_MESH = -4,        /* conn.min() != 0 or conn.max() != nnodes-1 (reference asserts, pyfem.py:680-681) */
    PFG_ERR_NOMEM = -5,
    PFG_ERR_NOCONV = -6       /* pfg_cg: max_iter reached before the tolerance (reference: RuntimeError "cg failed") */
} pfg_status;

typedef enum pfg_elem {
    PFG_QUAD4 = 4, /* BasisBilinear2D + QuadratureBilinear2D (pyfem.py:83-95, 253-284) */
    PFG_HEX8 = 8   /* BasisBlock3D + QuadratureBlock3D (pyfem.py:97-112, 287-338) */
} pfg_elem;

typedef enum pfg_mode {
    PFG_MODE_AUTO = 0,   /* the faster strategy for the operator: owner-computes (gather) when the mesh has a plan
                            for it, else atomic */
    PFG_MODE_ATOMIC = 1, /* element-per-thread, slot-indexed red.global.add.f64 scatter */
    PFG_MODE_GATHER = 2  /* owner-computes: element matrices staged in shared memory, every CSR value
                            summed in a fixed order and written exactly once (no atomics, deterministic) */
} pfg_mode;

typedef enum pfg_info {
    PFG_INFO_NNZ = 0,        /* CSR nnz of the owned rows */
    PFG_INFO_NROWS = 1,      /* owned dof rows = ndof_per_node * (own_end - own_begin) */
    PFG_INFO_NCOLS = 2,      /* global dof columns */
    PFG_INFO_IDX_BYTES = 3,  /* 4 or 8: the index width scipy's coo->csr would pick (scipy _coo.py:59-61,419) */
    PFG_INFO_NCHUNKS = 4,    /* row chunks of the gather plan (0 = no plan) */
    PFG_INFO_CHUNK_ELEMS = 5,/* sum over chunks of elements touching the chunk (>= nelems: halo recompute) */
    PFG_INFO_PLAN_BYTES = 6, /* device bytes of plan metadata read per assembly (chunk directory + distinct templates) */
    PFG_INFO_DEVICE_BYTES = 7,/* device bytes held by the handle */
    PFG_INFO_MAX_ROW_BLOCKS = 8, /* largest number of neighbour nodes of any owned node */
    PFG_INFO_MAX_VALENCE = 9,    /* largest number of elements around any node */
    PFG_INFO_HEX_ROWS = 10,      /* 1: hex8 elasticity runs the owner-computes geometry + chunk-row passes in AUTO / GATHER mode */
    PFG_INFO_TEMPLATES = 11      /* distinct chunk templates of the tile plan (== NCHUNKS when no two chunks share tables) */
} pfg_info;

PFG_API int pfg_abi_version(void);
PFG_API const char* pfg_last_error(void);

/*
 * Once per mesh.  Replaces the device-relevant part of ModelBase.__init__ (pyfem.py:640-757):
 * casts + sanity asserts (:659-681), utils.create_dof (utils.py:267-298), the COO pattern of
 * ModelBase._compute_nz_pattern (pyfem.py:837-858) -- here turned directly into the CSR pattern
 * that coo_matrix(...).tocsr() (pyfem.py:930-931) would produce, plus the element->slot map
 * and the row-chunk gather plans used by every later assembly.
 *
 *   X_dev     (nnodes, ndims) float64 row-major, ndims = 2 for QUAD4, 3 for HEX8
 *   conn_dev  (nelems, nnodes_per_elem) int64 row-major, values in [0, nnodes)
 *   own_begin/own_end   node rows [own_begin, own_end) this handle assembles (0, nnodes for
 *             a single GPU).  Multi-GPU: each rank passes its element block plus one layer of
 *             ghost elements and owns a contiguous slab of node rows.
 *   node_gid_dev  NULL, or (nnodes,) int64 strictly increasing local->global node ids used for
 *             the column indices the pattern reports; ncols_global_nodes = global node count
 *             (ignored when node_gid_dev is NULL).
 *   flags     0, or PFG_CREATE_NO_GATHER_PLAN to skip building the gather plan.
 */
#define PFG_CREATE_NO_GATHER_PLAN 1
#define PFG_CREATE_NO_REORDER 2 /* chunk nodes in id order instead of the coordinate tiling */
PFG_API int pfg_mesh_create(pfg_mesh** out, int elem_type, int ndof_per_node, int64_t nnodes, int64_t nelems,
                    const double* X_dev, const int64_t* conn_dev, int64_t own_begin, int64_t own_end,
                    const int64_t* node_gid_dev, int64_t ncols_global_nodes, int flags, void* stream);
PFG_API int pfg_mesh_destroy(pfg_mesh* mesh);
PFG_API int pfg_mesh_get(const pfg_mesh* mesh, int what, int64_t* value);

/*
 * Multi-GPU, interface rows summed by a reduce (the north_star's "NCCL reduce over NVLink" variant; the
 * default ghost-layer variant needs no exchange).  elem_skip_dev: (nelems,) uint8, non-zero for elements of
 * this handle's mesh whose contributions arrive from another rank: they stay in the pattern but are not
 * integrated.  NULL clears the mask.  Reference: none (the reference is single-process); the summation it
 * completes is the duplicate-summing of coo->csr (pyfem.py:930-931) across partitions.
 */
PFG_API int pfg_mesh_set_element_mask(pfg_mesh* mesh, const uint8_t* elem_skip_dev, void* stream);

/*
 * vals[idx[i]] += src[i] for i < n, idx unique: adds a neighbour's interface-row contributions (received with
 * NCCL into src_dev) into the owner's CSR values or rhs.  Part of the same cross-partition duplicate sum.
 */
PFG_API int pfg_add_indexed(double* vals_dev, const int64_t* idx_dev, const double* src_dev, int64_t n, void* stream);

/*
 * CSR pattern of the owned rows, identical to K.indptr / K.indices of the reference's
 * ModelBase._assemble_jacobian (pyfem.py:920-931): sorted unique columns, explicit zeros kept.
 *   indptr_dev  (nrows + 1) entries, indices_dev (nnz) entries, each idx_bytes (4 or 8) wide.
 */
PFG_API int pfg_mesh_pattern(const pfg_mesh* mesh, void* indptr_dev, void* indices_dev, int idx_bytes, void* stream);

/*
 * LinearPoisson.compute_jacobian(rho) (pyfem.py:1005-1030): RAMP material update (:1278-1301),
 * element matrices (:1175-1217) and CSR scatter (:920-931) in one pass.
 *   rho_dev  (nnodes,) nodal density or NULL for the constant rho_const (pyfem.py:1015-1016)
 *   vals_dev (nnz,) CSR values, overwritten.
 */
PFG_API int pfg_assemble_poisson(pfg_mesh* mesh, const double* rho_dev, double rho_const, double p, double* vals_dev,
                         int mode, void* stream);

/*
 * LinearElasticity.compute_jacobian(rho) (pyfem.py:1770-1795): plane stress for QUAD4
 * (ndof_per_node 2), 3-D for HEX8 (ndof_per_node 3); C0 from E, nu as pyfem.py:1746-1757.
 * HEX8: PFG_MODE_AUTO / PFG_MODE_GATHER run the owner-computes geometry + chunk-row passes when the mesh
 * allows (PFG_INFO_HEX_ROWS == 1; the first call allocates nelems * 640 B of scratch inside the handle),
 * otherwise AUTO is the atomic scatter and GATHER the staged-row-block kernel.
 */
PFG_API int pfg_assemble_elasticity(pfg_mesh* mesh, const double* rho_dev, double rho_const, double p, double E,
                            double nu, double* vals_dev, int mode, void* stream);

/*
 * Helmholtz.__init__ assembly (pyfem.py:2084-2097, element kernels :2126-2177):
 * K = r0^2 stiffness + mass and R = mass on one pattern.  Either output may be NULL.
 */
PFG_API int pfg_assemble_helmholtz(pfg_mesh* mesh, double r0, double* K_vals_dev, double* R_vals_dev, int mode,
                           void* stream);

/*
 * NonlinearPoisson2D.compute_jacobian(xdv, u) (pyfem.py:1390-1404, :1541-1610) and
 * compute_rhs(xdv, u) = residual (pyfem.py:1375-1388, :1474-1539) in one pass over the
 * elements.  xdv_host is a HOST array of nxdv (<= 32) design variables.  Either output may be
 * NULL.  res_dev has one entry per owned dof row.
 */
PFG_API int pfg_assemble_nlpoisson(pfg_mesh* mesh, const double* xdv_host, int nxdv, const double* u_dev,
                           double* K_vals_dev, double* res_dev, int mode, void* stream);

/*
 * LinearPoisson.compute_rhs (pyfem.py:996-1003, :1125-1173, scatter :860-875), split in two
 * because the source term gfunc is a user Python callable (pyfem.py:957,1127):
 *   pfg_quad_points  writes Xq (nelems, nquads, ndims), the physical quadrature coordinates
 *                    (utils.compute_elem_interp, utils.py:203-221) the callable is evaluated on;
 *   pfg_poisson_rhs  takes g at those points, (nelems, nquads), and writes
 *                    rhs[i] = sum_e sum_q detJ w N g for the owned rows.
 */
/*
 * K(rho) for a COMPLEX nodal density: the reference's complex-step checks (pyfem.py:1018-1020, 1289-1292, 1783-1785,
 * 1933-1936; tests/test_linear_poisson.py:57-89, tests/test_elasticity.py:68-104) call compute_jacobian(rho + 1j h p)
 * and read dK/drho . p off the imaginary part.  The element matrices are real multiples of the complex RAMP factor
 * c(rho_q) = rho_q / (1 + p (1 - rho_q)), so Re K and Im K are two real assemblies, one per part of the factor.
 *   rho_re_dev, rho_im_dev  (nnodes,) both parts of the nodal density
 *   vals_re_dev, vals_im_dev (nnz,) CSR values of Re K and Im K (either may be NULL)
 * Slot-indexed atomic scatter (this is a verification path, not a hot one); any mesh the handle was built for.
 */
PFG_API int pfg_assemble_poisson_complex(pfg_mesh* mesh, const double* rho_re_dev, const double* rho_im_dev, double p,
                                         double* vals_re_dev, double* vals_im_dev, void* stream);
PFG_API int pfg_assemble_elasticity_complex(pfg_mesh* mesh, const double* rho_re_dev, const double* rho_im_dev, double p,
                                            double E, double nu, double* vals_re_dev, double* vals_im_dev, void* stream);

PFG_API int pfg_quad_points(pfg_mesh* mesh, double* Xq_dev, void* stream);
PFG_API int pfg_poisson_rhs(pfg_mesh* mesh, const double* gq_dev, double* rhs_dev, int mode, void* stream);

/*
 * The scatter on its own, for caller-supplied element matrices / vectors -- the slot the reference's A2DWrapper uses
 * (pyfem.py:2255-2277: a native plugin fills the element Jacobians, pyfem assembles them):
 *   pfg_scatter_matrix  ModelBase._assemble_jacobian(Ke_mat) (pyfem.py:920-931); Ke_dev is (nelems, D, D) row-major
 *                       with D = nnodes_per_elem * ndof_per_node in the interleaved (node, axis) dof order of
 *                       utils.create_dof (utils.py:293-296); duplicates summed, explicit zeros kept.
 *   pfg_scatter_vector  ModelBase._assemble_rhs(rhs_e, rhs) (pyfem.py:860-875) for scalar handles; fe_dev is
 *                       (nelems, nnodes_per_elem).
 */
PFG_API int pfg_scatter_matrix(pfg_mesh* mesh, const double* Ke_dev, double* vals_dev, int mode, void* stream);
PFG_API int pfg_scatter_vector(pfg_mesh* mesh, const double* fe_dev, double* rhs_dev, int mode, void* stream);

/*
 * Element matrices / vectors without the scatter: the outputs of the reference's _compute_element_jacobian and
 * _compute_element_rhs methods (Ke_mat (nelems, D, D), rhs_e (nelems, nnodes_per_elem)) for callers that read them
 * (examples/SciTech2023/performance/performance_test.py:52).
 *   physics      PFG_PHYS_POISSON     pyfem.py:1188-1217   params {p}              field = rho
 *                PFG_PHYS_ELASTICITY  pyfem.py:2029-2068   params {p, E, nu}       field = rho
 *                PFG_PHYS_HELMHOLTZ   pyfem.py:2138-2177   params {r0}             Ke_dev = Ke, Ke2_dev = Re
 *                PFG_PHYS_NLPOISSON   pyfem.py:1541-1610, 1474-1539   params = xdv (nparams of them), field = u;
 *                                     Ke_dev = Jacobian, fe_dev = residual
 *   field_dev    nodal field or NULL for the constant field_const; params_host is a HOST array.
 */
typedef enum pfg_physics { PFG_PHYS_POISSON = 1, PFG_PHYS_ELASTICITY = 2, PFG_PHYS_HELMHOLTZ = 3, PFG_PHYS_NLPOISSON = 4 } pfg_physics;
PFG_API int pfg_element_matrices(pfg_mesh* mesh, int physics, const double* field_dev, double field_const,
                                 const double* params_host, int nparams, double* Ke_dev, double* Ke2_dev,
                                 double* fe_dev, void* stream);

/*
 * ModelBase.apply_dirichlet_bcs (pyfem.py:780-835) on the device CSR, keeping the pattern
 * (no eliminate_zeros): rows of fixed dofs zeroed, columns too when enforce_symmetric != 0,
 * unit diagonal, rhs[fixed] = vals (0 when fixed_vals_dev is NULL) and, in the symmetric
 * case with values, rhs[free] -= K_free,fixed . vals (uses the values before zeroing).
 *   fixed_dofs_dev (nfixed,) int64 global dof ids; rhs_dev may be NULL.
 */
PFG_API int pfg_apply_dirichlet(pfg_mesh* mesh, const int64_t* fixed_dofs_dev, const double* fixed_vals_dev,
                        int64_t nfixed, int enforce_symmetric, double* vals_dev, double* rhs_dev, void* stream);

/*
 * _compute_K_dv_sens(rho, phi, psi): d(phi^T K(rho) psi) / d rho at the nodes, the kernel behind compliance_grad
 * (LinearPoisson pyfem.py:1239-1276 with :1220-1236 and :1304-1329; LinearElasticity pyfem.py:1872-1920).  Fused:
 * no (nelems, D, D, nnodes_per_elem) derivative tensor is formed -- per quadrature point the scalar
 * ramp'(rho_q) detJ w (grad phi . grad psi   or   strain(phi)^T C0 strain(psi)) is spread with N[q, o] and added to
 * the element's nodes (np.add.at, pyfem.py:1272-1275).
 *   physics      PFG_PHYS_POISSON (handle with 1 dof per node, params {})  or
 *                PFG_PHYS_ELASTICITY (handle with ndims dofs per node, params_host = {E, nu})
 *   rho_dev      nodal density or NULL for the constant rho_const; p is the RAMP parameter
 *   phi_dev, psi_dev   dof vectors (nnodes * ndof_per_node,), local node numbering
 *   out_dev      one entry per owned node; zeroed here, then accumulated with atomic adds
 * For PFG_PHYS_ELASTICITY a scalar handle (1 dof per node) of the same mesh is accepted as well: phi / psi carry
 * ndims entries per node whatever the handle's own dof count.
 */
PFG_API int pfg_k_dv_sens(pfg_mesh* mesh, int physics, const double* rho_dev, double rho_const, double p,
                          const double* params_host, int nparams, const double* phi_dev, const double* psi_dev,
                          double* out_dev, void* stream);

/*
 * The same sensitivities without atomics, bitwise reproducible: rho / phi / psi of a chunk's window nodes are staged
 * once in shared memory and every node sums its elements' shares in plan order (the tile kernel's vector path).  Needs
 * a SCALAR handle (1 dof per node) built with a gather plan, for either physics.  Same arguments as pfg_k_dv_sens.
 * Measured slower than the atomic pass (the per-point arithmetic, not the scatter, is what bounds this operator, and
 * the owner-computes plan integrates chunk-border elements twice): use it where run-to-run reproducibility matters.
 */
PFG_API int pfg_k_dv_sens_ordered(pfg_mesh* mesh, int physics, const double* rho_dev, double rho_const, double p,
                                  const double* params_host, int nparams, const double* phi_dev,
                                  const double* psi_dev, double* out_dev, void* stream);

/*
 * y = A x on the device CSR of the owned rows (Helmholtz.compute_rhs = R.dot(x), pyfem.py:2117-2120).
 * x is indexed by global column (ncols entries), y has one entry per owned dof row.
 */
PFG_API int pfg_spmv(pfg_mesh* mesh, const double* vals_dev, const double* x_dev, double* y_dev, void* stream);

/*
 * y = A^T x (Helmholtz.apply_gradient = RT.dot(Ksolve.solve(g)), pyfem.py:2109-2115), without forming the
 * transpose: entry (c, r) is read from row c at the rank of r among c's columns.  Needs a handle that owns every
 * row (PFG_ERR_UNSUPPORTED for a rank's slab).
 */
PFG_API int pfg_spmv_t(pfg_mesh* mesh, const double* vals_dev, const double* x_dev, double* y_dev, void* stream);

/*
 * Jacobi-preconditioned conjugate gradients on the device CSR: the GPU stand-in for
 * Assembler._solve_linear_system(K, rhs, method="cg") (pyfem.py:2403-2423) and for the solves inside compliance
 * (pyfem.py:1050-1068, 1814-1828), so that a solve never copies K to the host.  The reference preconditions with
 * pyamg smoothed aggregation; the stopping rule is scipy's: |b - A x| <= max(rtol |b|, atol).
 *   vals_dev    CSR values (after pfg_apply_dirichlet); symmetric positive definite
 *   b_dev       right-hand side (nrows)
 *   x_dev       in: initial guess unless x_is_zero != 0 (then zero-filled here); out: the solution
 *   check_every iterations between two convergence checks (each check is one device->host copy of the partial sums
 *               of |r|^2; <= 0: 16).  Dot products are summed in a fixed order: results are reproducible.
 *   iters_out, resid_out   (host, may be NULL) iterations done and the final |r|
 * Returns PFG_OK, or PFG_ERR_NOCONV when max_iter was reached first (x_dev holds the last iterate).  Needs a handle
 * that owns every row.
 */
PFG_API int pfg_cg(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev, int x_is_zero,
                   double rtol, double atol, int max_iter, int check_every, int* iters_out, double* resid_out,
                   void* stream);

/*
 * The same conjugate gradients over the row slabs of several ranks (handles built with own_range / node_gid, one per
 * rank and GPU): what Assembler._solve_linear_system(K, rhs, method="cg") (pyfem.py:2403-2423) becomes when the
 * matrix is distributed by rows.  The library does not link NCCL: the two collective steps are the caller's
 * callbacks, which enqueue their work on `stream` (torch.distributed in the Python binding, slab_solve.py).
 *   x_full_dev  work vector of ncols doubles (global dof numbering): the search direction; entries of the owned rows
 *               are written here, `halo(user)` must make the ghost entries (columns of owned rows owned by other
 *               ranks) current before every product
 *   scal_dev    8 doubles; `reduce(user, offset, count)` must sum scal_dev[offset .. offset + count) over the ranks
 *   row0        global dof index of the rank's first owned row (rows are contiguous)
 *   b_dev, x_dev   the rank's rows of the right-hand side / the solution
 * Both callbacks return 0 on success.  Norms in the stopping rule are global, so every rank runs the same number of
 * iterations and returns the same status.
 */
typedef int (*pfg_reduce_fn)(void* user, int offset, int count);
typedef int (*pfg_halo_fn)(void* user);
PFG_API int pfg_cg_dist(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev, int x_is_zero,
                        double* x_full_dev, double* scal_dev, int64_t row0, double rtol, double atol, int max_iter,
                        int check_every, pfg_reduce_fn reduce, pfg_halo_fn halo, void* user, int* iters_out,
                        double* resid_out, void* stream);

/*
 * The two phases of pfg_cg_dist on their own.  Both only ENQUEUE work on `stream` (no host synchronisation), so a
 * caller can capture pfg_cg_dist_steps -- kernels and the callbacks' collectives -- in a CUDA graph and replay it
 * (slab_solve.py does, two iterations per graph), reading scal_dev[2] = |r|^2 (global) whenever it wants to test
 * convergence.  After pfg_cg_dist_begin: scal_dev[4] = |b|^2, scal_dev[2] = |r0|^2, both global.
 *   first_iter   index of the first of the n_steps iterations (its parity selects the r.z slot: iterations must be
 *                numbered consecutively from 0 across calls)
 */
PFG_API int pfg_cg_dist_begin(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev,
                              int x_is_zero, double* x_full_dev, double* scal_dev, int64_t row0, pfg_reduce_fn reduce,
                              pfg_halo_fn halo, void* user, void* stream);
PFG_API int pfg_cg_dist_steps(pfg_mesh* mesh, const double* vals_dev, double* x_dev, double* x_full_dev,
                              double* scal_dev, int64_t row0, int first_iter, int n_steps, pfg_reduce_fn reduce,
                              pfg_halo_fn halo, void* user, void* stream);

/*
 * Jacobi-preconditioned BiCGStab on the device CSR, for NON-SYMMETRIC systems: the Newton step of
 * Assembler.solve_nonlinear (pyfem.py:2337-2353: K is the Jacobian of NonlinearPoisson2D, the reference solves it with
 * gmres + pyamg or spsolve).  Same arguments and stopping rule as pfg_cg; x_dev is zero-filled here (x0 = 0).  Scalars
 * stay on the device (one-thread kernels between the vector kernels); the host reads |r|^2 every check_every
 * iterations (<= 0: 8).  PFG_ERR_NOCONV when max_iter is reached first.
 */
PFG_API int pfg_bicgstab(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev, double rtol,
                         double atol, int max_iter, int check_every, int* iters_out, double* resid_out, void* stream);

/*
 * pfg_bicgstab over the row slabs of several ranks (the Newton step of Assembler.solve_nonlinear, pyfem.py:2337-2353,
 * on a matrix distributed by rows).  As pfg_cg_dist, except that x_full_dev holds TWO global-length work vectors
 * (2 * ncols doubles: the preconditioned direction and the preconditioned residual) and halo(user, which) refreshes
 * the ghost entries of vector `which` (0 / 1, at x_full_dev + which * ncols) before the product that reads it.
 */
typedef int (*pfg_halo2_fn)(void* user, int which);
PFG_API int pfg_bicgstab_dist(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev,
                              double* x_full_dev, double* scal_dev, int64_t row0, double rtol, double atol, int max_iter,
                              int check_every, pfg_reduce_fn reduce, pfg_halo2_fn halo, void* user, int* iters_out,
                              double* resid_out, void* stream);

/*
 * Diagnostics for the bench's roofline: sustained FP64 FMA throughput of `device` in TFLOP/s (a DFMA-bound kernel,
 * 8 CTAs of 256 threads per SM, 16 independent chains per thread, best of 4 timed launches with CUDA events).  The
 * assembly kernels run their quadrature on the FP64 CUDA cores, so this is the second roof next to HBM bandwidth
 * (SURVEY.md section 8d).  No reference counterpart.
 */
PFG_API int pfg_probe_fp64(int device, int iters, double* tflops_out, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PYFEM_B200_H */
