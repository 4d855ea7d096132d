// GPU consumers of the device CSR: the steps that follow the assembly in every
// caller of the reference, kept in HBM so that a solve never copies the matrix to the host.
//
//   pfg_apply_dirichlet   ModelBase.apply_dirichlet_bcs (pyfem.py:780-835), pattern kept
//   pfg_spmv / pfg_spmv_t R.dot(x) (pyfem.py:2117-2120) and RT.dot(x) (pyfem.py:2109-2115)
//   pfg_cg                Jacobi-preconditioned conjugate gradients on the device CSR: the device stand-in for
//                         Assembler._solve_linear_system(method="cg") (pyfem.py:2403-2423) and the solves inside
//                         compliance (pyfem.py:1050-1068, 1814-1828)
//
// The CSR values are addressed through the node-level pattern of the handle (blk_ptr / nbr): the m dof rows of a
// node are consecutive, each holding k blocks of m values (pfg_internal.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "pfg_internal.cuh"

namespace pfg {

constexpr int kRowLanes = 8;        // lanes that share one dof row in the SpMV kernels
constexpr int kVecThreads = 256;    // CTA size of the vector kernels
constexpr int kMaxPartials = 2048;  // upper bound of CTAs (= partial sums per dot product) of the CG kernels

__device__ __forceinline__ double group_sum(double v) {  // sum over the kRowLanes lanes of a row group
#pragma unroll
    for (int o = kRowLanes / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y = A x: kRowLanes lanes per dof row stride over its k*m values (a warp reads 32 / kRowLanes consecutive rows, i.e.
// one contiguous piece of the values array); x is gathered through the node-level column list.  M (dofs per node) is a
// template parameter so that the index arithmetic folds to shifts / constant divisions, and the row loop is unrolled
// so that the column-index and x gathers of several steps are in flight together.
template <int M, int LANES>
__global__ void __launch_bounds__(256) k_spmv_rows(const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ nbr,
                                                   const int64_t* __restrict__ gid, int64_t nown,
                                                   const double* __restrict__ vals, const double* __restrict__ x,
                                                   double* __restrict__ y, const double* __restrict__ dot_with,
                                                   double* __restrict__ partial) {
    const int sub = threadIdx.x & (LANES - 1);
    const int64_t nrows = nown * M;
    constexpr int RPB = 256 / LANES;  // rows per CTA and step
    double dot = 0.0;                     // this row group's share of <dot_with, y>
    // the trip count is uniform over the CTA (the loop runs on the CTA's first row), so every lane reaches the shuffles
    for (int64_t row0 = blockIdx.x * (int64_t)RPB; row0 < nrows; row0 += (int64_t)gridDim.x * RPB) {
        const int64_t row = row0 + threadIdx.x / LANES;
        double s = 0.0;
        if (row < nrows) {
            const int64_t r = row / M;
            const int alpha = (int)(row - r * M);
            const int64_t p0 = __ldg(blk_ptr + r);
            const int km = (int)(__ldg(blk_ptr + r + 1) - p0) * M;
            const double* __restrict__ v = vals + p0 * (M * M) + (int64_t)alpha * km;
            const int32_t* __restrict__ cols = nbr + p0;
#pragma unroll 4
            for (int j = sub; j < km; j += LANES) {
                const int t = j / M, beta = j - t * M;
                int64_t cnode = __ldg(cols + t);
                if (gid != nullptr) cnode = __ldg(gid + cnode);
                s = fma(__ldcs(v + j), __ldg(x + cnode * M + beta), s);  // the values are streamed: read once
            }
        }
#pragma unroll
        for (int o = LANES / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // sum over the row's lanes
        if (sub == 0 && row < nrows) {
            y[row] = s;
            if (dot_with != nullptr) dot = fma(s, dot_with[row], dot);
        }
    }
    if (partial != nullptr) {  // fused dot product <dot_with, y> (CG: p . A p): one partial sum per CTA, fixed order
        __shared__ double red[RPB];
        if (sub == 0) red[threadIdx.x / LANES] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < RPB; ++i) t += red[i];
            partial[blockIdx.x] = t;
        }
    }
}

static void launch_spmv_rows(unsigned grid, cudaStream_t st, const MeshDev& d, const int64_t* gid, const double* vals,
                             const double* x, double* y, const double* dot_with, double* partial) {
    const int64_t nown = d.own_end - d.own_begin;
    // short rows (quad4: 9 or 18 values) share four lanes, longer ones (hex8: 27 .. 81) eight -- measured on 16.8 M quads:
    // scalar 0.83 -> 0.56 ms, 2 dofs per node 1.99 -> 1.73 ms
    const bool short_rows = d.max_k * d.m <= 24;
    if (d.m == 1 && short_rows) k_spmv_rows<1, 4><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, gid, nown, vals, x, y, dot_with, partial);
    else if (d.m == 1) k_spmv_rows<1, 8><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, gid, nown, vals, x, y, dot_with, partial);
    else if (d.m == 2 && short_rows) k_spmv_rows<2, 4><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, gid, nown, vals, x, y, dot_with, partial);
    else if (d.m == 2) k_spmv_rows<2, 8><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, gid, nown, vals, x, y, dot_with, partial);
    else k_spmv_rows<3, 8><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, gid, nown, vals, x, y, dot_with, partial);
}

// Transposed-slot map, built on the first transposed product and kept in the handle: for node block (r, t) with
// column node c, the rank of r among c's sorted neighbours -- where entry (c, r) of the matrix sits in row c (the
// node-level pattern is symmetric: c is a neighbour of r exactly when r is a neighbour of c).
__global__ void k_transpose_ranks(const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ nbr, int64_t nown,
                                  uint8_t* __restrict__ trank) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nown) return;
    const int64_t p0 = blk_ptr[r], p1 = blk_ptr[r + 1];
    for (int64_t p = p0; p < p1; ++p) {
        const int64_t c = nbr[p];
        const int64_t q0 = blk_ptr[c];
        int lo = 0, hi = (int)(blk_ptr[c + 1] - q0);
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (nbr[q0 + mid] < r) lo = mid + 1; else hi = mid;
        }
        trank[p] = (uint8_t)lo;  // a row has at most kMaxRowBlocks = 255 blocks
    }
}

// y = A^T x for a handle that owns every row: the same lanes-per-row layout as k_spmv_rows; lane j of row (r, alpha)
// takes column dof (c, beta) = j and reads A[(c, beta), (r, alpha)] from row c at the mapped rank.
template <int M, int LANES>
__global__ void __launch_bounds__(256) k_spmv_t_rows(const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ nbr,
                                                     const uint8_t* __restrict__ trank, int64_t nown,
                                                     const double* __restrict__ vals, const double* __restrict__ x,
                                                     double* __restrict__ y) {
    const int sub = threadIdx.x & (LANES - 1);
    const int64_t nrows = nown * M;
    constexpr int RPB = 256 / LANES;
    for (int64_t row0 = blockIdx.x * (int64_t)RPB; row0 < nrows; row0 += (int64_t)gridDim.x * RPB) {
        const int64_t row = row0 + threadIdx.x / LANES;
        double s = 0.0;
        if (row < nrows) {
            const int64_t r = row / M;
            const int alpha = (int)(row - r * M);
            const int64_t p0 = __ldg(blk_ptr + r);
            const int km = (int)(__ldg(blk_ptr + r + 1) - p0) * M;
#pragma unroll 4
            for (int j = sub; j < km; j += LANES) {
                const int t = j / M, beta = j - t * M;
                const int64_t c = __ldg(nbr + p0 + t);
                const int64_t q0 = __ldg(blk_ptr + c);
                const int kc = (int)(__ldg(blk_ptr + c + 1) - q0);
                const int rk = __ldg(trank + p0 + t);
                s = fma(__ldg(vals + q0 * (M * M) + (int64_t)beta * kc * M + rk * M + alpha), __ldg(x + c * M + beta), s);
            }
        }
#pragma unroll
        for (int o = LANES / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (sub == 0 && row < nrows) y[row] = s;
    }
}

// 1 / diag(A) of the owned rows (Jacobi preconditioner); a zero diagonal maps to 1
__global__ void k_inv_diag(const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ nbr, int64_t own_begin,
                           int64_t nown, int m, const double* __restrict__ vals, double* __restrict__ dinv) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= nown * m) return;
    const int64_t r = row / m;
    const int alpha = (int)(row - r * m);
    const int64_t p0 = blk_ptr[r];
    const int k = (int)(blk_ptr[r + 1] - p0);
    const int32_t self = (int32_t)(own_begin + r);
    int lo = 0, hi = k;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (nbr[p0 + mid] < self) lo = mid + 1; else hi = mid;
    }
    const double d = vals[p0 * m * m + (int64_t)alpha * k * m + (int64_t)lo * m + alpha];
    dinv[row] = (d != 0.0) ? 1.0 / d : 1.0;
}

__device__ __forceinline__ double block_sum(double v, double* red) {  // fixed-order sum over the CTA, result in thread 0
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kVecThreads / 32; ++i) t += red[i];
    __syncthreads();
    return t;
}

__device__ __forceinline__ double sum_partials(const double* __restrict__ partial, int n, double* red) {
    double v = 0.0;  // every CTA re-sums the n partial sums of the previous kernel in the same order
    for (int i = threadIdx.x; i < n; i += kVecThreads) v += partial[i];
    __shared__ double bcast;
    const double t = block_sum(v, red);
    if (threadIdx.x == 0) bcast = t;
    __syncthreads();
    return bcast;
}

// out[0] = sum of n partial sums, in a fixed order (one CTA): the value a rank contributes to an all-reduce
__global__ void __launch_bounds__(kVecThreads) k_sum_into(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double red[kVecThreads / 32];
    const double t = sum_partials(partial, n, red);
    if (threadIdx.x == 0) out[0] = t;
}

// r = b - A x0 is formed by the caller as r = b (x0 = 0) or through k_spmv_rows; this kernel starts the recurrences:
// z = dinv r, p = z, partial sums of r.z and r.r
__global__ void __launch_bounds__(kVecThreads) k_cg_init(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                                                         double* __restrict__ z, double* __restrict__ p,
                                                         double* __restrict__ part_rz, double* __restrict__ part_rr) {
    __shared__ double red[kVecThreads / 32];
    double rz = 0.0, rr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
        const double ri = r[i], zi = dinv[i] * ri;
        z[i] = zi, p[i] = zi;
        rz = fma(ri, zi, rz), rr = fma(ri, ri, rr);
    }
    const double a = block_sum(rz, red), b = block_sum(rr, red);
    if (threadIdx.x == 0) part_rz[blockIdx.x] = a, part_rr[blockIdx.x] = b;
}

// alpha = (r.z) / (p.Ap); x += alpha p; r -= alpha Ap; z = dinv r; partial sums of the new r.z and r.r
__global__ void __launch_bounds__(kVecThreads) k_cg_update(int64_t n, int n_spmv_parts, int n_vec_parts,
                                                           const double* __restrict__ part_pAp,
                                                           const double* __restrict__ part_rz_old,
                                                           const double* __restrict__ dinv, const double* __restrict__ p,
                                                           const double* __restrict__ Ap, double* __restrict__ x,
                                                           double* __restrict__ r, double* __restrict__ z,
                                                           double* __restrict__ part_rz_new, double* __restrict__ part_rr) {
    __shared__ double red[kVecThreads / 32];
    const double pAp = sum_partials(part_pAp, n_spmv_parts, red);
    const double rz_old = sum_partials(part_rz_old, n_vec_parts, red);
    const double alpha = (pAp != 0.0) ? rz_old / pAp : 0.0;
    double rz = 0.0, rr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, Ap[i], r[i]);
        const double zi = dinv[i] * ri;
        r[i] = ri, z[i] = zi;
        rz = fma(ri, zi, rz), rr = fma(ri, ri, rr);
    }
    const double a = block_sum(rz, red), b = block_sum(rr, red);
    if (threadIdx.x == 0) part_rz_new[blockIdx.x] = a, part_rr[blockIdx.x] = b;
}

// beta = (r.z)_new / (r.z)_old; p = z + beta p
__global__ void __launch_bounds__(kVecThreads) k_cg_direction(int64_t n, int n_vec_parts, const double* __restrict__ part_rz_new,
                                                              const double* __restrict__ part_rz_old,
                                                              const double* __restrict__ z, double* __restrict__ p) {
    __shared__ double red[kVecThreads / 32];
    const double rz_new = sum_partials(part_rz_new, n_vec_parts, red);
    const double rz_old = sum_partials(part_rz_old, n_vec_parts, red);
    const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads)
        p[i] = fma(beta, p[i], z[i]);
}

__global__ void __launch_bounds__(kVecThreads) k_residual(int64_t n, const double* __restrict__ b, const double* __restrict__ Ax,
                                                          double* __restrict__ r) {
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads)
        r[i] = b[i] - Ax[i];
}

__global__ void __launch_bounds__(kVecThreads) k_norm2_partials(int64_t n, const double* __restrict__ v, double* __restrict__ part) {
    __shared__ double red[kVecThreads / 32];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads)
        s = fma(v[i], v[i], s);
    const double a = block_sum(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}

// ---- Dirichlet rows / columns on the device CSR, pattern kept (pyfem.py:780-835 minus eliminate_zeros) ----------
__global__ void k_mark_fixed(const int64_t* __restrict__ fixed, const double* __restrict__ fixed_vals, int64_t nfixed,
                             int64_t ncols, uint8_t* __restrict__ is_fixed, double* __restrict__ u0) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nfixed) return;
    const int64_t dof = fixed[i];
    if (dof < 0 || dof >= ncols) return;
    is_fixed[dof] = 1;
    u0[dof] = fixed_vals ? fixed_vals[i] : 0.0;  // u0 is read at fixed dofs only
}

__global__ void k_apply_dirichlet(const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ nbr,
                                  const int64_t* __restrict__ gid, int64_t own_begin, int64_t nown, int m,
                                  const uint8_t* __restrict__ is_fixed, const double* __restrict__ u0, int symmetric,
                                  int have_vals, double* __restrict__ vals, double* __restrict__ rhs) {
    // one thread per owned dof row
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= nown * m) return;
    const int64_t r = row / m;
    const int alpha = (int)(row - r * m);
    const int64_t node = own_begin + r;
    const int64_t grow = (gid ? gid[node] : node) * m + alpha;
    const int64_t p0 = blk_ptr[r], k = blk_ptr[r + 1] - p0;
    double* v = vals + p0 * m * m + alpha * k * m;
    const bool row_fixed = is_fixed[grow];
    double corr = 0.0;
    for (int64_t t = 0; t < k; ++t) {
        const int64_t cnode = nbr[p0 + t];
        const int64_t gcol0 = (gid ? gid[cnode] : cnode) * m;
        for (int beta = 0; beta < m; ++beta) {
            const int64_t gcol = gcol0 + beta;
            double& x = v[t * m + beta];
            if (row_fixed) {
                x = (gcol == grow) ? 1.0 : 0.0;
            } else if (symmetric && is_fixed[gcol]) {
                if (have_vals) corr = fma(x, u0[gcol], corr);
                x = 0.0;
            }
        }
    }
    if (rhs) {
        if (row_fixed) rhs[row] = u0[grow];
        else if (symmetric && have_vals) rhs[row] -= corr;
    }
}

// ---- BiCGStab (non-symmetric systems: the Newton Jacobian of NonlinearPoisson2D, pyfem.py:1595-1609) ---------------
// Right-preconditioned with Jacobi.  Scalars live in a small device array `sc`: [0] rho, [1] alpha, [2] omega, [3] beta;
// one-thread kernels turn the fixed-order partial sums into the next scalar, so an iteration never syncs with the host.
__device__ __forceinline__ double serial_sum(const double* __restrict__ part, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += part[i];
    return s;
}
__global__ void k_bicg_scalar(int which, int n_a, const double* __restrict__ part_a, int n_b,
                              const double* __restrict__ part_b, double* __restrict__ sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (which == 0) {  // start: rho = rhat . r, alpha = omega = 1, beta = 0
        sc[0] = serial_sum(part_a, n_a), sc[1] = 1.0, sc[2] = 1.0, sc[3] = 0.0;
    } else if (which == 1) {  // alpha = rho / (rhat . v)
        const double d = serial_sum(part_a, n_a);
        sc[1] = (d != 0.0) ? sc[0] / d : 0.0;
    } else if (which == 2) {  // omega = (t . s) / (t . t)
        const double tt = serial_sum(part_b, n_b);
        sc[2] = (tt != 0.0) ? serial_sum(part_a, n_a) / tt : 0.0;
    } else {  // beta = (rho_new / rho) (alpha / omega); rho = rho_new
        const double rho_new = serial_sum(part_a, n_a);
        sc[3] = (sc[0] != 0.0 && sc[2] != 0.0) ? (rho_new / sc[0]) * (sc[1] / sc[2]) : 0.0;
        sc[0] = rho_new;
    }
}

// p = r + beta (p - omega v); y = dinv p
__global__ void __launch_bounds__(kVecThreads) k_bicg_p(int64_t n, const double* __restrict__ sc, const double* __restrict__ dinv,
                                                        const double* __restrict__ r, const double* __restrict__ v,
                                                        double* __restrict__ p, double* __restrict__ y) {
    const double beta = sc[3], omega = sc[2];
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
        const double pi = fma(beta, fma(-omega, v[i], p[i]), r[i]);
        p[i] = pi, y[i] = dinv[i] * pi;
    }
}

// s = r - alpha v; z = dinv s
__global__ void __launch_bounds__(kVecThreads) k_bicg_s(int64_t n, const double* __restrict__ sc, const double* __restrict__ dinv,
                                                        const double* __restrict__ r, const double* __restrict__ v,
                                                        double* __restrict__ s_, double* __restrict__ z) {
    const double alpha = sc[1];
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
        const double si = fma(-alpha, v[i], r[i]);
        s_[i] = si, z[i] = dinv[i] * si;
    }
}

// partial sums of a . b and a . a
__global__ void __launch_bounds__(kVecThreads) k_dot2_partials(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                               double* __restrict__ part_ab, double* __restrict__ part_aa) {
    __shared__ double red[kVecThreads / 32];
    double ab = 0.0, aa = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
        ab = fma(a[i], b[i], ab), aa = fma(a[i], a[i], aa);
    }
    const double x = block_sum(ab, red), y = block_sum(aa, red);
    if (threadIdx.x == 0) part_ab[blockIdx.x] = x, part_aa[blockIdx.x] = y;
}

// x += alpha y + omega z; r = s - omega t; partial sums of rhat . r and r . r
__global__ void __launch_bounds__(kVecThreads) k_bicg_x(int64_t n, const double* __restrict__ sc, const double* __restrict__ y,
                                                        const double* __restrict__ z, const double* __restrict__ s_,
                                                        const double* __restrict__ t, const double* __restrict__ rhat,
                                                        double* __restrict__ x, double* __restrict__ r,
                                                        double* __restrict__ part_rho, double* __restrict__ part_rr) {
    __shared__ double red[kVecThreads / 32];
    const double alpha = sc[1], omega = sc[2];
    double rho = 0.0, rr = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
        x[i] = fma(alpha, y[i], fma(omega, z[i], x[i]));
        const double ri = fma(-omega, t[i], s_[i]);
        r[i] = ri;
        rho = fma(rhat[i], ri, rho), rr = fma(ri, ri, rr);
    }
    const double a = block_sum(rho, red), b = block_sum(rr, red);
    if (threadIdx.x == 0) part_rho[blockIdx.x] = a, part_rr[blockIdx.x] = b;
}

static int ensure_solve_scratch(MeshDev& d, bool cg) {
    const int64_t ncols = d.ncols_nodes * d.m;
    if (!d.bc_fixed) {
        PFG_CUDA_TRY(cudaMalloc(&d.bc_fixed, std::max<int64_t>(ncols, 1)));
        PFG_CUDA_TRY(cudaMalloc(&d.bc_u0, std::max<int64_t>(ncols, 1) * sizeof(double)));
        d.device_bytes += ncols * 9;
    }
    if (cg && !d.cg_work) {
        const int64_t n = (d.own_end - d.own_begin) * d.m;
        // CG uses five vectors, BiCGStab nine; then the partial-sum arrays and a few device scalars
        PFG_CUDA_TRY(cudaMalloc(&d.cg_work, (9 * std::max<int64_t>(n, 1) + 8 * kMaxPartials + 16) * sizeof(double)));
        d.device_bytes += (9 * n + 8 * kMaxPartials + 16) * (int64_t)sizeof(double);
    }
    return PFG_OK;
}

}  // namespace pfg

using namespace pfg;

#define PFG_CHECK_MESH(mesh)                         \
    if (!(mesh)) {                                   \
        set_error("%s: mesh is NULL", __func__);     \
        return PFG_ERR_INVALID;                      \
    }                                                \
    PFG_CUDA_TRY(cudaSetDevice((mesh)->d.device));

extern "C" int pfg_apply_dirichlet(pfg_mesh* mesh, const int64_t* fixed_dofs_dev, const double* fixed_vals_dev,
                                   int64_t nfixed, int enforce_symmetric, double* vals_dev, double* rhs_dev,
                                   void* stream) {
    PFG_CHECK_MESH(mesh);
    MeshDev& d = mesh->d;
    if (!vals_dev || nfixed < 0 || (nfixed > 0 && !fixed_dofs_dev)) {
        set_error("pfg_apply_dirichlet: invalid argument");
        return PFG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ncols = d.ncols_nodes * d.m;
    PFG_TRY(ensure_solve_scratch(d, false));  // flags / values of the fixed dofs live in the handle
    PFG_CUDA_TRY(cudaMemsetAsync(d.bc_fixed, 0, ncols, st));
    if (nfixed)
        k_mark_fixed<<<(unsigned)((nfixed + 255) / 256), 256, 0, st>>>(fixed_dofs_dev, fixed_vals_dev, nfixed, ncols,
                                                                      d.bc_fixed, d.bc_u0);
    const int64_t nrows = (d.own_end - d.own_begin) * d.m;
    if (nrows)
        k_apply_dirichlet<<<(unsigned)((nrows + 127) / 128), 128, 0, st>>>(
            d.blk_ptr, d.nbr, d.gid, d.own_begin, d.own_end - d.own_begin, d.m, d.bc_fixed, d.bc_u0, enforce_symmetric,
            fixed_vals_dev != nullptr, vals_dev, rhs_dev);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

static unsigned spmv_grid(int64_t nrows, int sm_count) {  // grid-stride kernel: at most eight CTAs per SM
    const int64_t full = std::max<int64_t>(1, (nrows * kRowLanes + 255) / 256);
    return (unsigned)std::min<int64_t>(full, std::min<int64_t>(kMaxPartials, (int64_t)sm_count * 8));
}

extern "C" int pfg_spmv(pfg_mesh* mesh, const double* vals_dev, const double* x_dev, double* y_dev, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (!vals_dev || !x_dev || !y_dev) {
        set_error("pfg_spmv: NULL argument");
        return PFG_ERR_INVALID;
    }
    const int64_t nrows = (d.own_end - d.own_begin) * d.m;
    if (nrows)
        launch_spmv_rows(spmv_grid(nrows, d.sm_count), (cudaStream_t)stream, d, d.gid, vals_dev, x_dev, y_dev, nullptr, nullptr);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

static int whole_matrix(const MeshDev& d, const char* who) {
    if (d.own_begin != 0 || d.own_end != d.nnodes || d.gid != nullptr) {
        set_error("%s: needs a handle that owns every row (a rank's row slab does not hold the transposed entries / the "
                  "whole operator)", who);
        return PFG_ERR_UNSUPPORTED;
    }
    return PFG_OK;
}

extern "C" int pfg_spmv_t(pfg_mesh* mesh, const double* vals_dev, const double* x_dev, double* y_dev, void* stream) {
    PFG_CHECK_MESH(mesh);
    MeshDev& d = mesh->d;
    if (!vals_dev || !x_dev || !y_dev) {
        set_error("pfg_spmv_t: NULL argument");
        return PFG_ERR_INVALID;
    }
    PFG_TRY(whole_matrix(d, "pfg_spmv_t"));
    cudaStream_t st = (cudaStream_t)stream;
    if (!d.trank) {  // once per handle
        PFG_CUDA_TRY(cudaMalloc(&d.trank, std::max<int64_t>(d.nblocks, 1)));
        d.device_bytes += d.nblocks;
        k_transpose_ranks<<<(unsigned)((d.nnodes + 127) / 128), 128, 0, st>>>(d.blk_ptr, d.nbr, d.nnodes, d.trank);
    }
    const int64_t nrows = d.nnodes * d.m;
    const unsigned grid = spmv_grid(nrows, d.sm_count);
    const bool short_rows = d.max_k * d.m <= 24;
    if (d.m == 1 && short_rows) k_spmv_t_rows<1, 4><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, d.trank, d.nnodes, vals_dev, x_dev, y_dev);
    else if (d.m == 1) k_spmv_t_rows<1, 8><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, d.trank, d.nnodes, vals_dev, x_dev, y_dev);
    else if (d.m == 2 && short_rows) k_spmv_t_rows<2, 4><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, d.trank, d.nnodes, vals_dev, x_dev, y_dev);
    else if (d.m == 2) k_spmv_t_rows<2, 8><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, d.trank, d.nnodes, vals_dev, x_dev, y_dev);
    else k_spmv_t_rows<3, 8><<<grid, 256, 0, st>>>(d.blk_ptr, d.nbr, d.trank, d.nnodes, vals_dev, x_dev, y_dev);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_cg(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev, int x_is_zero,
                      double rtol, double atol, int max_iter, int check_every, int* iters_out, double* resid_out,
                      void* stream) {
    PFG_CHECK_MESH(mesh);
    MeshDev& d = mesh->d;
    if (!vals_dev || !b_dev || !x_dev || max_iter < 0) {
        set_error("pfg_cg: invalid argument");
        return PFG_ERR_INVALID;
    }
    PFG_TRY(whole_matrix(d, "pfg_cg"));
    PFG_TRY(ensure_solve_scratch(d, true));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = d.nnodes * d.m, nown = d.nnodes;
    double* r = d.cg_work;
    double *z = r + n, *p = z + n, *Ap = p + n, *dinv = Ap + n;
    double* parts = r + 9 * n;  // [0] p.Ap (one per SpMV CTA, capped), [1..2] r.z ping-pong, [3] r.r, [4] |b|^2
    const unsigned gs = spmv_grid(n, d.sm_count);  // persistent SpMV grid: one p.Ap partial per CTA
    double* part_pAp = parts;
    double* part_rz[2] = {parts + 1 * kMaxPartials, parts + 2 * kMaxPartials};
    double* part_rr = parts + 3 * kMaxPartials;
    double* part_bb = parts + 4 * kMaxPartials;
    const int gv = (int)std::min<int64_t>(kMaxPartials, std::max<int64_t>(1, (n + kVecThreads - 1) / kVecThreads));
    const unsigned g256 = (unsigned)((n + 255) / 256);

    k_inv_diag<<<g256, 256, 0, st>>>(d.blk_ptr, d.nbr, d.own_begin, nown, d.m, vals_dev, dinv);
    if (x_is_zero) {
        PFG_CUDA_TRY(cudaMemsetAsync(x_dev, 0, n * sizeof(double), st));
        PFG_CUDA_TRY(cudaMemcpyAsync(r, b_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    } else {
        launch_spmv_rows(gs, st, d, nullptr, vals_dev, x_dev, Ap, nullptr, nullptr);
        k_residual<<<gv, kVecThreads, 0, st>>>(n, b_dev, Ap, r);
    }
    k_norm2_partials<<<gv, kVecThreads, 0, st>>>(n, b_dev, part_bb);
    k_cg_init<<<gv, kVecThreads, 0, st>>>(n, dinv, r, z, p, part_rz[0], part_rr);
    std::vector<double> h(kMaxPartials);
    auto host_sum = [&](const double* dev, double* out) -> int {
        PFG_CUDA_TRY(cudaMemcpyAsync(h.data(), dev, gv * sizeof(double), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        double s = 0.0;
        for (int i = 0; i < gv; ++i) s += h[i];
        *out = s;
        return PFG_OK;
    };
    double bb = 0.0, rr = 0.0;
    PFG_TRY(host_sum(part_bb, &bb));
    PFG_TRY(host_sum(part_rr, &rr));
    const double target = std::max(rtol * std::sqrt(bb), atol);  // scipy's cg: |r| <= max(rtol |b|, atol)
    int it = 0;
    if (check_every <= 0) check_every = 16;
    while (std::sqrt(rr) > target && it < max_iter) {
        const int batch = std::min(check_every, max_iter - it);
        for (int j = 0; j < batch; ++j, ++it) {
            const int cur = it & 1;
            launch_spmv_rows(gs, st, d, nullptr, vals_dev, p, Ap, p, part_pAp);
            k_cg_update<<<gv, kVecThreads, 0, st>>>(n, (int)gs, gv, part_pAp, part_rz[cur], dinv, p, Ap, x_dev, r, z,
                                                    part_rz[cur ^ 1], part_rr);
            k_cg_direction<<<gv, kVecThreads, 0, st>>>(n, gv, part_rz[cur ^ 1], part_rz[cur], z, p);
        }
        PFG_CUDA_TRY(cudaGetLastError());
        PFG_TRY(host_sum(part_rr, &rr));
        if (!(rr == rr)) {
            set_error("pfg_cg: the residual became NaN after %d iterations (matrix not positive definite?)", it);
            return PFG_ERR_INVALID;
        }
    }
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = std::sqrt(rr);
    PFG_CUDA_TRY(cudaGetLastError());
    return (std::sqrt(rr) <= target) ? PFG_OK : PFG_ERR_NOCONV;
}

// Conjugate gradients over the row slabs of several ranks: the same kernels as pfg_cg; every dot product is summed
// on the rank in a fixed order, written to one slot of scal_dev and all-reduced by the caller's callback (the library
// itself does not link NCCL), the search direction lives in a global-length vector whose ghost entries the halo
// callback refreshes before every product.  Two phases that only ENQUEUE work (no host synchronisation, so a caller
// may capture the steps in a CUDA graph and replay them), and pfg_cg_dist = begin + steps + the convergence check.
namespace {
struct CgDist {  // scal_dev: [0] p.Ap  [1] / [3] r.z (alternating by iteration parity)  [2] r.r  [4] |b|^2
    MeshDev* d;
    int64_t n, nown;
    double *r, *z, *Ap, *dinv, *p, *part_a, *part_b, *part_c;
    unsigned gs;
    int gv;
};

int cg_dist_setup(pfg_mesh* mesh, const void* vals, const void* x, double* x_full, const void* scal, int64_t row0,
                  const void* reduce, const void* halo, const char* who, CgDist* c) {
    MeshDev& d = mesh->d;
    const int64_t nown = d.own_end - d.own_begin, n = nown * d.m, ncols = d.ncols_nodes * d.m;
    if (!vals || !x || !x_full || !scal || !reduce || !halo || row0 < 0 || row0 + n > ncols) {
        set_error("%s: invalid argument", who);
        return PFG_ERR_INVALID;
    }
    PFG_TRY(ensure_solve_scratch(d, true));
    c->d = &d, c->n = n, c->nown = nown;
    c->r = d.cg_work;
    c->z = c->r + n, c->Ap = c->r + 3 * n, c->dinv = c->r + 4 * n;  // (the slot pfg_cg uses for p stays free: p lives in x_full)
    c->p = x_full + row0;
    double* parts = c->r + 9 * n;
    c->part_a = parts, c->part_b = parts + kMaxPartials, c->part_c = parts + 2 * kMaxPartials;
    c->gs = spmv_grid(std::max<int64_t>(n, 1), d.sm_count);
    c->gv = (int)std::min<int64_t>(kMaxPartials, std::max<int64_t>(1, (n + kVecThreads - 1) / kVecThreads));
    return PFG_OK;
}

int cg_dist_cb(int rc, const char* what) {
    if (rc != 0) {
        set_error("pfg_cg_dist: the %s callback failed (%d)", what, rc);
        return PFG_ERR_INVALID;
    }
    return PFG_OK;
}

int cg_dist_begin(const CgDist& c, const double* vals, const double* b, double* x, int x_is_zero, double* x_full,
                  double* scal, pfg_reduce_fn reduce, pfg_halo_fn halo, void* user, cudaStream_t st) {
    const MeshDev& d = *c.d;
    const int64_t n = c.n;
    if (n) k_inv_diag<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.blk_ptr, d.nbr, d.own_begin, c.nown, d.m, vals, c.dinv);
    if (x_is_zero) {
        PFG_CUDA_TRY(cudaMemsetAsync(x, 0, n * sizeof(double), st));
        PFG_CUDA_TRY(cudaMemcpyAsync(c.r, b, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    } else {
        PFG_CUDA_TRY(cudaMemcpyAsync(c.p, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        PFG_TRY(cg_dist_cb(halo(user), "halo"));
        if (n) launch_spmv_rows(c.gs, st, d, d.gid, vals, x_full, c.Ap, nullptr, nullptr);
        k_residual<<<c.gv, kVecThreads, 0, st>>>(n, b, c.Ap, c.r);
    }
    k_norm2_partials<<<c.gv, kVecThreads, 0, st>>>(n, b, c.part_a);
    k_sum_into<<<1, kVecThreads, 0, st>>>(c.part_a, c.gv, scal + 4);
    k_cg_init<<<c.gv, kVecThreads, 0, st>>>(n, c.dinv, c.r, c.z, c.p, c.part_b, c.part_c);
    k_sum_into<<<1, kVecThreads, 0, st>>>(c.part_b, c.gv, scal + 1);
    k_sum_into<<<1, kVecThreads, 0, st>>>(c.part_c, c.gv, scal + 2);
    PFG_CUDA_TRY(cudaGetLastError());
    PFG_TRY(cg_dist_cb(reduce(user, 1, 2), "reduce"));
    PFG_TRY(cg_dist_cb(reduce(user, 4, 1), "reduce"));
    return PFG_OK;
}

int cg_dist_steps(const CgDist& c, const double* vals, double* x, double* x_full, double* scal, int first_iter,
                  int n_steps, pfg_reduce_fn reduce, pfg_halo_fn halo, void* user, cudaStream_t st) {
    const MeshDev& d = *c.d;
    const int64_t n = c.n;
    for (int it = first_iter; it < first_iter + n_steps; ++it) {
        const int rz_old = (it & 1) ? 3 : 1, rz_new = (it & 1) ? 1 : 3;
        PFG_TRY(cg_dist_cb(halo(user), "halo"));
        if (n) launch_spmv_rows(c.gs, st, d, d.gid, vals, x_full, c.Ap, c.p, c.part_a);
        k_sum_into<<<1, kVecThreads, 0, st>>>(c.part_a, n ? (int)c.gs : 0, scal + 0);
        PFG_TRY(cg_dist_cb(reduce(user, 0, 1), "reduce"));
        k_cg_update<<<c.gv, kVecThreads, 0, st>>>(n, 1, 1, scal + 0, scal + rz_old, c.dinv, c.p, c.Ap, x, c.r, c.z,
                                                  c.part_b, c.part_c);
        k_sum_into<<<1, kVecThreads, 0, st>>>(c.part_b, c.gv, scal + rz_new);
        k_sum_into<<<1, kVecThreads, 0, st>>>(c.part_c, c.gv, scal + 2);
        PFG_TRY(cg_dist_cb(reduce(user, std::min(rz_new, 2), 2), "reduce"));  // slots (1, 2) or (2, 3)
        k_cg_direction<<<c.gv, kVecThreads, 0, st>>>(n, 1, scal + rz_new, scal + rz_old, c.z, c.p);
    }
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}
}  // namespace

extern "C" int pfg_cg_dist_begin(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev,
                                 int x_is_zero, double* x_full_dev, double* scal_dev, int64_t row0,
                                 pfg_reduce_fn reduce, pfg_halo_fn halo, void* user, void* stream) {
    PFG_CHECK_MESH(mesh);
    CgDist c;
    if (!b_dev) {
        set_error("pfg_cg_dist_begin: invalid argument");
        return PFG_ERR_INVALID;
    }
    PFG_TRY(cg_dist_setup(mesh, vals_dev, x_dev, x_full_dev, scal_dev, row0, (const void*)reduce, (const void*)halo,
                          "pfg_cg_dist_begin", &c));
    return cg_dist_begin(c, vals_dev, b_dev, x_dev, x_is_zero, x_full_dev, scal_dev, reduce, halo, user, (cudaStream_t)stream);
}

extern "C" int pfg_cg_dist_steps(pfg_mesh* mesh, const double* vals_dev, double* x_dev, double* x_full_dev,
                                 double* scal_dev, int64_t row0, int first_iter, int n_steps, pfg_reduce_fn reduce,
                                 pfg_halo_fn halo, void* user, void* stream) {
    PFG_CHECK_MESH(mesh);
    CgDist c;
    if (first_iter < 0 || n_steps < 0) {
        set_error("pfg_cg_dist_steps: invalid argument");
        return PFG_ERR_INVALID;
    }
    PFG_TRY(cg_dist_setup(mesh, vals_dev, x_dev, x_full_dev, scal_dev, row0, (const void*)reduce, (const void*)halo,
                          "pfg_cg_dist_steps", &c));
    return cg_dist_steps(c, vals_dev, x_dev, x_full_dev, scal_dev, first_iter, n_steps, reduce, halo, user, (cudaStream_t)stream);
}

extern "C" int pfg_cg_dist(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev, int x_is_zero,
                           double* x_full_dev, double* scal_dev, int64_t row0, double rtol, double atol, int max_iter,
                           int check_every, pfg_reduce_fn reduce, pfg_halo_fn halo, void* user, int* iters_out,
                           double* resid_out, void* stream) {
    PFG_CHECK_MESH(mesh);
    CgDist c;
    if (!b_dev || max_iter < 0) {
        set_error("pfg_cg_dist: invalid argument");
        return PFG_ERR_INVALID;
    }
    PFG_TRY(cg_dist_setup(mesh, vals_dev, x_dev, x_full_dev, scal_dev, row0, (const void*)reduce, (const void*)halo,
                          "pfg_cg_dist", &c));
    cudaStream_t st = (cudaStream_t)stream;
    auto read_scal = [&](int slot, double* out) -> int {
        PFG_CUDA_TRY(cudaMemcpyAsync(out, scal_dev + slot, sizeof(double), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        return PFG_OK;
    };
    PFG_TRY(cg_dist_begin(c, vals_dev, b_dev, x_dev, x_is_zero, x_full_dev, scal_dev, reduce, halo, user, st));
    double bb = 0.0, rr = 0.0;
    PFG_TRY(read_scal(4, &bb));
    PFG_TRY(read_scal(2, &rr));
    const double target = std::max(rtol * std::sqrt(bb), atol);  // scipy's cg: |r| <= max(rtol |b|, atol), global norms
    int it = 0;
    if (check_every <= 0) check_every = 16;
    while (std::sqrt(rr) > target && it < max_iter) {  // rr is the same number on every rank: so is the trip count
        const int batch = std::min(check_every, max_iter - it);
        PFG_TRY(cg_dist_steps(c, vals_dev, x_dev, x_full_dev, scal_dev, it, batch, reduce, halo, user, st));
        it += batch;
        PFG_TRY(read_scal(2, &rr));
        if (!(rr == rr)) {
            set_error("pfg_cg_dist: the residual became NaN after %d iterations (matrix not positive definite?)", it);
            return PFG_ERR_INVALID;
        }
    }
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = std::sqrt(rr);
    PFG_CUDA_TRY(cudaGetLastError());
    return (std::sqrt(rr) <= target) ? PFG_OK : PFG_ERR_NOCONV;
}

extern "C" int pfg_bicgstab(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev, double rtol,
                            double atol, int max_iter, int check_every, int* iters_out, double* resid_out, void* stream) {
    PFG_CHECK_MESH(mesh);
    MeshDev& d = mesh->d;
    if (!vals_dev || !b_dev || !x_dev || max_iter < 0) {
        set_error("pfg_bicgstab: invalid argument");
        return PFG_ERR_INVALID;
    }
    PFG_TRY(whole_matrix(d, "pfg_bicgstab"));
    PFG_TRY(ensure_solve_scratch(d, true));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = d.nnodes * d.m, nown = d.nnodes;
    double* r = d.cg_work;
    double *rhat = r + n, *p = rhat + n, *v = p + n, *s_ = v + n, *t = s_ + n, *y = t + n, *z = y + n, *dinv = z + n;
    double* parts = r + 9 * n;
    double *part_a = parts, *part_b = parts + kMaxPartials, *part_rho = parts + 2 * kMaxPartials,
           *part_rr = parts + 3 * kMaxPartials, *part_bb = parts + 4 * kMaxPartials, *sc = parts + 8 * kMaxPartials;
    const unsigned gs = spmv_grid(n, d.sm_count);
    const int gv = (int)std::min<int64_t>(kMaxPartials, std::max<int64_t>(1, (n + kVecThreads - 1) / kVecThreads));
    k_inv_diag<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.blk_ptr, d.nbr, d.own_begin, nown, d.m, vals_dev, dinv);
    // x0 = 0: r = rhat = b, p = v = 0
    PFG_CUDA_TRY(cudaMemsetAsync(x_dev, 0, n * sizeof(double), st));
    PFG_CUDA_TRY(cudaMemcpyAsync(r, b_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PFG_CUDA_TRY(cudaMemcpyAsync(rhat, b_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PFG_CUDA_TRY(cudaMemsetAsync(p, 0, n * sizeof(double), st));
    PFG_CUDA_TRY(cudaMemsetAsync(v, 0, n * sizeof(double), st));
    k_norm2_partials<<<gv, kVecThreads, 0, st>>>(n, b_dev, part_bb);
    k_dot2_partials<<<gv, kVecThreads, 0, st>>>(n, r, rhat, part_rho, part_rr);
    k_bicg_scalar<<<1, 32, 0, st>>>(0, gv, part_rho, 0, nullptr, sc);
    std::vector<double> h(kMaxPartials);
    auto host_sum = [&](const double* dev, double* out) -> int {
        PFG_CUDA_TRY(cudaMemcpyAsync(h.data(), dev, gv * sizeof(double), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        double sum = 0.0;
        for (int i = 0; i < gv; ++i) sum += h[i];
        *out = sum;
        return PFG_OK;
    };
    double bb = 0.0, rr = 0.0;
    PFG_TRY(host_sum(part_bb, &bb));
    rr = bb;
    const double target = std::max(rtol * std::sqrt(bb), atol);
    int it = 0;
    if (check_every <= 0) check_every = 8;
    while (std::sqrt(rr) > target && it < max_iter) {
        const int batch = std::min(check_every, max_iter - it);
        for (int j = 0; j < batch; ++j, ++it) {
            k_bicg_p<<<gv, kVecThreads, 0, st>>>(n, sc, dinv, r, v, p, y);
            launch_spmv_rows(gs, st, d, nullptr, vals_dev, y, v, rhat, part_a);
            k_bicg_scalar<<<1, 32, 0, st>>>(1, (int)gs, part_a, 0, nullptr, sc);
            k_bicg_s<<<gv, kVecThreads, 0, st>>>(n, sc, dinv, r, v, s_, z);
            launch_spmv_rows(gs, st, d, nullptr, vals_dev, z, t, nullptr, nullptr);
            k_dot2_partials<<<gv, kVecThreads, 0, st>>>(n, t, s_, part_a, part_b);
            k_bicg_scalar<<<1, 32, 0, st>>>(2, gv, part_a, gv, part_b, sc);
            k_bicg_x<<<gv, kVecThreads, 0, st>>>(n, sc, y, z, s_, t, rhat, x_dev, r, part_rho, part_rr);
            k_bicg_scalar<<<1, 32, 0, st>>>(3, gv, part_rho, 0, nullptr, sc);
        }
        PFG_CUDA_TRY(cudaGetLastError());
        PFG_TRY(host_sum(part_rr, &rr));
        if (!(rr == rr)) {
            set_error("pfg_bicgstab: the residual became NaN after %d iterations (breakdown)", it);
            return PFG_ERR_INVALID;
        }
    }
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = std::sqrt(rr);
    PFG_CUDA_TRY(cudaGetLastError());
    return (std::sqrt(rr) <= target) ? PFG_OK : PFG_ERR_NOCONV;
}

// BiCGStab over the row slabs of several ranks: pfg_bicgstab's kernels; the two preconditioned vectors y and z live in
// global-length vectors (x_full_dev holds both, ncols doubles each) whose ghost entries halo(user, which) refreshes
// before the product that reads them, the five dot products per iteration travel in three all-reduces.
extern "C" int pfg_bicgstab_dist(pfg_mesh* mesh, const double* vals_dev, const double* b_dev, double* x_dev,
                                 double* x_full_dev, double* scal_dev, int64_t row0, double rtol, double atol,
                                 int max_iter, int check_every, pfg_reduce_fn reduce, pfg_halo2_fn halo, void* user,
                                 int* iters_out, double* resid_out, void* stream) {
    PFG_CHECK_MESH(mesh);
    MeshDev& d = mesh->d;
    const int64_t nown = d.own_end - d.own_begin, n = nown * d.m, ncols = d.ncols_nodes * d.m;
    if (!vals_dev || !b_dev || !x_dev || !x_full_dev || !scal_dev || !reduce || !halo || max_iter < 0 || row0 < 0 ||
        row0 + n > ncols) {
        set_error("pfg_bicgstab_dist: invalid argument");
        return PFG_ERR_INVALID;
    }
    PFG_TRY(ensure_solve_scratch(d, true));
    cudaStream_t st = (cudaStream_t)stream;
    double* r = d.cg_work;
    double *rhat = r + n, *p = rhat + n, *v = p + n, *s_ = v + n, *t = s_ + n, *dinv = t + 3 * n;
    double *y_full = x_full_dev, *z_full = x_full_dev + ncols;
    double *y = y_full + row0, *z = z_full + row0;
    double* parts = r + 9 * n;
    double *part_a = parts, *part_b = parts + kMaxPartials, *sc = parts + 8 * kMaxPartials;
    const unsigned gs = spmv_grid(std::max<int64_t>(n, 1), d.sm_count);
    const int gv = (int)std::min<int64_t>(kMaxPartials, std::max<int64_t>(1, (n + kVecThreads - 1) / kVecThreads));
    // scal_dev: [0] rhat.v  [1] t.s  [2] t.t  [3] rhat.r  [4] r.r  [5] |b|^2
    auto cb = [&](int rc, const char* what) -> int {
        if (rc != 0) {
            set_error("pfg_bicgstab_dist: the %s callback failed (%d)", what, rc);
            return PFG_ERR_INVALID;
        }
        return PFG_OK;
    };
    auto sum_into = [&](const double* part, int np, int slot) { k_sum_into<<<1, kVecThreads, 0, st>>>(part, np, scal_dev + slot); };
    auto read_scal = [&](int slot, double* out) -> int {
        PFG_CUDA_TRY(cudaMemcpyAsync(out, scal_dev + slot, sizeof(double), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        return PFG_OK;
    };
    if (n) k_inv_diag<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.blk_ptr, d.nbr, d.own_begin, nown, d.m, vals_dev, dinv);
    PFG_CUDA_TRY(cudaMemsetAsync(x_dev, 0, n * sizeof(double), st));
    PFG_CUDA_TRY(cudaMemcpyAsync(r, b_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PFG_CUDA_TRY(cudaMemcpyAsync(rhat, b_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PFG_CUDA_TRY(cudaMemsetAsync(p, 0, n * sizeof(double), st));
    PFG_CUDA_TRY(cudaMemsetAsync(v, 0, n * sizeof(double), st));
    k_norm2_partials<<<gv, kVecThreads, 0, st>>>(n, b_dev, part_a);
    sum_into(part_a, gv, 5);
    k_dot2_partials<<<gv, kVecThreads, 0, st>>>(n, r, rhat, part_a, part_b);
    sum_into(part_a, gv, 3), sum_into(part_b, gv, 4);
    PFG_CUDA_TRY(cudaGetLastError());
    PFG_TRY(cb(reduce(user, 3, 3), "reduce"));
    k_bicg_scalar<<<1, 32, 0, st>>>(0, 1, scal_dev + 3, 0, nullptr, sc);
    double bb = 0.0, rr = 0.0;
    PFG_TRY(read_scal(5, &bb));
    rr = bb;
    const double target = std::max(rtol * std::sqrt(bb), atol);
    int it = 0;
    if (check_every <= 0) check_every = 8;
    while (std::sqrt(rr) > target && it < max_iter) {
        const int batch = std::min(check_every, max_iter - it);
        for (int j = 0; j < batch; ++j, ++it) {
            k_bicg_p<<<gv, kVecThreads, 0, st>>>(n, sc, dinv, r, v, p, y);
            PFG_TRY(cb(halo(user, 0), "halo"));
            if (n) launch_spmv_rows(gs, st, d, d.gid, vals_dev, y_full, v, rhat, part_a);
            sum_into(part_a, n ? (int)gs : 0, 0);
            PFG_TRY(cb(reduce(user, 0, 1), "reduce"));
            k_bicg_scalar<<<1, 32, 0, st>>>(1, 1, scal_dev + 0, 0, nullptr, sc);
            k_bicg_s<<<gv, kVecThreads, 0, st>>>(n, sc, dinv, r, v, s_, z);
            PFG_TRY(cb(halo(user, 1), "halo"));
            if (n) launch_spmv_rows(gs, st, d, d.gid, vals_dev, z_full, t, nullptr, nullptr);
            k_dot2_partials<<<gv, kVecThreads, 0, st>>>(n, t, s_, part_a, part_b);
            sum_into(part_a, gv, 1), sum_into(part_b, gv, 2);
            PFG_TRY(cb(reduce(user, 1, 2), "reduce"));
            k_bicg_scalar<<<1, 32, 0, st>>>(2, 1, scal_dev + 1, 1, scal_dev + 2, sc);
            k_bicg_x<<<gv, kVecThreads, 0, st>>>(n, sc, y, z, s_, t, rhat, x_dev, r, part_a, part_b);
            sum_into(part_a, gv, 3), sum_into(part_b, gv, 4);
            PFG_TRY(cb(reduce(user, 3, 2), "reduce"));
            k_bicg_scalar<<<1, 32, 0, st>>>(3, 1, scal_dev + 3, 0, nullptr, sc);
        }
        PFG_CUDA_TRY(cudaGetLastError());
        PFG_TRY(read_scal(4, &rr));
        if (!(rr == rr)) {
            set_error("pfg_bicgstab_dist: the residual became NaN after %d iterations (breakdown)", it);
            return PFG_ERR_INVALID;
        }
    }
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = std::sqrt(rr);
    PFG_CUDA_TRY(cudaGetLastError());
    return (std::sqrt(rr) <= target) ? PFG_OK : PFG_ERR_NOCONV;
}
