// Per-element quadrature in registers: geometry + element matrices / vectors for quad4 and hex8.
//
// Math follows SURVEY.md Appendix A (reference utils.py:171-264 for the geometry; pyfem.py:1177-1185,
// 1132-1134, 2017-2026, 2127-2136, 1462-1471, 1530-1537, 1595-1609 for the element einsums), restated
// for one thread per element with everything unrolled so that basis values are immediates:
//   G_a = det(J) * grad N_a   (adjugate form, one reciprocal per quadrature point instead of
//                              dividing every inverse entry);
//   weights collapse to s_q = w_q c_q / det(J_q)  for gradient-gradient terms.
// Results go to a Sink (atomic scatter or shared-memory staging) one node-pair block at a time.
#pragma once
#include <utility>

#include "pfg_internal.cuh"

namespace pfg {

#define PFG_DEV __device__ __forceinline__
#define PFG_G 0.57735026918962576451  /* 1/sqrt(3) */

// Reciprocal to within a couple of ulp: hardware seed (rcp.approx.ftz.f64, ~2^-23) + two Newton steps.
// Replaces the IEEE division sequence (slow path + fix-up) at every quadrature point; the parity bound
// (1e-12 of max|K|) leaves six orders of magnitude of slack.  det(J) of a valid element is never subnormal.
PFG_DEV double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// ---------------------------------------------------------------------------------------------
// Reference element tables as constexpr functions (folded to immediates after unrolling)
// ---------------------------------------------------------------------------------------------
template <int NNE>
struct Elem;

template <>
struct Elem<4> {  // BasisBilinear2D / QuadratureBilinear2D (pyfem.py:83-95, 253-284)
    static constexpr int DIM = 2, NQ = 4;
    PFG_DEV static constexpr double sgn(int a, int k) {  // local node coordinates
        return k == 0 ? ((a == 1 || a == 2) ? 1.0 : -1.0) : ((a >= 2) ? 1.0 : -1.0);
    }
    PFG_DEV static constexpr double qp(int q, int k) {  // quadrature points, CCW order
        return sgn(q, k) * PFG_G;
    }
    PFG_DEV static constexpr double N(int q, int a) {
        return 0.25 * (1.0 + sgn(a, 0) * qp(q, 0)) * (1.0 + sgn(a, 1) * qp(q, 1));
    }
    PFG_DEV static constexpr double dN(int q, int a, int k) {
        return k == 0 ? 0.25 * sgn(a, 0) * (1.0 + sgn(a, 1) * qp(q, 1))
                      : 0.25 * sgn(a, 1) * (1.0 + sgn(a, 0) * qp(q, 0));
    }
};

template <>
struct Elem<8> {  // BasisBlock3D / QuadratureBlock3D (pyfem.py:97-112, 287-338)
    static constexpr int DIM = 3, NQ = 8;
    PFG_DEV static constexpr double sgn(int a, int k) {
        return k == 0 ? (((a & 3) == 1 || (a & 3) == 2) ? 1.0 : -1.0)
                      : (k == 1 ? (((a & 3) >= 2) ? 1.0 : -1.0) : ((a >= 4) ? 1.0 : -1.0));
    }
    PFG_DEV static constexpr double qp(int q, int k) {  // x slowest, z fastest
        return (k == 0 ? ((q & 4) ? 1.0 : -1.0) : (k == 1 ? ((q & 2) ? 1.0 : -1.0) : ((q & 1) ? 1.0 : -1.0))) * PFG_G;
    }
    PFG_DEV static constexpr double N(int q, int a) {
        return 0.125 * (1.0 + sgn(a, 0) * qp(q, 0)) * (1.0 + sgn(a, 1) * qp(q, 1)) * (1.0 + sgn(a, 2) * qp(q, 2));
    }
    PFG_DEV static constexpr double dN(int q, int a, int k) {
        return 0.125 * sgn(a, k) * (1.0 + sgn(a, (k + 1) % 3) * qp(q, (k + 1) % 3)) *
               (1.0 + sgn(a, (k + 2) % 3) * qp(q, (k + 2) % 3));
    }
};

// ---------------------------------------------------------------------------------------------
// What a kernel hands the element routines
// ---------------------------------------------------------------------------------------------
struct MeshView {
    const double* X;
    int64_t own_begin, own_end;
    // atomic path
    const int32_t* conn;
    const int64_t* blk_ptr;
    const uint8_t* rank;
    const uint8_t* elem_skip;  // nullptr: every element is integrated
    int64_t nelems;
    // gather path
    const ChunkHdr* chunks;
    const ChunkNode* cnodes;
    const int32_t* cnode_id;
    const int32_t* rec_nodes;
    const uint16_t* rec_dst;
    const int32_t* rec_elem;
    const uint8_t* plan_pool;
    // tile plan (second format)
    const TileDir* tile_dir;
    const uint8_t* tile_blob;
    const uint16_t* tile_codes;
    const uint32_t* win_nodes;
    const uint16_t* rec_local;
    const uint8_t* rec_skip;  // per element record: 1 = masked (nullptr: no mask)
    // shared-memory staging sizes (bytes) for the gather kernels: chunk node table, chunk plan
    int stage_nodes_bytes, stage_plan_bytes;
};

template <int NNE>
PFG_DEV void load_coords(const double* __restrict__ X, const int (&nodes)[NNE], double (&xe)[NNE][Elem<NNE>::DIM]) {
    if constexpr (NNE == 4) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double2 v = __ldg(reinterpret_cast<const double2*>(X) + nodes[a]);
            xe[a][0] = v.x;
            xe[a][1] = v.y;
        }
    } else {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const double* p = X + (int64_t)nodes[a] * 3;
            xe[a][0] = __ldg(p);
            xe[a][1] = __ldg(p + 1);
            xe[a][2] = __ldg(p + 2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Geometry at one quadrature point: det(J) and G_a = det(J) grad N_a
// (J[j][k] = sum_a dN_a/dxi_k X_a[j], utils.py:184; adjugate = det * inverse of utils.py:243-260)
// ---------------------------------------------------------------------------------------------
struct Quad4Coef {  // dx/dxi = ax + cx*eta, dx/deta = bx + cx*xi (bilinear map), same for y
    double ax, bx, cx, ay, by, cy;
};

PFG_DEV Quad4Coef quad4_coef(const double (&xe)[4][2]) {
    Quad4Coef c;
    c.ax = 0.25 * ((xe[1][0] - xe[0][0]) + (xe[2][0] - xe[3][0]));
    c.bx = 0.25 * ((xe[3][0] - xe[0][0]) + (xe[2][0] - xe[1][0]));
    c.cx = 0.25 * ((xe[0][0] - xe[1][0]) + (xe[2][0] - xe[3][0]));
    c.ay = 0.25 * ((xe[1][1] - xe[0][1]) + (xe[2][1] - xe[3][1]));
    c.by = 0.25 * ((xe[3][1] - xe[0][1]) + (xe[2][1] - xe[1][1]));
    c.cy = 0.25 * ((xe[0][1] - xe[1][1]) + (xe[2][1] - xe[3][1]));
    return c;
}

template <int Q>
PFG_DEV void quad4_geo(const Quad4Coef& c, double& det, double (&G)[4][2]) {
    constexpr double xi = Elem<4>::qp(Q, 0), eta = Elem<4>::qp(Q, 1);
    const double xxi = fma(c.cx, eta, c.ax), xeta = fma(c.cx, xi, c.bx);
    const double yxi = fma(c.cy, eta, c.ay), yeta = fma(c.cy, xi, c.by);
    det = xxi * yeta - xeta * yxi;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double dxi = Elem<4>::dN(Q, a, 0), deta = Elem<4>::dN(Q, a, 1);
        G[a][0] = dxi * yeta - deta * yxi;
        G[a][1] = deta * xxi - dxi * xeta;
    }
}

// bilinear nodal field on the reference square: f = (m + a xi + b eta + c xi eta) / 4 (the form quad4_coef gives the
// map, without its factor 1/4: the callers fold it into their scale)
struct Quad4Field {
    double m, a, b, c;
};
PFG_DEV Quad4Field quad4_field4(double f0, double f1, double f2, double f3) {
    const double d10 = f1 - f0, d23 = f2 - f3;
    Quad4Field r;
    r.m = (f0 + f1) + (f2 + f3);
    r.a = d10 + d23;
    r.b = (f3 - f0) + (f2 - f1);
    r.c = d23 - d10;
    return r;
}

// trilinear nodal field on the reference cube: 8 f = m + cx xi + cy eta + cz zeta + cxy xi eta + cyz eta zeta +
// cxz xi zeta + cxyz xi eta zeta; coefficients by a three-stage butterfly over the corner values (node order of
// Elem<8>: (-,-,-), (+,-,-), (+,+,-), (-,+,-), then the same with zeta = +1)
struct Hex8Field {
    double m, cx, cy, cz, cxy, cyz, cxz, cxyz;
};
PFG_DEV Hex8Field hex8_field8(double f0, double f1, double f2, double f3, double f4, double f5, double f6, double f7) {
    const double s00 = f1 + f0, d00 = f1 - f0, s10 = f2 + f3, d10 = f2 - f3;  // along xi, at (eta, zeta) = (-,-), (+,-)
    const double s01 = f5 + f4, d01 = f5 - f4, s11 = f6 + f7, d11 = f6 - f7;  // ... (-,+), (+,+)
    const double ss0 = s10 + s00, sd0 = s10 - s00, ds0 = d10 + d00, dd0 = d10 - d00;  // along eta, zeta = -1
    const double ss1 = s11 + s01, sd1 = s11 - s01, ds1 = d11 + d01, dd1 = d11 - d01;  // zeta = +1
    Hex8Field r;
    r.m = ss1 + ss0, r.cz = ss1 - ss0;
    r.cy = sd1 + sd0, r.cyz = sd1 - sd0;
    r.cx = ds1 + ds0, r.cxz = ds1 - ds0;
    r.cxy = dd1 + dd0, r.cxyz = dd1 - dd0;
    return r;
}
// 8 * (df/dxi, df/deta, df/dzeta) at quadrature point Q
template <int Q>
PFG_DEV void hex8_field_grad8(const Hex8Field& f, double (&g)[3]) {
    constexpr double xi = Elem<8>::qp(Q, 0), eta = Elem<8>::qp(Q, 1), zeta = Elem<8>::qp(Q, 2);
    g[0] = fma(f.cxyz, eta * zeta, fma(f.cxz, zeta, fma(f.cxy, eta, f.cx)));
    g[1] = fma(f.cxyz, xi * zeta, fma(f.cyz, zeta, fma(f.cxy, xi, f.cy)));
    g[2] = fma(f.cxyz, xi * eta, fma(f.cxz, xi, fma(f.cyz, eta, f.cz)));
}

template <int Q>
PFG_DEV void hex8_geo(const double (&xe)[8][3], double& det, double (&G)[8][3]) {
    double J[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < 8; ++a) s = fma(Elem<8>::dN(Q, a, k), xe[a][j], s);
            J[j][k] = s;
        }
    double A[3][3];  // adjugate: A = det * inv(J)
    A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    A[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
    A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    A[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]);
    A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    A[1][2] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]);
    A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    A[2][1] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]);
    A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int l = 0; l < 3; ++l)
            G[a][l] = Elem<8>::dN(Q, a, 0) * A[0][l] + Elem<8>::dN(Q, a, 1) * A[1][l] + Elem<8>::dN(Q, a, 2) * A[2][l];
}

template <int NNE>
struct GeoCtx;
template <>
struct GeoCtx<4> {
    Quad4Coef c;
    PFG_DEV explicit GeoCtx(const double (&xe)[4][2]) : c(quad4_coef(xe)) {}
    template <int Q>
    PFG_DEV void at(const double (&)[4][2], double& det, double (&G)[4][2]) const {
        quad4_geo<Q>(c, det, G);
    }
};
template <>
struct GeoCtx<8> {
    PFG_DEV explicit GeoCtx(const double (&)[8][3]) {}
    template <int Q>
    PFG_DEV void at(const double (&xe)[8][3], double& det, double (&G)[8][3]) const {
        hex8_geo<Q>(xe, det, G);
    }
};

// compile-time loop over quadrature points
template <class F, int... Qs>
PFG_DEV void for_each_q_impl(F&& f, std::integer_sequence<int, Qs...>) {
    (f(std::integral_constant<int, Qs>{}), ...);
}
template <int NQ, class F>
PFG_DEV void for_each_q(F&& f) {
    for_each_q_impl(f, std::make_integer_sequence<int, NQ>{});
}

template <int NNE, int Q>
PFG_DEV double interp(const double (&f)[NNE]) {  // node -> quadrature point (utils.py:218)
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < NNE; ++a) s = fma(Elem<NNE>::N(Q, a), f[a], s);
    return s;
}

// RAMP-penalised material factor at the quadrature points (pyfem.py:1294-1300, 1938-1944)
struct Material {
    const double* rho;  // nodal field or nullptr
    double rho_const, p;
    double c_const;  // RAMP factor of the constant field, rho_const / (1 + p (1 - rho_const)), formed once on the host
    // complex-step verification (pyfem.py:1018-1020, 1289-1292): imaginary part of a complex nodal field, and which
    // part of the complex RAMP factor this pass integrates (0 real, 1 imaginary).  nullptr on the real path.
    const double* rho_im = nullptr;
    int part = 0;
};

// c = rho / (1 + p (1 - rho)) for complex rho = rr + i ri, real or imaginary part
PFG_DEV double ramp_complex_part(double rr, double ri, double p, int part) {
    const double dr = fma(p, 1.0 - rr, 1.0), di = -p * ri;  // denominator
    const double inv = 1.0 / (dr * dr + di * di);
    return (part == 0 ? (rr * dr + ri * di) : (ri * dr - rr * di)) * inv;
}

template <int NNE>
PFG_DEV void material_at_quads_complex(const Material& mat, const double (&re)[NNE], const double (&im)[NNE],
                                       double (&cq)[Elem<NNE>::NQ]) {
    for_each_q<Elem<NNE>::NQ>([&](auto qc) {
        constexpr int Q = decltype(qc)::value;
        cq[Q] = ramp_complex_part(interp<NNE, Q>(re), interp<NNE, Q>(im), mat.p, mat.part);
    });
}

template <int NNE>
PFG_DEV void material_at_quads(const Material& mat, const double (&re)[NNE], double (&cq)[Elem<NNE>::NQ]) {
    if (mat.rho == nullptr) {
        const double c = mat.c_const;  // uniform field
#pragma unroll
        for (int q = 0; q < Elem<NNE>::NQ; ++q) cq[q] = c;
        return;
    }
    for_each_q<Elem<NNE>::NQ>([&](auto qc) {
        constexpr int Q = decltype(qc)::value;
        const double rq = interp<NNE, Q>(re);
        cq[Q] = rq * fast_rcp(fma(mat.p, 1.0 - rq, 1.0));
    });
}

template <int NNE>
PFG_DEV void load_field(const double* __restrict__ f, const int (&nodes)[NNE], double (&fe)[NNE]) {
#pragma unroll
    for (int a = 0; a < NNE; ++a) fe[a] = (f != nullptr) ? __ldg(f + nodes[a]) : 0.0;
}

// Partition of unity (sum_a grad N_a = 0): a gradient-gradient matrix K[a][b] = sum_q s G_a.G_b has zero row sums, so
// its last row / column follow from the others.  Fills K[a][NNE-1] for a < NNE-1 and K[NNE-1][NNE-1] from the upper
// triangle of the leading (NNE-1) x (NNE-1) block.
template <int NNE>
PFG_DEV void close_rows_upper(double (&K)[NNE][NNE]) {
    constexpr int L = NNE - 1;
    double last = 0.0;
#pragma unroll
    for (int a = 0; a < L; ++a) {
        double sum = 0.0;
#pragma unroll
        for (int b = 0; b < L; ++b) sum += (a <= b) ? K[a][b] : K[b][a];
        K[a][L] = -sum;
        last += sum;
    }
    K[L][L] = last;  // = -sum_a K[a][L]
}

// emit a symmetric scalar matrix held as upper-triangular accumulators
template <int NNE, class Sink>
PFG_DEV void emit_sym_scalar(Sink& sink, int mat, const double (&K)[NNE][NNE]) {
#pragma unroll
    for (int a = 0; a < NNE; ++a)
#pragma unroll
        for (int b = 0; b < NNE; ++b) {
            const double v = (a <= b) ? K[a][b] : K[b][a];
            sink.block(mat, a, b, &v);
        }
}

// ---------------------------------------------------------------------------------------------
// Physics operators, one thread per element.  Op::run(mv, prm, nodes, elem, sink)
// ---------------------------------------------------------------------------------------------
template <int NNE_>
struct PoissonOp {  // LinearPoisson._compute_element_jacobian (pyfem.py:1188-1217, einsum :1177-1185)
    static constexpr int NNE = NNE_, M = 1, NMAT = 1, NVEC = 0;
    static constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
    static constexpr bool NEEDS_ELEM = false, SYM = true;
    struct Params {
        Material mat;
    };
    __host__ __device__ __forceinline__ static const double* field(const Params& prm) { return prm.mat.rho; }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&xe)[NNE][DIM], const double (&fe)[NNE], int64_t,
                            Sink& sink) {
        double cq[NQ];
        material_at_quads<NNE>(prm.mat, fe, cq);
        integrate(xe, cq, sink);
    }
    template <class Sink>  // complex nodal density: one part of the complex RAMP factor per pass
    PFG_DEV static void run_complex(const Params& prm, const double (&xe)[NNE][DIM], const double (&fe)[NNE],
                                    const double (&fi)[NNE], Sink& sink) {
        double cq[NQ];
        material_at_quads_complex<NNE>(prm.mat, fe, fi, cq);
        integrate(xe, cq, sink);
    }
    template <class Sink>
    PFG_DEV static void integrate(const double (&xe)[NNE][DIM], const double (&cq)[NQ], Sink& sink) {
        double K[NNE][NNE];
#pragma unroll
        for (int a = 0; a < NNE; ++a)
#pragma unroll
            for (int b = 0; b < NNE; ++b) K[a][b] = 0.0;
        GeoCtx<NNE> geo(xe);
        for_each_q<NQ>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            double det, G[NNE][DIM];
            geo.template at<Q>(xe, det, G);
            const double s = cq[Q] * fast_rcp(det);  // w = 1 (pyfem.py:92,110)
#pragma unroll
            for (int a = 0; a < NNE - 1; ++a) {  // the last node's row / column: close_rows_upper
                double H[DIM];
#pragma unroll
                for (int l = 0; l < DIM; ++l) H[l] = s * G[a][l];
#pragma unroll
                for (int b = a; b < NNE - 1; ++b)
#pragma unroll
                    for (int l = 0; l < DIM; ++l) K[a][b] = fma(H[l], G[b][l], K[a][b]);
            }
        });
        close_rows_upper<NNE>(K);
        emit_sym_scalar<NNE>(sink, 0, K);
    }
};

template <int NNE_>
struct HelmholtzOp {  // Helmholtz._compute_element_jacobian_and_rhs (pyfem.py:2138-2177)
    static constexpr int NNE = NNE_, M = 1, NMAT = 2, NVEC = 0;  // matrix 0 = K, 1 = R
    static constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
    static constexpr bool NEEDS_ELEM = false, SYM = true;
    static constexpr bool DEEP_X = (NNE_ == 4);  // quad4 Helmholtz: 1.079 -> 1.066 ms with the window data a chunk ahead
    struct Params {
        double r0sq;
    };
    __host__ __device__ __forceinline__ static const double* field(const Params&) { return nullptr; }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&xe)[NNE][DIM], const double (&)[NNE], int64_t,
                            Sink& sink) {
        double K[NNE][NNE], R[NNE][NNE];
#pragma unroll
        for (int a = 0; a < NNE; ++a)
#pragma unroll
            for (int b = 0; b < NNE; ++b) K[a][b] = R[a][b] = 0.0;
        GeoCtx<NNE> geo(xe);
        for_each_q<NQ>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            double det, G[NNE][DIM];
            geo.template at<Q>(xe, det, G);
            const double s = prm.r0sq * fast_rcp(det);
#pragma unroll
            for (int a = 0; a < NNE; ++a) {
                double H[DIM];
#pragma unroll
                for (int l = 0; l < DIM; ++l) H[l] = s * G[a][l];
#pragma unroll
                for (int b = a; b < NNE; ++b) {
                    if (b < NNE - 1) {  // stiffness part: the last node's row / column follows from the zero row sums
#pragma unroll
                        for (int l = 0; l < DIM; ++l) K[a][b] = fma(H[l], G[b][l], K[a][b]);
                    }
                    R[a][b] = fma(det, Elem<NNE>::N(Q, a) * Elem<NNE>::N(Q, b), R[a][b]);
                }
            }
        });
        close_rows_upper<NNE>(K);
#pragma unroll
        for (int a = 0; a < NNE; ++a)
#pragma unroll
            for (int b = a; b < NNE; ++b) K[a][b] += R[a][b];  // Ke += Re (pyfem.py:2176)
        emit_sym_scalar<NNE>(sink, 0, K);
        emit_sym_scalar<NNE>(sink, 1, R);
    }
};

struct ElasticityQuad4Op {  // plane stress, LinearElasticity._compute_element_jacobian (pyfem.py:2029-2068)
    static constexpr int NNE = 4, M = 2, NMAT = 1, NVEC = 0, DIM = 2, NQ = 4;
    static constexpr bool NEEDS_ELEM = false, SYM = true;
    static constexpr bool DEEP_X = true;  // may gather the window data a whole chunk ahead (pfg_assemble.cu, k_tile XD)
    struct Params {
        Material mat;
        double c11, c12, c33;  // C0 entries (pyfem.py:1746-1750)
    };
    __host__ __device__ __forceinline__ static const double* field(const Params& prm) { return prm.mat.rho; }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&xe)[4][2], const double (&fe)[4], int64_t, Sink& sink) {
        double cq[4];
        material_at_quads<4>(prm.mat, fe, cq);
        integrate(prm, xe, cq, sink);
    }
    template <class Sink>  // complex nodal density: one part of the complex RAMP factor per pass
    PFG_DEV static void run_complex(const Params& prm, const double (&xe)[4][2], const double (&fe)[4],
                                    const double (&fi)[4], Sink& sink) {
        double cq[4];
        material_at_quads_complex<4>(prm.mat, fe, fi, cq);
        integrate(prm, xe, cq, sink);
    }
    template <class Sink>
    PFG_DEV static void integrate(const Params& prm, const double (&xe)[4][2], const double (&cq)[4], Sink& sink) {
        // Partition of unity: sum_a grad N_a = 0 at every point, so every row of the element matrix sums to zero.  Only
        // the node pairs among the first three nodes are integrated -- XX, YY (symmetric, 6 each) and XY (all 9;
        // YX[a][b] = XY[b][a]) -- and the blocks of the fourth node follow from the row / column sums: 21 instead of
        // 36 running sums per quadrature point, and no G / h for the fourth node.
        double XX[4][4], YY[4][4], XY[4][4];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) XX[a][b] = YY[a][b] = XY[a][b] = 0.0;
        const Quad4Coef c = quad4_coef(xe);
        for_each_q<4>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            double det, G[4][2];
            quad4_geo<Q>(c, det, G);
            const double s = cq[Q] * fast_rcp(det);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double hx = s * G[a][0], hy = s * G[a][1];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    XY[a][b] = fma(hx, G[b][1], XY[a][b]);
                    if (b >= a) {
                        XX[a][b] = fma(hx, G[b][0], XX[a][b]);
                        YY[a][b] = fma(hy, G[b][1], YY[a][b]);
                    }
                }
            }
        });
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < a; ++b) XX[a][b] = XX[b][a], YY[a][b] = YY[b][a];
#pragma unroll
        for (int a = 0; a < 3; ++a) {  // fourth node: minus the sums over the other three
            XX[a][3] = -((XX[a][0] + XX[a][1]) + XX[a][2]);
            YY[a][3] = -((YY[a][0] + YY[a][1]) + YY[a][2]);
            XY[a][3] = -((XY[a][0] + XY[a][1]) + XY[a][2]);
            XY[3][a] = -((XY[0][a] + XY[1][a]) + XY[2][a]);
        }
        XX[3][3] = -((XX[0][3] + XX[1][3]) + XX[2][3]);
        YY[3][3] = -((YY[0][3] + YY[1][3]) + YY[2][3]);
        XY[3][3] = -((XY[3][0] + XY[3][1]) + XY[3][2]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) {
                const double yx = XY[b][a];  // sum_q s Gy_a Gx_b
                double blk[4];  // [alpha][beta] for (row node a, col node b)
                blk[0] = fma(prm.c11, XX[a][b], prm.c33 * YY[a][b]);
                blk[1] = fma(prm.c12, XY[a][b], prm.c33 * yx);
                blk[2] = fma(prm.c12, yx, prm.c33 * XY[a][b]);
                blk[3] = fma(prm.c11, YY[a][b], prm.c33 * XX[a][b]);
                sink.block(0, a, b, blk);
                if (b != a) {
                    const double t[4] = {blk[0], blk[2], blk[1], blk[3]};
                    sink.block(0, b, a, t);
                }
            }
    }
};

constexpr int kMaxXdv = 32;

struct NlPoissonQuad4Op {  // NonlinearPoisson2D: Jacobian (pyfem.py:1541-1610) + residual (pyfem.py:1474-1539)
    static constexpr int NNE = 4, M = 1, NMAT = 1, NVEC = 1, DIM = 2, NQ = 4;
    static constexpr bool NEEDS_ELEM = false, SYM = false;  // the Newton Jacobian is not symmetric
    static constexpr bool DEEP_PREFETCH = true;             // register-limited (3 CTAs / SM): shared memory to spare
    struct Params {
        const double* u;
        int nxdv;
        double coef[kMaxXdv];  // xdv[k] * binom(nxdv-1, k)  (pyfem.py:1466-1470)
    };
    PFG_DEV static double gfun(double x, double y) {  // pyfem.py:1438-1446
        return 1e4 * x * (1.0 - x) * (1.0 - 2.0 * x) * y * (1.0 - y) * (1.0 - 2.0 * y);
    }
    __host__ __device__ __forceinline__ static const double* field(const Params& prm) { return prm.u; }
    // h = 1 + 4y(1-y) sum_k coef_k (1-x)^(n-1-k) x^k (pyfem.py:1450-1472).  Two-variable Horner: S_k = (1-x) S_{k-1} +
    // coef_k x^k needs no power table (three flops per term, coefficients straight from the constant bank).  The four
    // quadrature points share ONE coefficient loop: four independent Horner chains instead of four
    // dependent ones run one after the other (the loop bound is a run-time value, so the compiler cannot interleave
    // the per-point loops itself).
    PFG_DEV static void hfun4(const Params& prm, const double (&x)[4], const double (&y)[4], double (&h)[4]) {
        const int n = prm.nxdv;
        double om[4], s[4], xp[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) om[q] = 1.0 - x[q], s[q] = prm.coef[0], xp[q] = 1.0;
        for (int k = 1; k < n; ++k) {
            const double ck = prm.coef[k];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                xp[q] *= x[q];
                s[q] = fma(s[q], om[q], ck * xp[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = fma(s[q], 4.0 * y[q] * (1.0 - y[q]), 1.0);
    }
    static constexpr int qidx(int sx, int sy) {  // node / quadrature point with local coordinate signs (sx, sy), 1 = plus
        return sy ? (sx ? 2 : 3) : (sx ? 1 : 0);
    }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&xe)[4][2], const double (&ue)[4], int64_t, Sink& sink) {
        // K = S + T.  S[a][b] = sum_q c1 G_a.G_b is symmetric with zero row sums (partition of unity): only the six
        // node pairs among the first three nodes are integrated, the fourth node follows from close_rows_upper, and
        // the gradient part of the residual is S u.  T[a][b] = sum_q c2 (G_a.grad u) N_b, the non-symmetric Newton
        // term, is linear in eight numbers per element: G_a.grad u = dN_a/dxi P + dN_a/deta R with
        // P = y_eta gux - x_eta guy, R = x_xi guy - y_xi gux, and both basis tables are tensor products of the
        // one-dimensional values (1 +- g), so T is formed after the quadrature loop from (P_q, R_q) by two
        // one-dimensional contractions each (56 flops per element instead of 28 per point).
        double S[4][4], res[4], Pq[4], Rq[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            res[a] = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) S[a][b] = 0.0;
        }
        double xq[4], yq[4], hq[4];
        for_each_q<4>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            double x = 0.0, y = 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                x = fma(Elem<4>::N(Q, a), xe[a][0], x);
                y = fma(Elem<4>::N(Q, a), xe[a][1], y);
            }
            xq[Q] = x, yq[Q] = y;
        });
        hfun4(prm, xq, yq, hq);
        const Quad4Coef c = quad4_coef(xe);
        const Quad4Field fu = quad4_field4(ue[0], ue[1], ue[2], ue[3]);  // 4 u = m + a xi + b eta + c xi eta
        for_each_q<4>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            constexpr double xi = Elem<4>::qp(Q, 0), eta = Elem<4>::qp(Q, 1);
            const double xxi = fma(c.cx, eta, c.ax), xeta = fma(c.cx, xi, c.bx);
            const double yxi = fma(c.cy, eta, c.ay), yeta = fma(c.cy, xi, c.by);
            const double det = xxi * yeta - xeta * yxi;
            const double uq = fma(fu.c, 0.25 * xi * eta, fma(fu.b, 0.25 * eta, fma(fu.a, 0.25 * xi, 0.25 * fu.m)));
            const double uxi = fma(fu.c, eta, fu.a), ueta = fma(fu.c, xi, fu.b);  // 4 du/dxi, 4 du/deta
            const double P4 = yeta * uxi - yxi * ueta;   // 4 det du/dx
            const double R4 = xxi * ueta - xeta * uxi;   // 4 det du/dy
            const double h = hq[Q];
            const double inv = fast_rcp(det);
            const double c1 = h * fma(uq, uq, 1.0) * inv;  // detJ h (1+u^2) w / det^2
            const double c2 = 0.03125 * h * uq * inv;      // 2 h u / det, with 1/4 of the field gradient and 1/16 of T's tables
            const double gsrc = det * gfun(xq[Q], yq[Q]);
            Pq[Q] = c2 * (yeta * P4 - xeta * R4);
            Rq[Q] = c2 * (xxi * R4 - yxi * P4);
#pragma unroll
            for (int a = 0; a < 4; ++a) res[a] = fma(-gsrc, Elem<4>::N(Q, a), res[a]);
            double G[3][2];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double dxi = Elem<4>::dN(Q, a, 0), deta = Elem<4>::dN(Q, a, 1);
                G[a][0] = dxi * yeta - deta * yxi;
                G[a][1] = deta * xxi - dxi * xeta;
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double hx = c1 * G[a][0], hy = c1 * G[a][1];
#pragma unroll
                for (int b = a; b < 3; ++b) {
                    S[a][b] = fma(hx, G[b][0], S[a][b]);
                    S[a][b] = fma(hy, G[b][1], S[a][b]);
                }
            }
        });
        close_rows_upper<4>(S);
        // T[a][b] = sx_a B[sx_b][sy_a][sy_b] + sy_a B'[sy_b][sx_a][sx_b]:
        //   dN_a/dxi (q) = sx_a f(sy_a sy_q) / 4, dN_a/deta (q) = sy_a f(sx_a sx_q) / 4, N_b(q) = f(sx_b sx_q) f(sy_b sy_q) / 4,
        //   f(+1) = 1 + g, f(-1) = 1 - g (signs as indices: 0 = minus, 1 = plus)
        constexpr double fp = 1.0 + PFG_G, fm = 1.0 - PFG_G;
        double Bs[2][2], Bm[2], Cs[2][2], Cm[2];  // B / B' for equal (by sign) and for unequal second indices
#pragma unroll
        for (int sb = 0; sb < 2; ++sb) {
            const double a0 = fma(fp, Pq[qidx(sb, 0)], fm * Pq[qidx(1 - sb, 0)]);  // sum over sx_q, sy_q = minus
            const double a1 = fma(fp, Pq[qidx(sb, 1)], fm * Pq[qidx(1 - sb, 1)]);  // sy_q = plus
            Bs[sb][0] = fma(fp * fp, a0, (fm * fm) * a1);
            Bs[sb][1] = fma(fm * fm, a0, (fp * fp) * a1);
            Bm[sb] = (fp * fm) * (a0 + a1);
            const double d0 = fma(fp, Rq[qidx(0, sb)], fm * Rq[qidx(0, 1 - sb)]);  // sum over sy_q, sx_q = minus
            const double d1 = fma(fp, Rq[qidx(1, sb)], fm * Rq[qidx(1, 1 - sb)]);  // sx_q = plus
            Cs[sb][0] = fma(fp * fp, d0, (fm * fm) * d1);
            Cs[sb][1] = fma(fm * fm, d0, (fp * fp) * d1);
            Cm[sb] = (fp * fm) * (d0 + d1);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            constexpr int sxs[4] = {0, 1, 1, 0}, sys[4] = {0, 0, 1, 1};
            double r = res[a];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double pb = (sys[a] == sys[b]) ? Bs[sxs[b]][sys[a]] : Bm[sxs[b]];
                const double rb = (sxs[a] == sxs[b]) ? Cs[sys[b]][sxs[a]] : Cm[sys[b]];
                const double tab = (sxs[a] ? pb : -pb) + (sys[a] ? rb : -rb);
                const double sab = (a <= b) ? S[a][b] : S[b][a];
                r = fma(sab, ue[b], r);  // gradient part of the residual: S u
                const double v = tab + sab;
                sink.block(0, a, b, &v);
            }
            sink.vec(a, r);
        }
    }
};

template <int NNE_>
struct PoissonRhsOp {  // LinearPoisson._compute_element_rhs (pyfem.py:1137-1173, einsum :1132-1134)
    static constexpr int NNE = NNE_, M = 1, NMAT = 0, NVEC = 1;
    static constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
    static constexpr bool NEEDS_ELEM = true, SYM = false;
    struct Params {
        const double* gq;  // (nelems, NQ) source term at the quadrature points
    };
    __host__ __device__ __forceinline__ static const double* field(const Params&) { return nullptr; }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&xe)[NNE][DIM], const double (&)[NNE], int64_t elem,
                            Sink& sink) {
        double f[NNE];
#pragma unroll
        for (int a = 0; a < NNE; ++a) f[a] = 0.0;
        GeoCtx<NNE> geo(xe);
        const double* g = prm.gq + elem * NQ;
        for_each_q<NQ>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            double det, G[NNE][DIM];
            geo.template at<Q>(xe, det, G);
            const double wg = det * __ldg(g + Q);
            // The reference's einsum "ik,k,jk,ik->ij" (pyfem.py:1132-1134) indexes the (nquads, nnodes) table
            // N as N[j,k] with j the OUTPUT (node) index and k the quadrature index, i.e. it uses
            // N(quad=j, node=k).  For quad4 the table is symmetric; for hex8 (quadrature order differs from
            // node order) it is not, and parity means reproducing the reference, so the indices are swapped.
#pragma unroll
            for (int a = 0; a < NNE; ++a) f[a] = fma(wg, Elem<NNE>::N(a, Q), f[a]);
        });
#pragma unroll
        for (int a = 0; a < NNE; ++a) sink.vec(a, f[a]);
    }
};

// ---------------------------------------------------------------------------------------------
// Caller-supplied element matrices / vectors: the scatter of ModelBase._assemble_jacobian (pyfem.py:920-931) and
// ModelBase._assemble_rhs (pyfem.py:860-875) on their own -- the slot the reference's A2DWrapper uses
// (pyfem.py:2255-2277: a native plugin fills the element Jacobians, pyfem scatters them).
// ---------------------------------------------------------------------------------------------
template <int NNE_, int M_>
struct ScatterMatOp {
    static constexpr int NNE = NNE_, M = M_, NMAT = 1, NVEC = 0, D = NNE_ * M_;
    static constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
    static constexpr bool NEEDS_ELEM = true, SYM = false;  // nothing is assumed about the supplied matrices
    struct Params {
        const double* Ke;  // (nelems, D, D) row-major
    };
    __host__ __device__ __forceinline__ static const double* field(const Params&) { return nullptr; }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&)[NNE][DIM], const double (&)[NNE], int64_t elem,
                            Sink& sink) {
        const double* __restrict__ K = prm.Ke + elem * (int64_t)(D * D);
#pragma unroll
        for (int a = 0; a < NNE; ++a)
#pragma unroll
            for (int b = 0; b < NNE; ++b) {
                double blk[M * M];
#pragma unroll
                for (int al = 0; al < M; ++al)
#pragma unroll
                    for (int be = 0; be < M; ++be) blk[al * M + be] = __ldg(K + (a * M + al) * D + b * M + be);
                sink.block(0, a, b, blk);
            }
    }
};

template <int NNE_>
struct ScatterVecOp {  // rhs[conn[e, a]] += fe[e, a]
    static constexpr int NNE = NNE_, M = 1, NMAT = 0, NVEC = 1;
    static constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
    static constexpr bool NEEDS_ELEM = true, SYM = false;
    struct Params {
        const double* fe;  // (nelems, NNE)
    };
    __host__ __device__ __forceinline__ static const double* field(const Params&) { return nullptr; }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&)[NNE][DIM], const double (&)[NNE], int64_t elem,
                            Sink& sink) {
#pragma unroll
        for (int a = 0; a < NNE; ++a) sink.vec(a, __ldg(prm.fe + elem * NNE + a));
    }
};

// Element matrices / vectors to global memory, (nelems, D, D) and (nelems, NNE): the reference's
// _compute_element_jacobian / _compute_element_rhs outputs (Ke_mat, rhs_e) for callers that want them.
template <class Op>
struct GlobalSink {
    static constexpr int NNE = Op::NNE, M = Op::M, D = NNE * M;
    double* Ke[2];  // per matrix (may be null)
    double* fe;     // (may be null)
    int64_t e;
    PFG_DEV void block(int mat, int a, int b, const double* blk) const {
        if (Ke[mat] == nullptr) return;
        double* dst = Ke[mat] + e * (int64_t)(D * D) + (a * M) * D + b * M;
#pragma unroll
        for (int al = 0; al < M; ++al)
#pragma unroll
            for (int be = 0; be < M; ++be) dst[al * D + be] = blk[al * M + be];
    }
    PFG_DEV void vec(int a, double v) const {
        if (fe != nullptr) fe[e * NNE + a] = v;
    }
};

// ---------------------------------------------------------------------------------------------
// hex8 3-D elasticity: 8 threads per element.  Thread t first acts as quadrature point t (geometry
// into shared staging), then as local row node t (3 x 24 row block from 72 accumulators).
// LinearElasticity._compute_element_jacobian, 3-D branch (pyfem.py:2000-2011, 2017-2026, 1752-1757).
// ---------------------------------------------------------------------------------------------
struct ElasticityHex8Params {
    Material mat;
    double c11, c12, c44;
};

constexpr int kHexStageDoubles = 8 * 25 + 2;  // per element: [q][8 nodes x 3 + s_q], padded against bank conflicts

// `stage` points at this element's staging area; `lane8` is the thread's index inside its octet.
// All 8 threads of an octet must call this together (uses __syncwarp on the octet's lanes).
template <class Sink>
PFG_DEV void elasticity_hex8_octet(const MeshView& mv, const ElasticityHex8Params& prm, const int (&nodes)[8],
                                   double* __restrict__ stage, int lane8, unsigned octet_mask, bool row_wanted,
                                   Sink& sink) {
    // ---- role 1: quadrature point `lane8` (x slowest, z fastest; basis evaluated at run time so the
    //      eight lanes of an octet execute one instruction stream)
    double xe[8][3];
    load_coords<8>(mv.X, nodes, xe);
    const double qx = (lane8 & 4) ? PFG_G : -PFG_G, qy = (lane8 & 2) ? PFG_G : -PFG_G, qz = (lane8 & 1) ? PFG_G : -PFG_G;
    const double fx[2] = {1.0 - qx, 1.0 + qx}, fy[2] = {1.0 - qy, 1.0 + qy}, fz[2] = {1.0 - qz, 1.0 + qz};
    double det, G[8][3], cq;
    {
        double dn[8][3], shape[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            constexpr double e8 = 0.125;
            const int ix = Elem<8>::sgn(a, 0) > 0 ? 1 : 0, iy = Elem<8>::sgn(a, 1) > 0 ? 1 : 0, iz = Elem<8>::sgn(a, 2) > 0 ? 1 : 0;
            dn[a][0] = (Elem<8>::sgn(a, 0) * e8) * (fy[iy] * fz[iz]);
            dn[a][1] = (Elem<8>::sgn(a, 1) * e8) * (fx[ix] * fz[iz]);
            dn[a][2] = (Elem<8>::sgn(a, 2) * e8) * (fx[ix] * fy[iy]);
            shape[a] = e8 * fx[ix] * (fy[iy] * fz[iz]);
        }
        double J[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < 8; ++a) s = fma(dn[a][k], xe[a][j], s);
                J[j][k] = s;
            }
        double A[3][3];
        A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        A[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
        A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
        A[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]);
        A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
        A[1][2] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]);
        A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        A[2][1] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]);
        A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int l = 0; l < 3; ++l) G[a][l] = dn[a][0] * A[0][l] + dn[a][1] * A[1][l] + dn[a][2] * A[2][l];
        if (prm.mat.rho == nullptr) {
            cq = prm.mat.c_const;
        } else {
            double rq = 0.0;
#pragma unroll
            for (int a = 0; a < 8; ++a) rq = fma(shape[a], __ldg(prm.mat.rho + nodes[a]), rq);
            if (prm.mat.rho_im == nullptr) {
                cq = rq * fast_rcp(fma(prm.mat.p, 1.0 - rq, 1.0));
            } else {  // complex-step verification: one part of the complex RAMP factor per pass
                double iq = 0.0;
#pragma unroll
                for (int a = 0; a < 8; ++a) iq = fma(shape[a], __ldg(prm.mat.rho_im + nodes[a]), iq);
                cq = ramp_complex_part(rq, iq, prm.mat.p, prm.mat.part);
            }
        }
    }
    double* mine = stage + lane8 * 25;
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int l = 0; l < 3; ++l) mine[b * 3 + l] = G[b][l];
    mine[24] = cq * fast_rcp(det);
    __syncwarp(octet_mask);
    // ---- role 2: row node `lane8`
    if (row_wanted) {
        double P[8][3][3];
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) P[b][i][j] = 0.0;
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const double* gq = stage + q * 25;
            const double s = gq[24];
            const double h0 = s * gq[lane8 * 3 + 0], h1 = s * gq[lane8 * 3 + 1], h2 = s * gq[lane8 * 3 + 2];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const double g0 = gq[b * 3 + 0], g1 = gq[b * 3 + 1], g2 = gq[b * 3 + 2];
                P[b][0][0] = fma(h0, g0, P[b][0][0]);
                P[b][0][1] = fma(h0, g1, P[b][0][1]);
                P[b][0][2] = fma(h0, g2, P[b][0][2]);
                P[b][1][0] = fma(h1, g0, P[b][1][0]);
                P[b][1][1] = fma(h1, g1, P[b][1][1]);
                P[b][1][2] = fma(h1, g2, P[b][1][2]);
                P[b][2][0] = fma(h2, g0, P[b][2][0]);
                P[b][2][1] = fma(h2, g1, P[b][2][1]);
                P[b][2][2] = fma(h2, g2, P[b][2][2]);
            }
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double blk[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (i == j) {
                        const int l1 = (i + 1) % 3, l2 = (i + 2) % 3;
                        blk[i * 3 + j] = fma(prm.c11, P[b][i][i], prm.c44 * (P[b][l1][l1] + P[b][l2][l2]));
                    } else {
                        blk[i * 3 + j] = fma(prm.c12, P[b][i][j], prm.c44 * P[b][j][i]);
                    }
                }
            sink.block(0, lane8, b, blk);
        }
    }
    __syncwarp(octet_mask);  // staging may be reused by the caller's next element
}

// ---------------------------------------------------------------------------------------------
// hex8 3-D elasticity, owner-computes form (no atomics): a geometry pass stores the adjugate of J and the
// quadrature weight s_q = c_q / det(J_q) per (element, quadrature point) -- 10 doubles -- and a row pass gives
// every (node, incident element) pair one lane that forms the node's 3 x 24 row block of that element and adds
// it into a shared-memory image of the node's CSR rows (pfg_assemble.cu: k_hex8_geometry, k_hex8_chunk_rows).
// Same math as elasticity_hex8_octet above.
// ---------------------------------------------------------------------------------------------
constexpr int kHexGeoDoubles = 10;  // per (element, quadrature point): A[3][3] = det(J) inv(J), then s_q

// lane = (element, quadrature point `q`): nodes are the element's eight corners
PFG_DEV void hex8_geometry_point(const MeshView& mv, const Material& mat, const int (&nodes)[8], int q, bool skip,
                                 double* __restrict__ out) {
    double xe[8][3];
    load_coords<8>(mv.X, nodes, xe);
    const double qx = (q & 4) ? PFG_G : -PFG_G, qy = (q & 2) ? PFG_G : -PFG_G, qz = (q & 1) ? PFG_G : -PFG_G;
    const double fx[2] = {1.0 - qx, 1.0 + qx}, fy[2] = {1.0 - qy, 1.0 + qy}, fz[2] = {1.0 - qz, 1.0 + qz};
    double J[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) J[j][k] = 0.0;
    double rq = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        constexpr double e8 = 0.125;
        const int ix = Elem<8>::sgn(a, 0) > 0 ? 1 : 0, iy = Elem<8>::sgn(a, 1) > 0 ? 1 : 0, iz = Elem<8>::sgn(a, 2) > 0 ? 1 : 0;
        const double dn0 = (Elem<8>::sgn(a, 0) * e8) * (fy[iy] * fz[iz]);
        const double dn1 = (Elem<8>::sgn(a, 1) * e8) * (fx[ix] * fz[iz]);
        const double dn2 = (Elem<8>::sgn(a, 2) * e8) * (fx[ix] * fy[iy]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            J[j][0] = fma(dn0, xe[a][j], J[j][0]);
            J[j][1] = fma(dn1, xe[a][j], J[j][1]);
            J[j][2] = fma(dn2, xe[a][j], J[j][2]);
        }
        if (mat.rho != nullptr) rq = fma(e8 * fx[ix] * (fy[iy] * fz[iz]), __ldg(mat.rho + nodes[a]), rq);
    }
    double A[3][3];
    A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    A[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
    A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    A[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]);
    A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    A[1][2] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]);
    A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    A[2][1] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]);
    A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
    double cq;
    if (mat.rho == nullptr) cq = mat.c_const;
    else cq = rq * fast_rcp(fma(mat.p, 1.0 - rq, 1.0));
    const double s = skip ? 0.0 : cq * fast_rcp(det);  // a masked element (another rank integrates it) adds zeros
    double2* o = reinterpret_cast<double2*>(out);
    o[0] = make_double2(A[0][0], A[0][1]);
    o[1] = make_double2(A[0][2], A[1][0]);
    o[2] = make_double2(A[1][1], A[1][2]);
    o[3] = make_double2(A[2][0], A[2][1]);
    o[4] = make_double2(A[2][2], s);
}

// lane = (row node, incident element): the row node's local coordinate signs (/ 8) are run-time values, the NB
// column nodes B0 .. B0+NB-1 are compile time.  P[b][i][j] = sum_q s_q G_a,i G_b,j.
// `geo_e` points at the element's 8 x 10 doubles in shared memory.
template <int B0, int NB>
PFG_DEV void hex8_row_products(const double* __restrict__ geo_e, double sx8, double sy8, double sz8,
                               double (&P)[NB][3][3]) {
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) P[b][i][j] = 0.0;
    for_each_q<8>([&](auto qc) {
        constexpr int Q = decltype(qc)::value;
        const double2* g = reinterpret_cast<const double2*>(geo_e + Q * kHexGeoDoubles);
        const double2 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], g4 = g[4];
        const double A[3][3] = {{g0.x, g0.y, g1.x}, {g1.y, g2.x, g2.y}, {g3.x, g3.y, g4.x}};
        const double s = g4.y;
        // row node: basis derivatives at run time (s?8 = local coordinate sign / 8: 1 + sign * q = fma(s?8, 8 q, 1))
        const double ax = fma(sx8, 8.0 * Elem<8>::qp(Q, 0), 1.0), ay = fma(sy8, 8.0 * Elem<8>::qp(Q, 1), 1.0),
                     az = fma(sz8, 8.0 * Elem<8>::qp(Q, 2), 1.0);
        const double d0 = sx8 * (ay * az), d1 = sy8 * (ax * az), d2 = sz8 * (ax * ay);
        double h[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) h[l] = s * (d0 * A[0][l] + d1 * A[1][l] + d2 * A[2][l]);
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
            double gb[3];
#pragma unroll
            for (int l = 0; l < 3; ++l)
                gb[l] = Elem<8>::dN(Q, B0 + bb, 0) * A[0][l] + Elem<8>::dN(Q, B0 + bb, 1) * A[1][l] +
                        Elem<8>::dN(Q, B0 + bb, 2) * A[2][l];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) P[bb][i][j] = fma(h[i], gb[j], P[bb][i][j]);
        }
    });
}

// The same row block through the modes of the trilinear basis.  8 grad N_b = sum over the seven non-constant
// monomials sigma_k(b) of the node's corner signs (x, y, z, xy, xz, yz, xyz) of M_k, with
//   M_x = A[0], M_y = A[1], M_z = A[2], M_xy = eta A[0] + xi A[1], M_xz = zeta A[0] + xi A[2],
//   M_yz = zeta A[1] + eta A[2], M_xyz = eta zeta A[0] + xi zeta A[1] + xi eta A[2]     (rows of the adjugate),
// and (xi, eta, zeta) = (+-g, +-g, +-g) at a quadrature point: the modes cost 15 additions per point (the factors g, g^2
// leave the sum over the points) where the seven column gradients cost 63 FMAs.  V[k][i][j] = sum_q s_q G_a,i M'_k,j;
// the node blocks are P_ab = (1/8) sum_k sigma_k(b) scale_k V_k -- a three-stage butterfly per entry, which also yields
// the eighth block that hex8_rows_closed takes from the zero row sums (there is no constant mode).
PFG_DEV void hex8_row_modes(const double* __restrict__ geo_e, double sx8, double sy8, double sz8, double (&V)[7][3][3]) {
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) V[k][i][j] = 0.0;
    for_each_q<8>([&](auto qc) {
        constexpr int Q = decltype(qc)::value;
        const double2* g = reinterpret_cast<const double2*>(geo_e + Q * kHexGeoDoubles);
        const double2 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], g4 = g[4];
        const double A[3][3] = {{g0.x, g0.y, g1.x}, {g1.y, g2.x, g2.y}, {g3.x, g3.y, g4.x}};
        const double s = g4.y;
        const double ax = fma(sx8, 8.0 * Elem<8>::qp(Q, 0), 1.0), ay = fma(sy8, 8.0 * Elem<8>::qp(Q, 1), 1.0),
                     az = fma(sz8, 8.0 * Elem<8>::qp(Q, 2), 1.0);
        const double d0 = sx8 * (ay * az), d1 = sy8 * (ax * az), d2 = sz8 * (ax * ay);
        double h[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) h[l] = s * (d0 * A[0][l] + d1 * A[1][l] + d2 * A[2][l]);
        constexpr bool px = Elem<8>::qp(Q, 0) > 0, py = Elem<8>::qp(Q, 1) > 0, pz = Elem<8>::qp(Q, 2) > 0;  // signs of the point
        double M[7][3];
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const double a0 = A[0][l], a1 = A[1][l], a2 = A[2][l];
            M[0][l] = a0, M[1][l] = a1, M[2][l] = a2;
            M[3][l] = (py ? a0 : -a0) + (px ? a1 : -a1);  // xy / g
            M[4][l] = (pz ? a0 : -a0) + (px ? a2 : -a2);  // xz / g
            M[5][l] = (pz ? a1 : -a1) + (py ? a2 : -a2);  // yz / g
            M[6][l] = ((py == pz) ? a0 : -a0) + ((px == pz) ? a1 : -a1) + ((px == py) ? a2 : -a2);  // xyz / g^2
        }
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) V[k][i][j] = fma(h[i], M[k][j], V[k][i][j]);
    });
}

// modes -> the eight node blocks, entry by entry: 8 P_b = sx Vx + sy Vy + sz Vz + sx sy g Vxy + sx sz g Vxz + sy sz g Vyz +
// sx sy sz g^2 Vxyz with (sx, sy, sz) the corner signs of node b (the factor 1/8 is left to the caller's constants)
PFG_DEV void hex8_modes_to_nodes(const double (&V)[7][3][3], double (&P)[8][3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double vx = V[0][i][j], vy = V[1][i][j], vz = V[2][i][j];
            const double vxy = PFG_G * V[3][i][j], vxz = PFG_G * V[4][i][j], vyz = PFG_G * V[5][i][j];
            const double vxyz = (PFG_G * PFG_G) * V[6][i][j];
            // along z: index [sz] with 0 = minus, 1 = plus
            const double a10[2] = {vx - vxz, vx + vxz}, a01[2] = {vy - vyz, vy + vyz}, a11[2] = {vxy - vxyz, vxy + vxyz};
            const double a00[2] = {-vz, vz};
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int sx = Elem<8>::sgn(b, 0) > 0 ? 1 : 0, sy = Elem<8>::sgn(b, 1) > 0 ? 1 : 0,
                          sz = Elem<8>::sgn(b, 2) > 0 ? 1 : 0;
                const double b0 = sy ? a00[sz] + a01[sz] : a00[sz] - a01[sz];
                const double b1 = sy ? a10[sz] + a11[sz] : a10[sz] - a11[sz];
                P[b][i][j] = sx ? b0 + b1 : b0 - b1;
            }
        }
}

// C0 applied to one summed product block (pyfem.py:1752-1757, 2017-2026)
PFG_DEV void hex8_apply_c0(const ElasticityHex8Params& prm, const double (&P)[3][3], double (&blk)[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (i == j) {
                const int l1 = (i + 1) % 3, l2 = (i + 2) % 3;
                blk[i * 3 + j] = fma(prm.c11, P[i][i], prm.c44 * (P[l1][l1] + P[l2][l2]));
            } else {
                blk[i * 3 + j] = fma(prm.c12, P[i][j], prm.c44 * P[j][i]);
            }
        }
}

}  // namespace pfg
