// Assembly kernels and their C-ABI entry points.
//
// Two scatter strategies share the element routines of pfg_elem.cuh:
//   atomic  one thread per element, every entry added with red.global.add.f64 at the slot the
//           rank map gives (the reference's coo->csr duplicates-summed scatter, pyfem.py:930-931,
//           and np.add.at for vectors, pyfem.py:872-874);
//   gather  one CTA per row chunk: phase A integrates every element touching the chunk once and
//           stages the row blocks of chunk nodes in shared memory, phase B sums each CSR block's
//           contributions in plan order and writes it exactly once.  No atomics, no zero-fill,
//           bitwise reproducible.
#include <algorithm>
#include <cstdlib>

#include "pfg_elem.cuh"

namespace pfg {

static inline int align16(int x) { return (x + 15) & ~15; }

// ---------------------------------------------------------------------------------------------
// shared-memory row-block stride (doubles) per incidence; odd multiples of the access width keep
// element-per-thread stores and plan-ordered loads spread over the banks
// ---------------------------------------------------------------------------------------------
template <class Op>
struct Layout {
    static constexpr int BLK = Op::M * Op::M;
    static constexpr int RAW = Op::NMAT * Op::NNE * BLK;
    static constexpr int RB = (RAW == 0) ? 0 : ((Op::M == 2) ? RAW + 2 : RAW + 1);
    static constexpr int VEC = Op::NVEC;  // doubles per incidence for the vector output
};

struct Outputs {
    double* vals[2];  // CSR values per matrix (may be null)
    double* vec;      // owned-row vector (may be null)
};

// ---------------------------------------------------------------------------------------------
// sinks
// ---------------------------------------------------------------------------------------------
template <class Op>
struct AtomicSink {
    static constexpr int NNE = Op::NNE, M = Op::M;
    const Outputs& out;
    int64_t base[NNE];  // first value slot of the row node's dof rows, -1 when the row is not owned
    int k[NNE];
    int row[NNE];
    const uint8_t* rank_e;  // rank[(e*NNE + a)*NNE + b]

    PFG_DEV AtomicSink(const MeshView& mv, const Outputs& o, const int (&nodes)[NNE], int64_t e) : out(o) {
        rank_e = mv.rank + e * NNE * NNE;
#pragma unroll
        for (int a = 0; a < NNE; ++a) {
            const int64_t r = nodes[a] - mv.own_begin;
            if (nodes[a] >= mv.own_begin && nodes[a] < mv.own_end) {
                const int64_t p0 = __ldg(mv.blk_ptr + r), p1 = __ldg(mv.blk_ptr + r + 1);
                base[a] = p0 * M * M;
                k[a] = (int)(p1 - p0);
                row[a] = (int)r;
            } else {
                base[a] = -1;
                k[a] = 0;
                row[a] = 0;
            }
        }
    }
    PFG_DEV void block(int mat, int a, int b, const double* blk) const {
        if (base[a] < 0 || out.vals[mat] == nullptr) return;
        const int t = rank_e[a * NNE + b];
        double* dst = out.vals[mat] + base[a] + (int64_t)M * t;
#pragma unroll
        for (int al = 0; al < M; ++al)
#pragma unroll
            for (int be = 0; be < M; ++be) atomicAdd(dst + (int64_t)al * M * k[a] + be, blk[al * M + be]);
    }
    PFG_DEV void vec(int a, double v) const {
        if (base[a] < 0 || out.vec == nullptr) return;
        atomicAdd(out.vec + row[a], v);
    }
};

template <class Op>
struct SmemSink {
    static constexpr int NNE = Op::NNE, M = Op::M, BLK = M * M, RB = Layout<Op>::RB;
    double* rb;    // row-block staging [n_inc][RB]
    double* vecs;  // vector staging [n_inc]
    uint16_t dst[NNE];
    PFG_DEV void block(int mat, int a, int b, const double* blk) const {
        if (dst[a] == kNoDst) return;
        double* p = rb + (int)dst[a] * RB + (mat * NNE + b) * BLK;
        if constexpr (M == 2) {
            reinterpret_cast<double2*>(p)[0] = make_double2(blk[0], blk[1]);
            reinterpret_cast<double2*>(p)[1] = make_double2(blk[2], blk[3]);
        } else {
#pragma unroll
            for (int i = 0; i < BLK; ++i) p[i] = blk[i];
        }
    }
    PFG_DEV void vec(int a, double v) const {
        if (dst[a] == kNoDst) return;
        vecs[dst[a]] = v;
    }
};

// single-row sinks for the 8-threads-per-element hex8 elasticity kernels (row node fixed per thread)
struct HexAtomicRowSink {
    double* dst0;  // values of the row node's first dof row, nullptr when not owned
    int k;
    const uint8_t* rank_row;  // rank[(e*8 + a)*8 + b]
    PFG_DEV void block(int, int, int b, const double* blk) const {
        double* dst = dst0 + 3 * (int)rank_row[b];
#pragma unroll
        for (int al = 0; al < 3; ++al)
#pragma unroll
            for (int be = 0; be < 3; ++be) atomicAdd(dst + (int64_t)al * 3 * k + be, blk[al * 3 + be]);
    }
};

struct HexSmemRowSink {
    double* row;  // this incidence's staged row block [b][3][3]
    PFG_DEV void block(int, int, int b, const double* blk) const {
#pragma unroll
        for (int i = 0; i < 9; ++i) row[b * 9 + i] = blk[i];
    }
};

// ---------------------------------------------------------------------------------------------
// atomic kernels
// ---------------------------------------------------------------------------------------------
template <class Op>
__global__ void __launch_bounds__(128) k_assemble_atomic(MeshView mv, typename Op::Params prm, Outputs out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= mv.nelems) return;
    if (mv.elem_skip != nullptr && mv.elem_skip[e]) return;  // integrated by another rank
    constexpr int NNE = Op::NNE;
    int nodes[NNE];
    if constexpr (NNE == 4) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(mv.conn) + e);
        nodes[0] = v.x, nodes[1] = v.y, nodes[2] = v.z, nodes[3] = v.w;
    } else {
        const int4 v0 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e);
        const int4 v1 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e + 1);
        nodes[0] = v0.x, nodes[1] = v0.y, nodes[2] = v0.z, nodes[3] = v0.w;
        nodes[4] = v1.x, nodes[5] = v1.y, nodes[6] = v1.z, nodes[7] = v1.w;
    }
    AtomicSink<Op> sink(mv, out, nodes, e);
    double xe[NNE][Elem<NNE>::DIM], fe[NNE];
    load_coords<NNE>(mv.X, nodes, xe);
    load_field<NNE>(Op::field(prm), nodes, fe);
    Op::run(prm, xe, fe, e, sink);
}

// complex nodal density (complex-step verification, pyfem.py:1018-1020): the same scatter with one part of the
// complex RAMP factor per pass (prm.mat.part); Poisson and quad4 elasticity operators
template <class Op>
__global__ void __launch_bounds__(128) k_assemble_atomic_complex(MeshView mv, typename Op::Params prm, Outputs out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= mv.nelems) return;
    if (mv.elem_skip != nullptr && mv.elem_skip[e]) return;
    constexpr int NNE = Op::NNE;
    int nodes[NNE];
#pragma unroll
    for (int a = 0; a < NNE; ++a) nodes[a] = __ldg(mv.conn + e * NNE + a);
    AtomicSink<Op> sink(mv, out, nodes, e);
    double xe[NNE][Elem<NNE>::DIM], fe[NNE], fi[NNE];
    load_coords<NNE>(mv.X, nodes, xe);
    load_field<NNE>(prm.mat.rho, nodes, fe);
    load_field<NNE>(prm.mat.rho_im, nodes, fi);
    Op::run_complex(prm, xe, fe, fi, sink);
}

// element matrices only (no scatter): one thread per element
template <class Op>
__global__ void __launch_bounds__(128) k_element_matrices(MeshView mv, typename Op::Params prm, double* Ke0, double* Ke1,
                                                          double* fe) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= mv.nelems) return;
    constexpr int NNE = Op::NNE;
    int nodes[NNE];
#pragma unroll
    for (int a = 0; a < NNE; ++a) nodes[a] = __ldg(mv.conn + e * NNE + a);
    GlobalSink<Op> sink{{Ke0, Ke1}, fe, e};
    double xe[NNE][Elem<NNE>::DIM], fld[NNE];
    load_coords<NNE>(mv.X, nodes, xe);
    load_field<NNE>(Op::field(prm), nodes, fld);
    Op::run(prm, xe, fld, e, sink);
}

struct HexGlobalRowSink {  // octet kernel: this thread's 3 x 24 row block of the element matrix
    double* row;  // Ke + e*576 + lane8*3*24
    PFG_DEV void block(int, int, int b, const double* blk) const {
#pragma unroll
        for (int al = 0; al < 3; ++al)
#pragma unroll
            for (int be = 0; be < 3; ++be) row[al * 24 + b * 3 + be] = blk[al * 3 + be];
    }
};

struct ElasticityHex8Tag {  // layout / sink traits of the octet kernel
    static constexpr int NNE = 8, M = 3, NMAT = 1, NVEC = 0;
};

__global__ void __launch_bounds__(128) k_elasticity_hex8_matrices(MeshView mv, ElasticityHex8Params prm, double* Ke) {
    __shared__ double stage[16 * kHexStageDoubles];
    const int64_t e = blockIdx.x * 16ll + (threadIdx.x >> 3);
    const int lane8 = threadIdx.x & 7;
    const unsigned octet_mask = 0xffu << ((threadIdx.x & 31) & ~7);
    if (e >= mv.nelems) return;
    int nodes[8];
    const int4 v0 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e);
    const int4 v1 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e + 1);
    nodes[0] = v0.x, nodes[1] = v0.y, nodes[2] = v0.z, nodes[3] = v0.w;
    nodes[4] = v1.x, nodes[5] = v1.y, nodes[6] = v1.z, nodes[7] = v1.w;
    HexGlobalRowSink sink{Ke + e * 576 + lane8 * 72};
    elasticity_hex8_octet(mv, prm, nodes, stage + (threadIdx.x >> 3) * kHexStageDoubles, lane8, octet_mask, true, sink);
}

__global__ void __launch_bounds__(128) k_elasticity_hex8_atomic(MeshView mv, ElasticityHex8Params prm, Outputs out) {
    __shared__ double stage[16 * kHexStageDoubles];
    const int64_t e = blockIdx.x * 16ll + (threadIdx.x >> 3);
    const int lane8 = threadIdx.x & 7;
    const unsigned octet_mask = 0xffu << ((threadIdx.x & 31) & ~7);
    if (e >= mv.nelems) return;  // whole octets leave together
    if (mv.elem_skip != nullptr && mv.elem_skip[e]) return;
    int nodes[8];
    const int4 v0 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e);
    const int4 v1 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e + 1);
    nodes[0] = v0.x, nodes[1] = v0.y, nodes[2] = v0.z, nodes[3] = v0.w;
    nodes[4] = v1.x, nodes[5] = v1.y, nodes[6] = v1.z, nodes[7] = v1.w;
    HexAtomicRowSink sink{nullptr, 0, mv.rank + (e * 8 + lane8) * 8};
    int my_node = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
        if (a == lane8) my_node = nodes[a];
    if (my_node >= mv.own_begin && my_node < mv.own_end && out.vals[0] != nullptr) {
        const int64_t r = my_node - mv.own_begin;
        const int64_t p0 = __ldg(mv.blk_ptr + r), p1 = __ldg(mv.blk_ptr + r + 1);
        sink.dst0 = out.vals[0] + p0 * 9;
        sink.k = (int)(p1 - p0);
    }
    elasticity_hex8_octet(mv, prm, nodes, stage + (threadIdx.x >> 3) * kHexStageDoubles, lane8, octet_mask,
                          sink.dst0 != nullptr, sink);
}

// ---------------------------------------------------------------------------------------------
// gather kernels
// ---------------------------------------------------------------------------------------------
// TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) bring the chunk's node table and plan bytes into
// shared memory while phase A integrates; completion is tracked by one mbarrier.
PFG_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PFG_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
PFG_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
PFG_DEV void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PFG_DEV void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
PFG_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA bulk store shared -> global (SASS UBLKCP.G.S), tracked by the issuing thread's bulk async-group
PFG_DEV void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
PFG_DEV void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
PFG_DEV void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
PFG_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory carve-up of a gather CTA: [mbarrier | chunk node table | chunk plan bytes | row blocks | vectors]
struct GatherSmem {
    uint64_t* bar;
    const ChunkNode* nodes;
    const uint8_t* plan;  // byte address of the chunk's first plan word
    double* rb;
    PFG_DEV GatherSmem(unsigned char* base, const MeshView& mv, const ChunkHdr& h) {
        bar = reinterpret_cast<uint64_t*>(base);
        unsigned char* nodes_s = base + 16;
        unsigned char* plan_s = nodes_s + mv.stage_nodes_bytes;
        rb = reinterpret_cast<double*>(plan_s + mv.stage_plan_bytes);
        const size_t plan_lo = (size_t)h.plan_begin * 4;
        const size_t plan_lo16 = plan_lo & ~(size_t)15;
        nodes = reinterpret_cast<const ChunkNode*>(nodes_s);
        plan = plan_s + (plan_lo - plan_lo16);
        if (threadIdx.x == 0) {
            const size_t plan_hi16 = ((size_t)(h.plan_begin + h.plan_words) * 4 + 15) & ~(size_t)15;
            const uint32_t nbytes = h.n_nodes * (uint32_t)sizeof(ChunkNode);
            const uint32_t pbytes = (uint32_t)(plan_hi16 - plan_lo16);
            mbar_init(bar, 1);
            mbar_expect_tx(bar, nbytes + pbytes);
            tma_load_1d(nodes_s, mv.cnodes + h.node_begin, nbytes, bar);
            if (pbytes) tma_load_1d(plan_s, mv.plan_pool + plan_lo16, pbytes, bar);
        }
    }
    // call after the __syncthreads that ends phase A (so every thread sees the initialised barrier)
    PFG_DEV void wait_metadata() const { mbar_wait(bar, 0); }
};
// phase B: sum plan-ordered contributions of every (chunk node, neighbour) block and store it
template <class Op>
PFG_DEV void gather_phase_b(const MeshView& mv, const ChunkHdr& h, const ChunkNode* __restrict__ cnodes_s,
                            const uint8_t* __restrict__ plan_s, const double* __restrict__ rb,
                            const double* __restrict__ vecs, const Outputs& out) {
    constexpr int NNE = Op::NNE, M = Op::M, BLK = M * M, NMAT = Op::NMAT, RB = Layout<Op>::RB;
    if constexpr (NMAT > 0) {
        const int kpad = (int)h.kpad;
        const int items = (int)h.n_nodes * kpad;
        for (int idx = threadIdx.x; idx < items; idx += blockDim.x) {
            const int p = idx / kpad;
            const int t = idx - p * kpad;
            const ChunkNode cn = cnodes_s[p];
            if (t >= cn.k) continue;
            const uint8_t* rec = plan_s + (size_t)(cn.plan - h.plan_begin) * 4;
            const int s0 = rec[t], s1 = rec[t + 1];
            const uint8_t* src = rec + cn.k + 1;
            double acc[NMAT][BLK];
#pragma unroll
            for (int mt = 0; mt < NMAT; ++mt)
#pragma unroll
                for (int i = 0; i < BLK; ++i) acc[mt][i] = 0.0;
            for (int s = s0; s < s1; ++s) {
                const int code = src[s];
                const double* q = rb + ((int)cn.inc_base + (code >> 3)) * RB + (code & 7) * BLK;
#pragma unroll
                for (int mt = 0; mt < NMAT; ++mt) {
                    if constexpr (M == 2) {
                        const double2 v0 = reinterpret_cast<const double2*>(q + mt * NNE * BLK)[0];
                        const double2 v1 = reinterpret_cast<const double2*>(q + mt * NNE * BLK)[1];
                        acc[mt][0] += v0.x, acc[mt][1] += v0.y, acc[mt][2] += v1.x, acc[mt][3] += v1.y;
                    } else {
#pragma unroll
                        for (int i = 0; i < BLK; ++i) acc[mt][i] += q[mt * NNE * BLK + i];
                    }
                }
            }
#pragma unroll
            for (int mt = 0; mt < NMAT; ++mt) {
                if (out.vals[mt] == nullptr) continue;
                double* dst = out.vals[mt] + cn.gslot + (int64_t)M * t;
#pragma unroll
                for (int al = 0; al < M; ++al) {
                    double* row = dst + (int64_t)al * M * cn.k;
                    if constexpr (M == 2) {
                        __stcs(reinterpret_cast<double2*>(row), make_double2(acc[mt][al * 2], acc[mt][al * 2 + 1]));
                    } else {
#pragma unroll
                        for (int be = 0; be < M; ++be) __stcs(row + be, acc[mt][al * M + be]);
                    }
                }
            }
        }
    }
    if constexpr (Op::NVEC > 0) {
        if (out.vec != nullptr) {
            for (int p = threadIdx.x; p < (int)h.n_nodes; p += blockDim.x) {
                const ChunkNode cn = cnodes_s[p];
                double s = 0.0;
                for (int j = 0; j < cn.valence; ++j) s += vecs[(int)cn.inc_base + j];
                out.vec[mv.cnode_id[h.node_begin + p] - mv.own_begin] = s;
            }
        }
    }
}

// ---- cp.async helpers ------------------------------------------------------------------------------
PFG_DEV void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
PFG_DEV void cp_async_8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
PFG_DEV void cp_async_4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
PFG_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
PFG_DEV void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int NNE>
PFG_DEV void load_rec_nodes(const int32_t* __restrict__ recn, int r, int (&nodes)[NNE]) {
    if constexpr (NNE == 4) {
        const int4 v = reinterpret_cast<const int4*>(recn)[r];
        nodes[0] = v.x, nodes[1] = v.y, nodes[2] = v.z, nodes[3] = v.w;
    } else {
        const int4 v0 = reinterpret_cast<const int4*>(recn)[2 * r];
        const int4 v1 = reinterpret_cast<const int4*>(recn)[2 * r + 1];
        nodes[0] = v0.x, nodes[1] = v0.y, nodes[2] = v0.z, nodes[3] = v0.w;
        nodes[4] = v1.x, nodes[5] = v1.y, nodes[6] = v1.z, nodes[7] = v1.w;
    }
}

// ---- tile kernel: persistent, software-pipelined owner-computes assembly -----------------------------
// Staging layout of one element record for operator Op (doubles): NMAT matrices of NB node-pair blocks
// (upper triangle only when the operator is symmetric), then NVEC*NNE vector entries.  The record stride S
// is chosen so that consecutive records (= consecutive lanes in phase A) hit different banks with the
// widest store the operator uses (16 B for 2x2 blocks, 8 B for scalars).
template <class Op>
struct TileStage {
    static constexpr int NNE = Op::NNE, M = Op::M, BLK = M * M;
    static constexpr bool SYM = Op::SYM;
    static constexpr int NB = SYM ? NNE * (NNE + 1) / 2 : NNE * NNE;
    static constexpr int MATD = Op::NMAT * NB * BLK;
    static constexpr int RAW = MATD + Op::NVEC * NNE;
    static constexpr bool WIDE = (M == 2 && Op::NMAT > 0);  // 16-byte code units / accesses
    static constexpr int S = WIDE ? ((((RAW + 1) / 2) % 2 == 1) ? ((RAW + 1) & ~1) : ((RAW + 1) & ~1) + 2) : (RAW | 1);
    static constexpr int UNIT_D = WIDE ? 2 : 1;  // doubles per code unit
    PFG_DEV static constexpr int block_index(int a, int b) {
        return SYM ? (a * (2 * NNE - 1 - a) / 2 + b) : (a * NNE + b);
    }
    static TileLayout layout() {
        TileLayout L;
        L.nne = NNE;
        L.unit_shift = WIDE ? 4 : 3;
        L.rec_units = S / UNIT_D;
        L.sym = SYM ? 1 : 0;
        L.blk_units = BLK / UNIT_D;
        L.has_mat = Op::NMAT > 0 ? 1 : 0;
        L.vec_units = Op::NVEC > 0 ? MATD / UNIT_D : -1;
        return L;
    }
};

template <class Op>
struct TileSink {
    using St = TileStage<Op>;
    static constexpr int NNE = Op::NNE, M = Op::M, BLK = M * M;
    double* rec;  // this record's staging area
    PFG_DEV void block(int mat, int a, int b, const double* blk) const {
        if (St::SYM && a > b) return;  // lower triangle is read transposed in phase B
        double* p = rec + (mat * St::NB + St::block_index(a, b)) * BLK;
        if constexpr (M == 2) {
            reinterpret_cast<double2*>(p)[0] = make_double2(blk[0], blk[1]);
            reinterpret_cast<double2*>(p)[1] = make_double2(blk[2], blk[3]);
        } else {
#pragma unroll
            for (int i = 0; i < BLK; ++i) p[i] = blk[i];
        }
    }
    PFG_DEV void vec(int a, double v) const { rec[St::MATD + a] = v; }
};

struct TileCfg {
    int off_dir;                  // ring of 8 TileDir entries
    int off_blob, off_codes;      // chunk blob / codes buffers (one stage each)
    int off_win, win_stride;      // node-window ids, two stages
    int off_loc, loc_stride;      // window indices of the record corners (DEEP: two stages, loc_stride apart)
    int off_x, off_field;         // coordinates / nodal field of the window nodes (DEEP: two stages)
    int x_stride, field_stride;
    int off_stage;                // element-record staging
    int off_image;                // image of the chunk's CSR values (TMA bulk-store source)
    int image_stride;             // doubles between the images of an operator's matrices (Helmholtz: K and R)
    int nchunks;
};

// phase B: one thread per chunk node walks the node's codes in groups of eight, sums each (node, neighbour) block in
// plan order and, at the block's end flag, drops its rows into the shared-memory image of the CSR values.
// Neighbouring lanes read the same block of neighbouring records, so on a regular mesh the gather runs at one
// bank-conflict-free wavefront per 128 bytes.  The running sums restart without a branch: acc = fma(acc, keep, v)
// with keep = 0 after an end flag.  Scalar handles then sum the node's vector entries (residual / right-hand side).
template <class Op, int THREADS>
PFG_DEV void tile_phase_b_nodes(const MeshView& mv, const TileHdr& h, const unsigned char* __restrict__ blob,
                                const uint16_t* __restrict__ codes, const double* __restrict__ stage,
                                double* __restrict__ image, int image_stride, const Outputs& out, int rot,
                                uint32_t row_base) {
    using St = TileStage<Op>;
    constexpr int M = Op::M, NMAT = Op::NMAT;
    static_assert(M == 1 || (M == 2 && NMAT == 1), "tile kernel: scalar operators or one matrix of 2x2 blocks");
    const TileNode* __restrict__ nodes = reinterpret_cast<const TileNode*>(blob + sizeof(TileHdr));
    const unsigned char* __restrict__ stage_b = reinterpret_cast<const unsigned char*>(stage);
    const int n_nodes = (int)h.n_nodes, gmax = (int)h.gmax;
    // a chunk has fewer nodes than the CTA has threads, so one warp idles through this phase; `rot` moves that
    // warp around from chunk to chunk so that no scheduler (warp id mod 4) is systematically under-used
    for (int p = (int)((threadIdx.x + 32u * (unsigned)rot) % (unsigned)THREADS); p < n_nodes; p += THREADS) {
        const TileNode tn = nodes[p];
        const uint4* __restrict__ cp = reinterpret_cast<const uint4*>(codes) + p;  // [group][node]
        if constexpr (NMAT > 0) {
            if constexpr (M == 2) {
                const int row_bytes = (int)tn.k * 16;  // k neighbours: 2k doubles per dof row
                unsigned char* o = reinterpret_cast<unsigned char*>(image) + (size_t)tn.aux * 16;
                double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0, keep = 0.0;
#pragma unroll 1
                for (int g = 0; g < gmax; ++g) {
                    const uint4 c = cp[(size_t)g * n_nodes];
                    const unsigned code[8] = {c.x & 0xFFFFu, c.x >> 16, c.y & 0xFFFFu, c.y >> 16,
                                              c.z & 0xFFFFu, c.z >> 16, c.w & 0xFFFFu, c.w >> 16};
                    // all sixteen loads of the group are issued before the dependent sums start
                    double2 v0[8], v1[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double2* __restrict__ q = reinterpret_cast<const double2*>(stage_b + ((code[j] & 0xFFFCu) << 2));
                        v0[j] = q[0], v1[j] = q[1];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const bool tr = St::SYM && (code[j] & 1u);
                        a00 = fma(a00, keep, v0[j].x);
                        a01 = fma(a01, keep, tr ? v1[j].x : v0[j].y);
                        a10 = fma(a10, keep, tr ? v0[j].y : v1[j].x);
                        a11 = fma(a11, keep, v1[j].y);
                        const bool end = (code[j] & 2u) != 0;  // last contribution of this block
                        if (end) {
                            *reinterpret_cast<double2*>(o) = make_double2(a00, a01);
                            *reinterpret_cast<double2*>(o + row_bytes) = make_double2(a10, a11);
                            o += 16;
                        }
                        keep = __hiloint2double(end ? 0 : 0x3FF00000, 0);
                    }
                }
            } else {
                double* o = image + tn.aux;
                double acc[NMAT], keep = 0.0;
#pragma unroll
                for (int mt = 0; mt < NMAT; ++mt) acc[mt] = 0.0;
#pragma unroll 1
                for (int g = 0; g < gmax; ++g) {
                    const uint4 c = cp[(size_t)g * n_nodes];
                    const unsigned code[8] = {c.x & 0xFFFFu, c.x >> 16, c.y & 0xFFFFu, c.y >> 16,
                                              c.z & 0xFFFFu, c.z >> 16, c.w & 0xFFFFu, c.w >> 16};
                    double v[8][NMAT];  // the group's loads are issued before the dependent sums start
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double* __restrict__ q = reinterpret_cast<const double*>(stage_b + ((code[j] & 0xFFFCu) << 1));
#pragma unroll
                        for (int mt = 0; mt < NMAT; ++mt) v[j][mt] = q[mt * St::NB];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
#pragma unroll
                        for (int mt = 0; mt < NMAT; ++mt) acc[mt] = fma(acc[mt], keep, v[j][mt]);
                        const bool end = (code[j] & 2u) != 0;
                        if (end) {
#pragma unroll
                            for (int mt = 0; mt < NMAT; ++mt) o[mt * image_stride] = acc[mt];
                            o += 1;
                        }
                        keep = __hiloint2double(end ? 0 : 0x3FF00000, 0);
                    }
                }
            }
        }
        if constexpr (Op::NVEC > 0) {
            if (out.vec != nullptr) {  // the node's incidences, in element order; padding codes read the zero record
                const uint2* __restrict__ vp = reinterpret_cast<const uint2*>(codes + (size_t)gmax * n_nodes * 8) + p;
                const uint32_t* __restrict__ row = reinterpret_cast<const uint32_t*>(nodes + n_nodes);
                double sum = 0.0;
                for (int g = 0; g < (int)h.gvmax; ++g) {
                    const uint2 c = vp[(size_t)g * n_nodes];
                    sum += *reinterpret_cast<const double*>(stage_b + (((c.x & 0xFFFFu) & 0xFFFCu) << 1));
                    sum += *reinterpret_cast<const double*>(stage_b + (((c.x >> 16) & 0xFFFCu) << 1));
                    sum += *reinterpret_cast<const double*>(stage_b + (((c.y & 0xFFFFu) & 0xFFFCu) << 1));
                    sum += *reinterpret_cast<const double*>(stage_b + (((c.y >> 16) & 0xFFFCu) << 1));
                }
                out.vec[row_base + row[p]] = sum;
            }
        }
    }
}

// One run of consecutive node ids: image -> CSR values.  2x2 blocks: everything is 16-byte aligned, one TMA bulk
// store.  Scalars: the run sits on the same 16-byte phase in the image as in the CSR values; its aligned middle part
// leaves as a bulk store, a leading / trailing odd value by a plain store.
template <int M>
PFG_DEV void tile_store_run(double* __restrict__ vals, const double* __restrict__ image, int64_t gbase, const TileRun& run) {
    if constexpr (M == 2) {
        tma_store_1d(vals + gbase + run.gslot_rel, image + (size_t)run.out_off * 2, (uint32_t)run.len * 16u);
    } else {
        double* g = vals + gbase + run.gslot_rel;
        const double* s = image + run.out_off;
        const int len = (int)run.len;
        const int head = (int)((gbase + run.gslot_rel) & 1);
        if (head) g[0] = s[0];
        const int body = (len - head) & ~1;
        if (body > 0) tma_store_1d(g + head, s + head, (uint32_t)body * 8u);
        if ((len - head) & 1) g[len - 1] = s[len - 1];
    }
}

PFG_DEV void cp_async_mbar_arrive(uint64_t* bar) {  // the mbarrier sees this thread's earlier cp.async copies land
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent CTA over a contiguous range of chunks.  Per chunk i:
//   phase A  one thread per element record: coordinates (and nodal field) come from the chunk's node window in
//            shared memory, the element matrices go to the record staging;
//   phase B  plan-ordered sums -> CSR values (row format: shared-memory image + TMA bulk stores).
// Everything a chunk needs is in flight one or two chunks ahead and is tracked by mbarriers, never by
// register scoreboards: TMA bulk copies bring the directory-addressed blob + codes (one ahead), the corner
// indices (one ahead) and the window's node ids (two ahead); cp.async gathers the window's coordinates one
// chunk ahead and reports to an mbarrier.
// Operators with `static constexpr bool DEEP_PREFETCH = true` double-buffer the window coordinates / field and the
// corner indices and fetch them a whole chunk ahead (costs 3-4 KB of shared memory: pays for the register-limited
// nonlinear Poisson kernel, 1.99 -> 1.90 ms, not where it would cost a CTA per SM).
template <class Op, class = void>
struct op_deep_prefetch : std::false_type {};
template <class Op>
struct op_deep_prefetch<Op, std::enable_if_t<Op::DEEP_PREFETCH>> : std::true_type {};
// `static constexpr bool DEEP_X = true`: the operator MAY run the XD variant of the kernel, in which only the window
// coordinates / field are double-buffered and gathered a whole chunk ahead (2-3 KB of shared memory for the 2 x 2-block
// kernel; the corner indices keep their single stage).  With the plan tables resident (chunk templates) the coordinate
// gather is the one transfer left on the critical path: 1.374 -> 1.349 ms on the 16.8 M-quad case.  The launcher takes
// the variant only when the extra stage costs no resident CTA (with a nodal field it would: 3 -> 2 per SM).
template <class Op, class = void>
struct op_deep_x : std::false_type {};
template <class Op>
struct op_deep_x<Op, std::enable_if_t<Op::DEEP_X>> : std::true_type {};

// Operators with more than one nodal field (sensitivities: rho, phi, psi) declare `static constexpr int FW` doubles
// per window node and a `gather_fields(prm, node, dst)` that issues the cp.async copies of one node.
template <class Op, class = void>
struct op_field_width : std::integral_constant<int, 1> {};
template <class Op>
struct op_field_width<Op, std::enable_if_t<(Op::FW > 1)>> : std::integral_constant<int, Op::FW> {};

template <class Op, int THREADS, int MINB, bool XD>
__global__ void __launch_bounds__(THREADS, MINB)
    k_tile(MeshView mv, typename Op::Params prm, Outputs out, TileCfg cfg) {
    constexpr bool DEEP = op_deep_prefetch<Op>::value;  // corner indices double-buffered, fetched at the top
    constexpr bool DEEPX = DEEP || XD;                  // window data double-buffered, gathered at the top
    extern __shared__ __align__(128) unsigned char smem[];
    using St = TileStage<Op>;
    constexpr int NNE = Op::NNE, DIM = Elem<NNE>::DIM, FW = op_field_width<Op>::value;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0] blob+codes, [1..2] window ids, [3] (and [5]) corner indices, [4] window data
    const int64_t c_begin = (int64_t)cfg.nchunks * blockIdx.x / gridDim.x;
    const int64_t c_end = (int64_t)cfg.nchunks * (blockIdx.x + 1) / gridDim.x;
    const int nloc = (int)(c_end - c_begin);
    if (nloc <= 0) return;
    const double* __restrict__ field = Op::field(prm);
    TileDir* dir_s = reinterpret_cast<TileDir*>(smem + cfg.off_dir);
    unsigned char* blob_s = smem + cfg.off_blob;
    uint16_t* codes_s = reinterpret_cast<uint16_t*>(smem + cfg.off_codes);
    auto loc_stage = [&](int j) -> unsigned char* { return smem + cfg.off_loc + (DEEP ? (j & 1) * cfg.loc_stride : 0); };
    auto x_stage = [&](int j) -> double* {
        return reinterpret_cast<double*>(smem + cfg.off_x + (DEEPX ? (j & 1) * cfg.x_stride : 0));
    };
    auto f_stage = [&](int j) -> double* {
        return reinterpret_cast<double*>(smem + cfg.off_field + (DEEPX ? (j & 1) * cfg.field_stride : 0));
    };
    double* stage = reinterpret_cast<double*>(smem + cfg.off_stage);
    double* image = reinterpret_cast<double*>(smem + cfg.off_image);
    const TileDir* __restrict__ dir_g = mv.tile_dir + c_begin;

    auto n_recs_of = [&](int j) -> int { return (int)dir_s[j & 7].n_recs; };
    auto n_win_of = [&](int j) -> int { return (int)dir_s[j & 7].n_win; };
    auto same_tmpl = [&](int j, int back) -> bool { return j >= back && dir_s[j & 7].tmpl == dir_s[(j - back) & 7].tmpl; };
    auto win_stage = [&](int j) -> unsigned char* { return smem + cfg.off_win + (j & 1) * cfg.win_stride; };
    // thread 0: bulk copies of chunk j's tables.  A table buffer still holds the tables of the chunk that used it last:
    // when chunk j runs on the same TEMPLATE (pfg_internal.cuh) nothing is fetched and the barrier's phase is completed
    // by a plain arrive -- on a lattice-like mesh a CTA loads its tables a handful of times per launch.
    // Sources are only 4- or 8-byte aligned, so each copy starts at the enclosing 16-byte boundary and the readers
    // skip the same number of bytes (the pools are padded at both ends).
    auto issue_win = [&](int j) {  // two stages: the last user of stage j & 1 was chunk j - 2
        uint64_t* bar = &bars[1 + (j & 1)];
        if (same_tmpl(j, 2)) return mbar_arrive(bar);
        const size_t lo = (size_t)dir_s[j & 7].win_off * 4, lo16 = lo & ~(size_t)15;
        const uint32_t bytes = (uint32_t)(((lo + (size_t)n_win_of(j) * 4 + 15) & ~(size_t)15) - lo16);
        mbar_expect_tx(bar, bytes);
        if (bytes) tma_load_1d(win_stage(j), reinterpret_cast<const unsigned char*>(mv.win_nodes) + lo16, bytes, bar);
    };
    // DEEP: the corner indices of chunk i + 1 are issued at the top of iteration i, while slower threads may still be
    // waiting for those of chunk i -- with one barrier an instantly completed phase (same template: plain arrive) could
    // overtake such a waiter, so each of the two stages has its own barrier.  Single stage: one barrier, the issue
    // comes after the __syncthreads that ends phase A.
    auto loc_bar = [&](int j) -> uint64_t* { return &bars[(DEEP && (j & 1)) ? 5 : 3]; };
    auto loc_parity = [&](int j) -> uint32_t { return (uint32_t)((DEEP ? (j >> 1) : j) & 1); };
    auto issue_loc = [&](int j) {
        uint64_t* bar = loc_bar(j);
        if (same_tmpl(j, DEEP ? 2 : 1)) return mbar_arrive(bar);
        const size_t lo = (size_t)dir_s[j & 7].loc_off * NNE * 2, lo16 = lo & ~(size_t)15;
        const uint32_t bytes = (uint32_t)(((lo + (size_t)n_recs_of(j) * NNE * 2 + 15) & ~(size_t)15) - lo16);
        mbar_expect_tx(bar, bytes);
        if (bytes) tma_load_1d(loc_stage(j), reinterpret_cast<const unsigned char*>(mv.rec_local) + lo16, bytes, bar);
    };
    auto issue_meta = [&](int j) {
        if (same_tmpl(j, 1)) return mbar_arrive(&bars[0]);
        const TileDir t = dir_s[j & 7];
        const uint32_t bb = (uint32_t)t.blob_len16 * 16u, cb = (uint32_t)t.code_len16 * 16u;
        mbar_expect_tx(&bars[0], bb + cb);
        tma_load_1d(blob_s, mv.tile_blob + (size_t)t.blob_off16 * 16, bb, &bars[0]);
        if (cb) tma_load_1d(codes_s, reinterpret_cast<const unsigned char*>(mv.tile_codes) + (size_t)t.code_off16 * 16, cb, &bars[0]);
    };
    // all threads: gather the coordinates (and nodal field) of chunk j's window nodes, one node per thread
    auto gather_window = [&](int j) {
        const uint32_t* __restrict__ win =
            reinterpret_cast<const uint32_t*>(win_stage(j) + (((size_t)dir_s[j & 7].win_off * 4) & 15));
        const int n_win = n_win_of(j);
        const size_t node_base = dir_s[j & 7].node_base;
        double* xs = x_stage(j);
        double* fs = f_stage(j);
        for (int t = threadIdx.x; t < n_win; t += THREADS) {
            const size_t node = node_base + win[t];
            double* dst = xs + (size_t)t * DIM;
            if constexpr (DIM == 2) {
                cp_async_16(dst, mv.X + node * 2);
            } else {
                cp_async_8(dst, mv.X + node * 3);
                cp_async_8(dst + 1, mv.X + node * 3 + 1);
                cp_async_8(dst + 2, mv.X + node * 3 + 2);
            }
            if constexpr (FW > 1) Op::gather_fields(prm, node, fs + (size_t)t * FW);
            else if (field != nullptr) cp_async_8(fs + t, field + node);
        }
        cp_async_mbar_arrive(&bars[4]);
    };
    auto copy_dir_piece = [&](int j, int piece) {  // 16 bytes of directory entry j into the ring
        cp_async_16(reinterpret_cast<unsigned char*>(&dir_s[j & 7]) + 16 * piece,
                    reinterpret_cast<const unsigned char*>(&dir_g[j]) + 16 * piece);
    };

    for (int i = threadIdx.x; i < St::S; i += THREADS) stage[i] = 0.0;  // record slot 0: zeros (code 0)
    if (threadIdx.x < 15 && (int)threadIdx.x / 3 < nloc) {               // directory entries 0 .. 4
        copy_dir_piece((int)threadIdx.x / 3, (int)threadIdx.x % 3);
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) mbar_init(&bars[s], 1);
        mbar_init(&bars[4], THREADS);
        mbar_init(&bars[5], 1);
        issue_win(0);
        if (nloc > 1) issue_win(1);
        issue_loc(0);
        issue_meta(0);
    }
    __syncthreads();
    mbar_wait(&bars[1], 0);
    gather_window(0);

    for (int i = 0; i < nloc; ++i) {
        // directory entry i + 5 rides on this thread's next arrive on bars[4]
        if (threadIdx.x < 3 && i + 5 < nloc) copy_dir_piece(i + 5, (int)threadIdx.x);
        const int n_recs = n_recs_of(i);
        // skip flags of masked elements (halo.py's reduce variant): fetched before the waits, used in phase A
        const uint8_t* __restrict__ skip_g = mv.rec_skip != nullptr ? mv.rec_skip + dir_s[i & 7].rec_begin : nullptr;
        const unsigned skip_mine = (skip_g != nullptr && (int)threadIdx.x < n_recs) ? skip_g[threadIdx.x] : 0u;
        mbar_wait(&bars[4], i & 1);  // the window's coordinates have landed (all threads' copies)
        mbar_wait(loc_bar(i), loc_parity(i));  // and so have the corner indices
        // ---- prefetch: window ids two chunks ahead; corner indices and window coordinates one chunk ahead
        auto prefetch_window = [&]() {
            if (threadIdx.x == 0 && i + 2 < nloc) issue_win(i + 2);
            if (i + 1 < nloc) {
                mbar_wait(&bars[1 + ((i + 1) & 1)], ((i + 1) >> 1) & 1);
                gather_window(i + 1);
            }
        };
        auto prefetch_corners = [&]() {
            if (threadIdx.x == 0 && i + 1 < nloc) issue_loc(i + 1);
        };
        // into the other stage: phases A and B of this chunk to land
        if constexpr (DEEPX) prefetch_window();
        if constexpr (DEEP) prefetch_corners();
        // ---- phase A: one thread per element record -> staged element matrices
        const unsigned char* loc_i = loc_stage(i) + (((size_t)dir_s[i & 7].loc_off * NNE * 2) & 15);
        const double* xs = x_stage(i);
        const double* fs = f_stage(i);
        for (int r = threadIdx.x; r < n_recs; r += THREADS) {
            TileSink<Op> sink{stage + (size_t)(r + 1) * St::S};  // slot 0 is the all-zero record
            if (skip_g != nullptr && (r == (int)threadIdx.x ? skip_mine : (unsigned)skip_g[r])) {
                // masked element (another rank integrates it): its record contributes zeros
                for (int i = 0; i < St::RAW; ++i) sink.rec[i] = 0.0;
                continue;
            }
            unsigned loc[NNE];
            if constexpr (NNE == 4) {
                const uint2 v = reinterpret_cast<const uint2*>(loc_i)[r];
                loc[0] = v.x & 0xFFFFu, loc[1] = v.x >> 16, loc[2] = v.y & 0xFFFFu, loc[3] = v.y >> 16;
            } else {
                const uint4 v = reinterpret_cast<const uint4*>(loc_i)[r];
                loc[0] = v.x & 0xFFFFu, loc[1] = v.x >> 16, loc[2] = v.y & 0xFFFFu, loc[3] = v.y >> 16;
                loc[4] = v.z & 0xFFFFu, loc[5] = v.z >> 16, loc[6] = v.w & 0xFFFFu, loc[7] = v.w >> 16;
            }
            double xe[NNE][DIM], fe[NNE * FW];
#pragma unroll
            for (int a = 0; a < NNE; ++a) {
                const double* src = xs + (size_t)loc[a] * DIM;
                if constexpr (DIM == 2) {
                    const double2 v = *reinterpret_cast<const double2*>(src);
                    xe[a][0] = v.x, xe[a][1] = v.y;
                } else {
                    xe[a][0] = src[0], xe[a][1] = src[1], xe[a][2] = src[2];
                }
#pragma unroll
                for (int c = 0; c < FW; ++c) fe[a * FW + c] = (field != nullptr) ? fs[loc[a] * FW + c] : 0.0;
            }
            int64_t elem = 0;
            if constexpr (Op::NEEDS_ELEM) elem = __ldg(mv.rec_elem + dir_s[i & 7].rec_begin + r);
            Op::run(prm, xe, fe, elem, sink);
        }
        tma_store_wait_read();  // the previous chunk's bulk stores have read the image
        __syncthreads();
        // single stage: free once phase A of this chunk has read it
        if constexpr (!DEEPX) prefetch_window();
        if constexpr (!DEEP) prefetch_corners();
        // ---- phase B: plan-ordered sums into the CSR image, each CSR value written once
        mbar_wait(&bars[0], i & 1);
        {
            const TileHdr h = *reinterpret_cast<const TileHdr*>(blob_s);
            const int64_t gbase = dir_s[i & 7].gbase;
            tile_phase_b_nodes<Op, THREADS>(mv, h, blob_s, codes_s, stage, image, cfg.image_stride, out,
                                            (i + (int)blockIdx.x) & 3, dir_s[i & 7].row_base);
            // runs of consecutive node ids leave as TMA bulk stores, spread over the warps' leading lanes;
            // the run entry is read before the barrier so that the blob may be overwritten right after it
            // (a chunk has at most 128 runs: checked when the plan is built)
            constexpr int NW = THREADS / 32;
            const int my_run = (int)(threadIdx.x & 31) * NW + (int)(threadIdx.x >> 5);
            TileRun run;
            run.len = 0;
            if (Op::NMAT > 0 && my_run < (int)h.n_runs)
                run = reinterpret_cast<const TileRun*>(blob_s + tile_blob_tables(h.n_nodes, Op::M))[my_run];
            fence_proxy_async();  // image writes (generic proxy) -> visible to the bulk-copy engine
            __syncthreads();
            if (threadIdx.x == 0 && i + 1 < nloc) issue_meta(i + 1);
            if constexpr (Op::NMAT > 0) {
                if (run.len) {
#pragma unroll
                    for (int mt = 0; mt < Op::NMAT; ++mt)
                        if (out.vals[mt] != nullptr) tile_store_run<Op::M>(out.vals[mt], image + (size_t)mt * cfg.image_stride, gbase, run);
                }
                tma_store_commit();
            }
        }
    }
    tma_store_wait_read();
}

struct ElasticityHex8GatherOp : ElasticityHex8Tag {};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_elasticity_hex8_gather(MeshView mv, ElasticityHex8Params prm, Outputs out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using L = Layout<ElasticityHex8GatherOp>;
    const ChunkHdr h = mv.chunks[blockIdx.x];
    const GatherSmem sm(smem_raw, mv, h);
    double* rb = sm.rb;
    double* stage = rb + (size_t)h.n_inc * L::RB;  // [THREADS/8][kHexStageDoubles]
    const int lane8 = threadIdx.x & 7;
    const unsigned octet_mask = 0xffu << ((threadIdx.x & 31) & ~7);
    for (int r = threadIdx.x >> 3; r < (int)h.n_recs; r += THREADS / 8) {
        const int64_t rr = h.rec_begin + r;
        int nodes[8];
        const int4 v0 = __ldg(reinterpret_cast<const int4*>(mv.rec_nodes) + 2 * rr);
        const int4 v1 = __ldg(reinterpret_cast<const int4*>(mv.rec_nodes) + 2 * rr + 1);
        nodes[0] = v0.x, nodes[1] = v0.y, nodes[2] = v0.z, nodes[3] = v0.w;
        nodes[4] = v1.x, nodes[5] = v1.y, nodes[6] = v1.z, nodes[7] = v1.w;
        const uint16_t my_dst = __ldg(mv.rec_dst + rr * 8 + lane8);
        HexSmemRowSink sink{rb + (int)my_dst * L::RB};
        if (mv.elem_skip != nullptr && mv.elem_skip[__ldg(mv.rec_elem + rr)]) {  // integrated by another rank
            if (my_dst != kNoDst)
                for (int i = 0; i < 72; ++i) sink.row[i] = 0.0;
            continue;
        }
        elasticity_hex8_octet(mv, prm, nodes, stage + (threadIdx.x >> 3) * kHexStageDoubles, lane8, octet_mask,
                              my_dst != kNoDst, sink);
    }
    __syncthreads();
    sm.wait_metadata();
    gather_phase_b<ElasticityHex8GatherOp>(mv, h, sm.nodes, sm.plan, sm.rb, nullptr, out);
}

// ---- hex8 3-D elasticity, owner-computes: geometry pass + node-row pass -----------------------------------
// Pass 1: one lane per (element, quadrature point) -> hex_geo[(e*8 + q)*10 ..] (adjugate of J and s_q).
__global__ void __launch_bounds__(256) k_hex8_geometry(MeshView mv, ElasticityHex8Params prm, double* __restrict__ geo) {
    const int64_t t = blockIdx.x * 256ll + threadIdx.x;
    const int64_t e = t >> 3;
    if (e >= mv.nelems) return;
    int nodes[8];
    const int4 v0 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e);
    const int4 v1 = __ldg(reinterpret_cast<const int4*>(mv.conn) + 2 * e + 1);
    nodes[0] = v0.x, nodes[1] = v0.y, nodes[2] = v0.z, nodes[3] = v0.w;
    nodes[4] = v1.x, nodes[5] = v1.y, nodes[6] = v1.z, nodes[7] = v1.w;
    const bool skip = mv.elem_skip != nullptr && mv.elem_skip[e];
    hex8_geometry_point(mv, prm.mat, nodes, (int)(t & 7), skip, geo + t * kHexGeoDoubles);
}

// Pass 2: persistent CTAs over contiguous ranges of row chunks (first-format plan: ~28 nodes and the element
// records touching them).  A producer warp brings the records' geometry into a two-stage shared-memory ring by
// TMA bulk copies (640 B each, full / empty mbarriers); seven consumer warps free-run over the chunks.  Eight
// lanes serve a chunk node, one per incident element (a node of this path has at most eight).  A lane forms its
// element's 3 x 24 row block for the node and adds the 3 x 3 blocks into the warp's shared-memory image of its
// four nodes' CSR rows at the slots the rank map gives; the lanes of a node take the same local column index at
// the same time, which puts them on different neighbours (checked when the mesh handle is built:
// MeshDev::hex_rows_ok).  Each node's rows leave as one contiguous run of the CSR values.  Node and incidence
// tables of the next chunk are fetched while the current one is computed.  No atomics, no zero-fill of the
// output, every CSR value written once, bitwise reproducible.
#ifndef PFG_HEX_NB
#define PFG_HEX_NB 0  // row products of a (node, element) lane.  0: through the modes of the trilinear basis -- 63 running sums
#endif                // against seven modes whose values cost 15 additions per point, node blocks by a butterfly at the end
                      // (244 registers, no spills): 128^3 hex 3.890 -> 3.281 ms.  7: seven column nodes with their
                      // gradients (63 FMAs per point), the eighth block from the zero row sums (255 registers, 116 B of
                      // spills); 4: two passes of four column nodes (36 sums each, geometry read twice); 8: one pass with
                      // 72 sums (spills: 4.16 ms)
constexpr int kHexRowsThreads = (kHexRowWarps + 1) * 32;
constexpr int kHexGeoBytes = 8 * kHexGeoDoubles * (int)sizeof(double);  // 640 B per element
constexpr int kHexGeoStride = kHexGeoRecordBytes / 8;  // doubles per staged record: 656 B keeps 16-byte reads of
                                                       // consecutive records on different banks
static_assert(kHexGeoRecordBytes == (8 * kHexGeoDoubles + 2) * 8, "hex8 geometry record layout");

PFG_DEV void hex8_add_block(const ElasticityHex8Params& prm, const double (&P)[3][3], int b, bool active,
                            double* __restrict__ rows, int k3, uint2 ranks) {
    if (active) {
        double blk[9];
        hex8_apply_c0(prm, P, blk);
        const unsigned rk = (b < 4) ? ranks.x : ranks.y;  // four 8-bit ranks per word
        double* dst = rows + 3 * (int)((rk >> (8 * (b & 3))) & 0xFFu);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) dst[i * k3 + j] += blk[i * 3 + j];
    }
    __syncwarp();  // the next column index may land on a block another lane of the node has just updated
}

template <int B0, int NB>
PFG_DEV void hex8_rows_part(const ElasticityHex8Params& prm, const double* __restrict__ geo_e, double sx8, double sy8,
                            double sz8, bool active, double* __restrict__ rows, int k3, uint2 ranks) {
    double P[NB][3][3];
    if (active) hex8_row_products<B0, NB>(geo_e, sx8, sy8, sz8, P);
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) hex8_add_block(prm, P[bb], B0 + bb, active, rows, k3, ranks);
}

// One pass over the quadrature points for seven column nodes; the eighth block follows from the zero row sums of the
// element matrix (sum_b grad N_b = 0 at every point): 63 running sums fit the register file where 72 do not, and the
// element's geometry is read from shared memory once instead of twice.
PFG_DEV void hex8_rows_closed(const ElasticityHex8Params& prm, const double* __restrict__ geo_e, double sx8, double sy8,
                              double sz8, bool active, double* __restrict__ rows, int k3, uint2 ranks) {
    double P[7][3][3], L[3][3];
    if (active) {
        hex8_row_products<0, 7>(geo_e, sx8, sy8, sz8, P);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double sum = P[0][i][j];
#pragma unroll
                for (int bb = 1; bb < 7; ++bb) sum += P[bb][i][j];
                L[i][j] = -sum;
            }
    }
#pragma unroll
    for (int bb = 0; bb < 7; ++bb) hex8_add_block(prm, P[bb], bb, active, rows, k3, ranks);
    hex8_add_block(prm, L, 7, active, rows, k3, ranks);
}

// The same through the modes of the trilinear basis (pfg_elem.cuh: hex8_row_modes): 15 additions per point instead of
// the 63 FMAs of seven column gradients, a butterfly per entry at the end; `prm` carries C0 / 8.
PFG_DEV void hex8_rows_modal(const ElasticityHex8Params& prm, const double* __restrict__ geo_e, double sx8, double sy8,
                             double sz8, bool active, double* __restrict__ rows, int k3, uint2 ranks) {
    double P[8][3][3];
    if (active) {
        double V[7][3][3];
        hex8_row_modes(geo_e, sx8, sy8, sz8, V);
        hex8_modes_to_nodes(V, P);
    }
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) hex8_add_block(prm, P[bb], bb, active, rows, k3, ranks);
}

struct HexRowsCfg {
    int off_geo, geo_stage_bytes;  // two geometry stages
    int off_image, image_stride;   // per consumer warp: four nodes x image_stride doubles (9 * max neighbours)
    int off_meta;                  // per consumer thread 2 x 32 B of node / incidence tables, per warp 2 x 8 B
    int nchunks;
};

struct HexLaneMeta {  // what a lane needs for its (node, incident element) pair
    int64_t gslot;    // first CSR value of the node
    int k9;           // 9 * neighbours of the node (0: no node)
    int ra;           // record * 8 + local node, 0xFFFF: no incidence
    uint2 ranks;      // rank[(e*8 + a)*8 + 0..7]
};

__global__ void __launch_bounds__(kHexRowsThreads, 1)
    k_hex8_chunk_rows(MeshView mv, ElasticityHex8Params prm, const double* __restrict__ geo,
                      const uint32_t* __restrict__ inc_rec8, const uint2* __restrict__ inc_ranks8,
                      double* __restrict__ vals, HexRowsCfg cfg) {
    extern __shared__ __align__(128) unsigned char hex_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(hex_smem);  // [0..1] stage full, [2..3] stage empty
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c_begin = (int64_t)cfg.nchunks * blockIdx.x / gridDim.x;
    const int nloc = (int)((int64_t)cfg.nchunks * (blockIdx.x + 1) / gridDim.x - c_begin);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1), mbar_init(&bars[1], 1);
        mbar_init(&bars[2], kHexRowWarps), mbar_init(&bars[3], kHexRowWarps);
    }
    __syncthreads();
    if (warp == kHexRowWarps) {
        // ---- producer warp: geometry of chunk i's records -> stage i & 1
        for (int i = 0; i < nloc; ++i) {
            const int s = i & 1, f = i >> 1;
            const ChunkHdr h = mv.chunks[c_begin + i];
            const int e_lo = ((int)h.n_recs > lane) ? __ldg(mv.rec_elem + h.rec_begin + lane) : 0;
            mbar_wait(&bars[2 + s], (f & 1) ^ 1);  // consumers have left the stage's previous chunk
            if (lane == 0) mbar_expect_tx(&bars[s], h.n_recs * (uint32_t)kHexGeoBytes);
            __syncwarp();
            unsigned char* stage = hex_smem + cfg.off_geo + s * cfg.geo_stage_bytes;
            for (int r = lane; r < (int)h.n_recs; r += 32) {
                const int e = (r < 32) ? e_lo : __ldg(mv.rec_elem + h.rec_begin + r);
                tma_load_1d(stage + (size_t)r * (kHexGeoStride * 8), geo + (size_t)e * (8 * kHexGeoDoubles),
                            (uint32_t)kHexGeoBytes, &bars[s]);
            }
        }
        return;
    }
    // ---- consumer warps
#if PFG_HEX_NB == 0
    ElasticityHex8Params prm8 = prm;  // the modal row products carry a factor 8
    prm8.c11 *= 0.125, prm8.c12 *= 0.125, prm8.c44 *= 0.125;
#endif
    const int j = lane & 7;
    const int64_t nown = mv.own_end - mv.own_begin;
    double* image = reinterpret_cast<double*>(hex_smem + cfg.off_image) + (size_t)warp * 4 * cfg.image_stride;
    double* rows = image + (size_t)(lane >> 3) * cfg.image_stride;
    // Node / incidence tables travel one chunk ahead by cp.async into a per-thread slot (two buffers), the next
    // chunk's (node_begin, n_nodes) into a per-warp slot: no register is held across a chunk's arithmetic.
    unsigned char* meta_s = hex_smem + cfg.off_meta + (size_t)threadIdx.x * 64;
    unsigned char* hd_s = hex_smem + cfg.off_meta + (size_t)kHexRowWarps * 32 * 64 + (size_t)warp * 16;
    auto issue_meta = [&](int64_t first_slot, int i_chunk) {  // tables of local chunk i_chunk, whose nodes start at first_slot
        const int buf = i_chunk & 1;
        const int64_t slot = min(first_slot + warp * 4 + (lane >> 3), nown - 1);  // past the chunk: loaded, not used
        cp_async_16(meta_s + buf * 32, mv.cnodes + slot);
        cp_async_8(meta_s + buf * 32 + 16, inc_ranks8 + slot * 8 + j);
        cp_async_4(meta_s + buf * 32 + 24, inc_rec8 + slot * 8 + j);
        if (lane == 0 && i_chunk < nloc) cp_async_8(hd_s + buf * 8, mv.chunks + c_begin + i_chunk);
        cp_async_commit();
    };
    auto read_meta = [&](uint2 hd, int buf) -> HexLaneMeta {
        HexLaneMeta m;
        m.gslot = 0, m.k9 = 0, m.ra = 0xFFFF, m.ranks = make_uint2(0u, 0u);
        if (warp * 4 + (lane >> 3) < (int)hd.y) {
            const ChunkNode cn = *reinterpret_cast<const ChunkNode*>(meta_s + buf * 32);
            m.gslot = cn.gslot;
            m.k9 = (int)cn.k * 9;
            m.ranks = *reinterpret_cast<const uint2*>(meta_s + buf * 32 + 16);
            m.ra = (int)*reinterpret_cast<const uint32_t*>(meta_s + buf * 32 + 24);
        }
        return m;
    };
    auto load_meta = [&](uint2 hd, int p0) -> HexLaneMeta {  // direct loads: second round of a large chunk
        HexLaneMeta m;
        m.gslot = 0, m.k9 = 0, m.ra = 0xFFFF, m.ranks = make_uint2(0u, 0u);
        const int p = p0 + (lane >> 3);
        if (p < (int)hd.y) {
            const size_t slot = (size_t)hd.x + p;
            const ChunkNode cn = mv.cnodes[slot];
            m.gslot = cn.gslot;
            m.k9 = (int)cn.k * 9;
            m.ra = (int)__ldg(inc_rec8 + slot * 8 + j);
            m.ranks = __ldg(inc_ranks8 + slot * 8 + j);
        }
        return m;
    };
    if (nloc > 0) issue_meta((int64_t)__ldg(&mv.chunks[c_begin].node_begin), 0);
    for (int i = 0; i < nloc; ++i) {
        const int s = i & 1, f = i >> 1;
        cp_async_wait<0>();
        __syncwarp();
        const uint2 hd = *reinterpret_cast<const uint2*>(hd_s + s * 8);  // (node_begin, n_nodes)
        HexLaneMeta meta = read_meta(hd, s);
        issue_meta((int64_t)hd.x + hd.y, i + 1);
        const double* geo_s = reinterpret_cast<const double*>(hex_smem + cfg.off_geo + s * cfg.geo_stage_bytes);
        bool waited = false;
        for (int p0 = warp * 4; p0 < (int)hd.y; p0 += kHexRowWarps * 4) {
            if (p0 != warp * 4) meta = load_meta(hd, p0);
            for (int t = lane; t < 4 * cfg.image_stride; t += 32) image[t] = 0.0;
            const bool active = meta.ra != 0xFFFF;
            const int a = meta.ra & 7;
            const double sx8 = ((a & 3) == 1 || (a & 3) == 2) ? 0.125 : -0.125;  // local coordinate signs / 8
            const double sy8 = ((a & 3) >= 2) ? 0.125 : -0.125, sz8 = (a >= 4) ? 0.125 : -0.125;
            const double* geo_e = geo_s + (size_t)(active ? (meta.ra >> 3) : 0) * kHexGeoStride;
            if (!waited) mbar_wait(&bars[s], f & 1);  // the chunk's geometry has landed
            waited = true;
            __syncwarp();
#if PFG_HEX_NB == 0
            hex8_rows_modal(prm8, geo_e, sx8, sy8, sz8, active, rows, meta.k9 / 3, meta.ranks);
#elif PFG_HEX_NB == 7
            hex8_rows_closed(prm, geo_e, sx8, sy8, sz8, active, rows, meta.k9 / 3, meta.ranks);
#elif PFG_HEX_NB == 8
            hex8_rows_part<0, 8>(prm, geo_e, sx8, sy8, sz8, active, rows, meta.k9 / 3, meta.ranks);
#else
            hex8_rows_part<0, 4>(prm, geo_e, sx8, sy8, sz8, active, rows, meta.k9 / 3, meta.ranks);
            hex8_rows_part<4, 4>(prm, geo_e, sx8, sy8, sz8, active, rows, meta.k9 / 3, meta.ranks);
#endif
            // the warp's four nodes, one after the other: 256 contiguous bytes per store instruction
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int64_t g = __shfl_sync(0xffffffffu, meta.gslot, q4 * 8);
                const int n = __shfl_sync(0xffffffffu, meta.k9, q4 * 8);
                const double* __restrict__ src = image + (size_t)q4 * cfg.image_stride;
                for (int t = lane; t < n; t += 32) __stcs(vals + g + t, src[t]);
            }
            __syncwarp();  // the image is zeroed again by the next round
        }
        if (!waited) mbar_wait(&bars[s], f & 1);  // keep the ring in step even without a node in this chunk
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2 + s]);
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// auxiliary kernels
// ---------------------------------------------------------------------------------------------
template <int NNE>
__global__ void k_quad_points(MeshView mv, double* __restrict__ Xq) {  // utils.compute_elem_interp (utils.py:203-221)
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= mv.nelems) return;
    constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
    int nodes[NNE];
#pragma unroll
    for (int a = 0; a < NNE; ++a) nodes[a] = __ldg(mv.conn + e * NNE + a);
    double xe[NNE][DIM];
    load_coords<NNE>(mv.X, nodes, xe);
    for_each_q<NQ>([&](auto qc) {
        constexpr int Q = decltype(qc)::value;
#pragma unroll
        for (int l = 0; l < DIM; ++l) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < NNE; ++a) s = fma(Elem<NNE>::N(Q, a), xe[a][l], s);
            Xq[(e * NQ + Q) * DIM + l] = s;
        }
    });
}

// d(phi^T K psi) / d rho at the nodes, one thread per element (pyfem.py:1239-1276, 1872-1920 fused: the reference's
// (nelems, D, D, nnodes_per_elem) derivative tensor is never formed).  M = 1: Poisson, M = DIM: elasticity.
struct SensParams {
    Material mat;
    const double* phi;
    const double* psi;
    double c11, c12, c33;  // elasticity: C0 diagonal, off-diagonal and shear entries
};

// energy density phi-strain : C0 : psi-strain (times det^2) from the det-scaled gradients of the two fields
template <int M, int DIM>
PFG_DEV double sens_energy(const SensParams& prm, const double (&gu)[M][DIM], const double (&gv)[M][DIM]) {
    if constexpr (M == 1) {
        double energy = 0.0;
#pragma unroll
        for (int l = 0; l < DIM; ++l) energy = fma(gu[0][l], gv[0][l], energy);
        return energy;
    } else if constexpr (DIM == 2) {  // strains [ex, ey, gxy] (pyfem.py:1988-1998)
        const double gxu = gu[0][1] + gu[1][0], gxv = gv[0][1] + gv[1][0];
        return prm.c11 * (gu[0][0] * gv[0][0] + gu[1][1] * gv[1][1]) +
               prm.c12 * (gu[0][0] * gv[1][1] + gu[1][1] * gv[0][0]) + prm.c33 * gxu * gxv;
    } else {  // [ex, ey, ez, gxy, gyz, gxz] (pyfem.py:2000-2011)
        const double su = gu[0][0] + gu[1][1] + gu[2][2], sv = gv[0][0] + gv[1][1] + gv[2][2];
        double diag = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) diag = fma(gu[l][l], gv[l][l], diag);
        const double sh = (gu[0][1] + gu[1][0]) * (gv[0][1] + gv[1][0]) + (gu[1][2] + gu[2][1]) * (gv[1][2] + gv[2][1]) +
                          (gu[0][2] + gu[2][0]) * (gv[0][2] + gv[2][0]);
        return prm.c11 * diag + prm.c12 * (su * sv - diag) + prm.c33 * sh;
    }
}

// Nodal shares inner[o] = sum_q N[q, o] ramp'(rho_q) (strain(phi) : C0 : strain(psi)) detJ_q of one element
// (pyfem.py:1239-1276, 1872-1920).  hex8: gradients through G_a = det grad N_a.  quad4: the fields are bilinear like
// the map, so det * df/dx = y_eta f_xi - y_xi f_eta and det * df/dy = x_xi f_eta - x_eta f_xi with f_xi = a + c eta,
// f_eta = b + c xi -- six flops per field component and point, and no G_a at all (30 instead of 54 flops per point
// ahead of the energy for two components; 285 instead of 325 FP64 instructions per element, 78 instead of 128 registers).
template <int NNE, int M>
PFG_DEV void sens_element(const SensParams& prm, const double (&xe)[NNE][Elem<NNE>::DIM], const double (&re)[NNE],
                          const double (&ue)[NNE][M], const double (&ve)[NNE][M], double (&inner)[NNE]) {
    constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
#pragma unroll
    for (int a = 0; a < NNE; ++a) inner[a] = 0.0;
    const double scale = 1.0 + prm.mat.p;
    if constexpr (NNE == 4) {
        const Quad4Coef c = quad4_coef(xe);
        Quad4Field fu[M], fv[M];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            fu[k] = quad4_field4(ue[0][k], ue[1][k], ue[2][k], ue[3][k]);
            fv[k] = quad4_field4(ve[0][k], ve[1][k], ve[2][k], ve[3][k]);
        }
        const Quad4Field fr = quad4_field4(re[0], re[1], re[2], re[3]);
        const double scale16 = 0.0625 * scale;  // the two field gradients carry a factor 4 each
        for_each_q<4>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            constexpr double xi = Elem<4>::qp(Q, 0), eta = Elem<4>::qp(Q, 1);
            const double xxi = fma(c.cx, eta, c.ax), xeta = fma(c.cx, xi, c.bx);
            const double yxi = fma(c.cy, eta, c.ay), yeta = fma(c.cy, xi, c.by);
            const double det = xxi * yeta - xeta * yxi;
            const double rq = (prm.mat.rho != nullptr)
                                  ? fma(fr.c, 0.25 * xi * eta, fma(fr.b, 0.25 * eta, fma(fr.a, 0.25 * xi, 0.25 * fr.m)))
                                  : prm.mat.rho_const;
            const double den = fma(prm.mat.p, 1.0 - rq, 1.0);  // ramp'(rho_q) = (1 + p) / den^2, pyfem.py:1325
            double gu[M][2], gv[M][2];                          // det * gradients of the two fields
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const double uxi = fma(fu[k].c, eta, fu[k].a), ueta = fma(fu[k].c, xi, fu[k].b);
                const double vxi = fma(fv[k].c, eta, fv[k].a), veta = fma(fv[k].c, xi, fv[k].b);
                gu[k][0] = yeta * uxi - yxi * ueta, gu[k][1] = xxi * ueta - xeta * uxi;
                gv[k][0] = yeta * vxi - yxi * veta, gv[k][1] = xxi * veta - xeta * vxi;
            }
            // ramp' * (gu / det) . (gv / det) * det * w with w = 1: one reciprocal for both quotients
            const double t = scale16 * sens_energy<M, 2>(prm, gu, gv) * fast_rcp(den * den * det);
#pragma unroll
            for (int a = 0; a < 4; ++a) inner[a] = fma(Elem<4>::N(Q, a), t, inner[a]);
        });
    } else {
        // hex8, the same idea with trilinear fields: 8 f = m + cx xi + cy eta + cz zeta + cxy xi eta + cyz eta zeta +
        // cxz xi zeta + cxyz xi eta zeta (a three-stage butterfly over the corner values, 24 adds per field), so a
        // reference-space derivative is three FMAs, the Jacobian is the derivative of the coordinate fields (27 FMAs
        // per point instead of 72) and det * df/dx_l = sum_k df/dxi_k A[k][l] needs no G_a (18 flops per field
        // component and point instead of 24 + the 72 of G).  The factors 8 of the coefficient form are folded into one
        // constant: gradients carry 8 * 64 each, the determinant 512.
        Hex8Field fx[3], fu[M], fv[M];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            fx[j] = hex8_field8(xe[0][j], xe[1][j], xe[2][j], xe[3][j], xe[4][j], xe[5][j], xe[6][j], xe[7][j]);
#pragma unroll
        for (int k = 0; k < M; ++k) {
            fu[k] = hex8_field8(ue[0][k], ue[1][k], ue[2][k], ue[3][k], ue[4][k], ue[5][k], ue[6][k], ue[7][k]);
            fv[k] = hex8_field8(ve[0][k], ve[1][k], ve[2][k], ve[3][k], ve[4][k], ve[5][k], ve[6][k], ve[7][k]);
        }
        const double scale512 = scale * (1.0 / 512.0);
        for_each_q<NQ>([&](auto qc) {
            constexpr int Q = decltype(qc)::value;
            constexpr double xi = Elem<8>::qp(Q, 0), eta = Elem<8>::qp(Q, 1), zeta = Elem<8>::qp(Q, 2);
            double J[3][3];  // 8 * J[j][k] = 8 * d x_j / d xi_k
#pragma unroll
            for (int j = 0; j < 3; ++j) hex8_field_grad8<Q>(fx[j], J[j]);
            double A[3][3];  // adjugate (of 8 J): A = det * inv
            A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
            A[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
            A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
            A[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]);
            A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
            A[1][2] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]);
            A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
            A[2][1] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]);
            A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            const double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];  // 512 det J
            const double rq = (prm.mat.rho != nullptr) ? interp<NNE, Q>(re) : prm.mat.rho_const;
            const double den = fma(prm.mat.p, 1.0 - rq, 1.0);
            double gu[M][DIM], gv[M][DIM];
#pragma unroll
            for (int k = 0; k < M; ++k) {
                double du[3], dv[3];  // 8 * reference-space derivatives
                hex8_field_grad8<Q>(fu[k], du);
                hex8_field_grad8<Q>(fv[k], dv);
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    gu[k][l] = fma(du[2], A[2][l], fma(du[1], A[1][l], du[0] * A[0][l]));
                    gv[k][l] = fma(dv[2], A[2][l], fma(dv[1], A[1][l], dv[0] * A[0][l]));
                }
            }
            (void)xi, (void)eta, (void)zeta;
            const double t = scale512 * sens_energy<M, DIM>(prm, gu, gv) * fast_rcp(den * den * det);
#pragma unroll
            for (int a = 0; a < NNE; ++a) inner[a] = fma(Elem<NNE>::N(Q, a), t, inner[a]);
        });
    }
}

template <int NNE, int M>
__global__ void __launch_bounds__(128) k_dv_sens(MeshView mv, SensParams prm, double* __restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // (whole warps stay in the kernel for the shuffles below; a masked element is integrated by another rank)
    const bool active = e < mv.nelems && !(mv.elem_skip != nullptr && mv.elem_skip[e]);
    constexpr int DIM = Elem<NNE>::DIM;
    int nodes[NNE];
    double inner[NNE];
#pragma unroll
    for (int a = 0; a < NNE; ++a) nodes[a] = -1 - a, inner[a] = 0.0;
    if (active) {
#pragma unroll
        for (int a = 0; a < NNE; ++a) nodes[a] = __ldg(mv.conn + e * NNE + a);
        double xe[NNE][DIM], re[NNE], ue[NNE][M], ve[NNE][M];
        load_coords<NNE>(mv.X, nodes, xe);
        load_field<NNE>(prm.mat.rho, nodes, re);
#pragma unroll
        for (int a = 0; a < NNE; ++a) {
#pragma unroll
            for (int c = 0; c < M; ++c) {
                ue[a][c] = __ldg(prm.phi + (int64_t)nodes[a] * M + c);
                ve[a][c] = __ldg(prm.psi + (int64_t)nodes[a] * M + c);
            }
        }
        sens_element<NNE, M>(prm, xe, re, ue, ve, inner);
    }
    // Elements that follow each other along a mesh line share an edge (quad4) or a face (hex8): local nodes
    // (1, 2[, 5, 6]) of one are nodes (0, 3[, 4, 7]) of the next.  Where the ids say so (checked per lane pair, any mesh)
    // the next lane's share travels by shuffle and one atomic serves both: half the atomics on a lattice.
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NNE / 2; ++k) {
        const int mine = (k & 1) ? (k / 2) * 4 + 2 : (k / 2) * 4 + 1;    // 1, 2, 5, 6
        const int theirs = (k & 1) ? (k / 2) * 4 + 3 : (k / 2) * 4 + 0;  // 0, 3, 4, 7
        const int next_id = __shfl_down_sync(0xffffffffu, nodes[theirs], 1);
        const double next_val = __shfl_down_sync(0xffffffffu, inner[theirs], 1);
        const int prev_id = __shfl_up_sync(0xffffffffu, nodes[mine], 1);
        if (lane < 31 && next_id == nodes[mine]) inner[mine] += next_val;
        if (lane > 0 && prev_id == nodes[theirs]) nodes[theirs] = -1;  // taken by the previous lane
    }
#pragma unroll
    for (int a = 0; a < NNE; ++a)
        if (nodes[a] >= mv.own_begin && nodes[a] < mv.own_end) atomicAdd(out + (nodes[a] - mv.own_begin), inner[a]);
}

// The same sensitivities as a tile operator for scalar handles (one dof row per node): rho, phi and psi of the chunk's
// window nodes are staged once in shared memory, every element record leaves its NNE nodal shares in the staging
// area and phase B sums a node's shares in plan order -- no atomics, bitwise reproducible, every input read once.
template <int NNE_, int M_>
struct SensOp {
    static constexpr int NNE = NNE_, M = 1, NMAT = 0, NVEC = 1, MF = M_;
    static constexpr int DIM = Elem<NNE>::DIM, NQ = Elem<NNE>::NQ;
    static constexpr int FW = 1 + 2 * M_;  // per window node: rho, phi[MF], psi[MF]
    static constexpr bool NEEDS_ELEM = false, SYM = false;
    using Params = SensParams;
    __host__ __device__ __forceinline__ static const double* field(const Params& prm) { return prm.phi; }
    PFG_DEV static void gather_fields(const Params& prm, size_t node, double* dst) {
        if (prm.mat.rho != nullptr) cp_async_8(dst, prm.mat.rho + node);
#pragma unroll
        for (int c = 0; c < MF; ++c) {
            cp_async_8(dst + 1 + c, prm.phi + node * MF + c);
            cp_async_8(dst + 1 + MF + c, prm.psi + node * MF + c);
        }
    }
    template <class Sink>
    PFG_DEV static void run(const Params& prm, const double (&xe)[NNE][DIM], const double (&fe)[NNE * FW], int64_t,
                            Sink& sink) {
        double inner[NNE], re[NNE], ue[NNE][MF], ve[NNE][MF];
#pragma unroll
        for (int a = 0; a < NNE; ++a) {
            re[a] = fe[a * FW];
#pragma unroll
            for (int c = 0; c < MF; ++c) ue[a][c] = fe[a * FW + 1 + c], ve[a][c] = fe[a * FW + 1 + MF + c];
        }
        sens_element<NNE, MF>(prm, xe, re, ue, ve, inner);
#pragma unroll
        for (int a = 0; a < NNE; ++a) sink.vec(a, inner[a]);
    }
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static MeshView view_of(const MeshDev& d) {
    MeshView mv;
    mv.X = d.X;
    mv.own_begin = d.own_begin;
    mv.own_end = d.own_end;
    mv.conn = d.conn;
    mv.blk_ptr = d.blk_ptr;
    mv.rank = d.rank;
    mv.elem_skip = d.elem_skip;
    mv.nelems = d.nelems;
    mv.chunks = d.chunks;
    mv.cnodes = d.cnodes;
    mv.cnode_id = d.cnode_id;
    mv.rec_nodes = d.rec_nodes;
    mv.rec_dst = d.rec_dst;
    mv.rec_elem = d.rec_elem;
    mv.plan_pool = d.plan_pool;
    mv.tile_dir = d.tile_dir;
    mv.tile_blob = d.tile_blob;
    mv.tile_codes = d.tile_codes;
    mv.win_nodes = d.win_nodes;
    mv.rec_local = d.rec_local;
    mv.rec_skip = d.rec_skip;
    mv.stage_nodes_bytes = d.max_chunk_nodes * (int)sizeof(ChunkNode);
    mv.stage_plan_bytes = ((d.max_chunk_plan_words * 4 + 15) / 16) * 16 + 32;
    return mv;
}

static int resolve_mode(const MeshDev& d, int mode, bool* gather) {
    if (mode != PFG_MODE_AUTO && mode != PFG_MODE_ATOMIC && mode != PFG_MODE_GATHER) {
        set_error("unknown scatter mode %d", mode);
        return PFG_ERR_INVALID;
    }
    if (mode == PFG_MODE_GATHER && d.nchunks == 0) {
        set_error("this mesh has no gather plan (PFG_CREATE_NO_GATHER_PLAN, or node valence > %d)", kMaxValence);
        return PFG_ERR_UNSUPPORTED;
    }
    *gather = (mode == PFG_MODE_GATHER) || (mode == PFG_MODE_AUTO && d.nchunks > 0);
    return PFG_OK;
}

static int zero_outputs(const MeshDev& d, const Outputs& out, cudaStream_t st) {
    for (int i = 0; i < 2; ++i)
        if (out.vals[i]) PFG_CUDA_TRY(cudaMemsetAsync(out.vals[i], 0, d.nnz * sizeof(double), st));
    if (out.vec) PFG_CUDA_TRY(cudaMemsetAsync(out.vec, 0, (d.own_end - d.own_begin) * d.m * sizeof(double), st));
    return PFG_OK;
}


template <class Op, bool XD>
static size_t tile_smem_layout(const MeshDev& d, const typename Op::Params& prm, TileCfg& cfg) {
    using St = TileStage<Op>;
    constexpr int NNE = Op::NNE, DIM = Elem<NNE>::DIM;
    cfg.off_dir = 64;
    cfg.off_blob = cfg.off_dir + 8 * (int)sizeof(TileDir);
    cfg.off_codes = cfg.off_blob + align16(d.max_blob_bytes);
    cfg.off_win = cfg.off_codes + align16(d.max_code_bytes) + 16;  // slack: phase B may read a few codes past the end
    cfg.win_stride = align16(d.max_chunk_win * 4) + 16;            // copies start at the enclosing 16-byte boundary
    constexpr int NSTL = op_deep_prefetch<Op>::value ? 2 : 1;  // stages of the corner-index buffer
    constexpr int NSTX = (NSTL == 2 || XD) ? 2 : 1;            // stages of the window-data buffers
    cfg.off_loc = cfg.off_win + 2 * cfg.win_stride;
    cfg.loc_stride = align16(d.max_chunk_recs * NNE * 2) + 16;
    cfg.off_x = cfg.off_loc + NSTL * cfg.loc_stride;
    cfg.x_stride = align16(d.max_chunk_win * DIM * 8);
    cfg.off_field = cfg.off_x + NSTX * cfg.x_stride;
    cfg.field_stride = Op::field(prm) ? align16(d.max_chunk_win * 8 * op_field_width<Op>::value) : 0;
    cfg.off_stage = cfg.off_field + NSTX * cfg.field_stride;
    cfg.nchunks = (int)d.nchunks;
    cfg.off_image = align16(cfg.off_stage + (d.max_chunk_recs + 1) * St::S * 8);
    cfg.image_stride = align16(d.max_out_bytes) / 8;
    return (size_t)cfg.off_image + (size_t)Op::NMAT * cfg.image_stride * 8;
}

// attribute / occupancy queries cost more than a small assembly: repeated only when the configuration changes
template <class Op, int THREADS, int MINB, bool XD>
static int tile_resident_ctas(const MeshDev& d, size_t smem, int* per_sm) {
    static thread_local size_t cached_smem = 0;
    static thread_local int cached_device = -1, cached_per_sm = 0;
    if (cached_smem != smem || cached_device != d.device) {
        auto kern = k_tile<Op, THREADS, MINB, XD>;
        int q = 0;
        if (smem <= kMaxDynamicSmem) {
            PFG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PFG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PFG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, THREADS, smem));
        }
        cached_per_sm = q;
        cached_smem = smem;
        cached_device = d.device;
    }
    *per_sm = cached_per_sm;
    return PFG_OK;
}

template <class Op, int THREADS, int MINB, bool XD>
static int launch_tile_variant(const MeshDev& d, const MeshView& mv, const typename Op::Params& prm, const Outputs& out,
                               const TileCfg& cfg, size_t smem, int per_sm, cudaStream_t st) {
    using St = TileStage<Op>;
    const unsigned grid = (unsigned)std::min<int64_t>(d.nchunks, (int64_t)std::max(1, per_sm) * d.sm_count);
    static const bool debug = getenv("PFG_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr, "[pfg] k_tile%s: %d threads, %zu B smem (stage %d, image %d, blob %d, codes %d, win %d, recs<=%d), %d CTA/SM, grid %u\n",
                XD ? " [window data a chunk ahead]" : "", THREADS, smem, (d.max_chunk_recs + 1) * St::S * 8,
                Op::NMAT * cfg.image_stride * 8, d.max_blob_bytes, d.max_code_bytes, d.max_chunk_win, d.max_chunk_recs, per_sm,
                grid);
    k_tile<Op, THREADS, MINB, XD><<<grid, THREADS, smem, st>>>(mv, prm, out, cfg);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

template <class Op, int THREADS, int MINB>
static int launch_tile(MeshDev& d, const MeshView& mv, const typename Op::Params& prm, const Outputs& out,
                       cudaStream_t st) {
    using St = TileStage<Op>;
    PFG_TRY(tile_prepare_layout(d, St::layout(), st));
    if ((Op::M == 2) != (d.m == 2)) {
        set_error("tile plan was built for %d dofs per node, the operator has %d", d.m, Op::M);
        return PFG_ERR_INVALID;
    }
    // the image leaves through cp.async.bulk stores whose global addresses are formed from 16-byte aligned slots:
    // a values buffer that is itself misaligned (a tensor view at an odd offset, a packed inbox block) would fault
    for (int mt = 0; mt < 2; ++mt)
        if (Op::NMAT > mt && out.vals[mt] != nullptr && (reinterpret_cast<uintptr_t>(out.vals[mt]) & 15) != 0) {
            set_error("CSR values buffer %p is not 16-byte aligned (bulk stores need it)", (void*)out.vals[mt]);
            return PFG_ERR_INVALID;
        }
    TileCfg cfg;
    const size_t smem = tile_smem_layout<Op, false>(d, prm, cfg);
    if (smem > kMaxDynamicSmem) {
        set_error("chunk staging of %zu bytes exceeds shared memory", smem);
        return PFG_ERR_UNSUPPORTED;
    }
    int per_sm = 1;
    PFG_TRY((tile_resident_ctas<Op, THREADS, MINB, false>(d, smem, &per_sm)));
    if constexpr (op_deep_x<Op>::value) {
        // the variant that gathers the window data a whole chunk ahead, if its second stage costs no resident CTA
        static const bool no_xd = getenv("PFG_NO_DEEP_X") != nullptr;
        TileCfg cfg_x;
        const size_t smem_x = tile_smem_layout<Op, true>(d, prm, cfg_x);
        int per_sm_x = 0;
        PFG_TRY((tile_resident_ctas<Op, THREADS, MINB, true>(d, smem_x, &per_sm_x)));
        if (!no_xd && per_sm_x >= per_sm)
            return launch_tile_variant<Op, THREADS, MINB, true>(d, mv, prm, out, cfg_x, smem_x, per_sm_x, st);
    }
    return launch_tile_variant<Op, THREADS, MINB, false>(d, mv, prm, out, cfg, smem, per_sm, st);
}

template <class Op, int THREADS, int MINB>
static int launch(const MeshDev& d, const typename Op::Params& prm, const Outputs& out, bool gather,
                  cudaStream_t st) {
    const MeshView mv = view_of(d);
    if (!gather) {
        PFG_TRY(zero_outputs(d, out, st));
        const unsigned grid = (unsigned)((d.nelems + 127) / 128);
        k_assemble_atomic<Op><<<grid, 128, 0, st>>>(mv, prm, out);
    } else {
        if constexpr (Op::M <= 2) {
            return launch_tile<Op, 128, MINB>(const_cast<MeshDev&>(d), mv, prm, out, st);
        } else {
            set_error("no tile kernel for %d dofs per node", Op::M);
            return PFG_ERR_UNSUPPORTED;
        }
    }
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

static Material material_of(const double* rho_dev, double rho_const, double p) {
    Material m;
    m.rho = rho_dev;
    m.rho_const = rho_const;
    m.c_const = rho_const / (1.0 + p * (1.0 - rho_const));
    m.p = p;
    return m;
}

}  // namespace pfg

using namespace pfg;

#define PFG_CHECK_MESH(mesh)                         \
    if (!(mesh)) {                                   \
        set_error("%s: mesh is NULL", __func__);     \
        return PFG_ERR_INVALID;                      \
    }                                                \
    PFG_CUDA_TRY(cudaSetDevice((mesh)->d.device));

extern "C" int pfg_assemble_poisson(pfg_mesh* mesh, const double* rho_dev, double rho_const, double p,
                                    double* vals_dev, int mode, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != 1 || !vals_dev) {
        set_error("pfg_assemble_poisson: needs ndof_per_node == 1 and a values buffer");
        return PFG_ERR_INVALID;
    }
    bool gather;
    PFG_TRY(resolve_mode(d, mode, &gather));
    Outputs out{{vals_dev, nullptr}, nullptr};
    if (d.nne == 4) {
        PoissonOp<4>::Params prm{material_of(rho_dev, rho_const, p)};
        return launch<PoissonOp<4>, 128, 6>(d, prm, out, gather, (cudaStream_t)stream);
    }
    PoissonOp<8>::Params prm{material_of(rho_dev, rho_const, p)};
    return launch<PoissonOp<8>, 128, 2>(d, prm, out, gather, (cudaStream_t)stream);
}

extern "C" int pfg_assemble_helmholtz(pfg_mesh* mesh, double r0, double* K_vals_dev, double* R_vals_dev, int mode,
                                      void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != 1 || (!K_vals_dev && !R_vals_dev)) {
        set_error("pfg_assemble_helmholtz: needs ndof_per_node == 1 and at least one output");
        return PFG_ERR_INVALID;
    }
    bool gather;
    PFG_TRY(resolve_mode(d, mode, &gather));
    Outputs out{{K_vals_dev, R_vals_dev}, nullptr};
    if (d.nne == 4) {
        HelmholtzOp<4>::Params prm{r0 * r0};
        return launch<HelmholtzOp<4>, 128, 5>(d, prm, out, gather, (cudaStream_t)stream);
    }
    HelmholtzOp<8>::Params prm{r0 * r0};
    return launch<HelmholtzOp<8>, 128, 2>(d, prm, out, gather, (cudaStream_t)stream);
}

extern "C" int pfg_assemble_elasticity(pfg_mesh* mesh, const double* rho_dev, double rho_const, double p, double E,
                                       double nu, double* vals_dev, int mode, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != d.ndims || !vals_dev) {
        set_error("pfg_assemble_elasticity: needs ndof_per_node == ndims and a values buffer");
        return PFG_ERR_INVALID;
    }
    bool gather;
    PFG_TRY(resolve_mode(d, mode, &gather));
    cudaStream_t st = (cudaStream_t)stream;
    Outputs out{{vals_dev, nullptr}, nullptr};
    if (d.nne == 4) {
        // plane stress C0 (pyfem.py:1746-1750)
        const double f = E / (1.0 - nu * nu);
        ElasticityQuad4Op::Params prm{material_of(rho_dev, rho_const, p), f, f * nu, f * 0.5 * (1.0 - nu)};
        return launch<ElasticityQuad4Op, 128, 3>(d, prm, out, gather, st);
    }
    // 3-D C0 (pyfem.py:1752-1757)
    const double f = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
    ElasticityHex8Params prm{material_of(rho_dev, rho_const, p), f * (1.0 - nu), f * nu, f * (0.5 - nu)};
    const MeshView mv = view_of(d);
    const bool rows_ok = d.hex_rows_ok == 1 && d.nchunks > 0 && d.inc_rec8;
    if (mode != PFG_MODE_ATOMIC && rows_ok) {
        // owner-computes: geometry pass into the handle's scratch, then the chunk-row pass
        MeshDev& dm = const_cast<MeshDev&>(d);
        if (!dm.hex_geo) {
            PFG_CUDA_TRY(cudaMalloc(&dm.hex_geo, (size_t)d.nelems * 8 * kHexGeoDoubles * sizeof(double)));
            dm.device_bytes += d.nelems * 8 * kHexGeoDoubles * (int64_t)sizeof(double);
        }
        const HexRowsSmem L = hex_rows_smem(d.max_chunk_recs, d.max_k);  // fits: checked when the plan was built
        HexRowsCfg cfg;
        cfg.image_stride = L.image_stride;
        cfg.nchunks = (int)d.nchunks;
        cfg.off_geo = L.off_geo;
        cfg.geo_stage_bytes = L.geo_stage_bytes;
        cfg.off_image = L.off_image;
        cfg.off_meta = L.off_meta;
        const size_t smem = L.total;
        static thread_local size_t cached_smem = 0;
        static thread_local int cached_device = -1;  // function attributes are per device
        if (cached_smem != smem || cached_device != d.device) {
            PFG_CUDA_TRY(cudaFuncSetAttribute(k_hex8_chunk_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PFG_CUDA_TRY(cudaFuncSetAttribute(k_hex8_chunk_rows, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            cached_smem = smem;
            cached_device = d.device;
        }
        const unsigned grid = (unsigned)std::min<int64_t>(d.nchunks, d.sm_count);
        static const bool debug = getenv("PFG_DEBUG") != nullptr;
        if (debug)
            fprintf(stderr, "[pfg] k_hex8_chunk_rows: %zu B smem (records<=%d, nodes<=%d), %lld chunks, grid %u\n", smem,
                    d.max_chunk_recs, d.max_chunk_nodes, (long long)d.nchunks, grid);
        k_hex8_geometry<<<(unsigned)((d.nelems * 8 + 255) / 256), 256, 0, st>>>(mv, prm, dm.hex_geo);
        k_hex8_chunk_rows<<<grid, kHexRowsThreads, smem, st>>>(mv, prm, dm.hex_geo, d.inc_rec8,
                                                               reinterpret_cast<const uint2*>(d.inc_ranks8), vals_dev, cfg);
        PFG_CUDA_TRY(cudaGetLastError());
        return PFG_OK;
    }
    // meshes the chunk-row pass does not cover (a node with more than eight elements or 48 neighbours, or two
    // elements of a node that reach one neighbour through the same local index): the slot-indexed atomic scatter
    // (measured on B200, 128^3 hex: 6.4 ms) beats the first-format gather kernel (9.1 ms), so AUTO means atomic
    if (mode == PFG_MODE_AUTO) gather = false;
    if (gather && d.plan_pool == nullptr) {
        set_error("this hex8 handle has no plan for the staged-row-block gather kernel");
        return PFG_ERR_UNSUPPORTED;
    }
    if (!gather) {
        PFG_TRY(zero_outputs(d, out, st));
        const unsigned grid = (unsigned)((d.nelems + 15) / 16);
        k_elasticity_hex8_atomic<<<grid, 128, 0, st>>>(mv, prm, out);
    } else {
        constexpr int THREADS = 128;
        const size_t smem = 16 + mv.stage_nodes_bytes + mv.stage_plan_bytes +
                            ((size_t)(THREADS / 8) * kHexStageDoubles +
                             (size_t)d.max_chunk_inc * Layout<ElasticityHex8GatherOp>::RB) * sizeof(double);
        if (smem > 227 * 1024) {
            set_error("chunk staging of %zu bytes exceeds shared memory", smem);
            return PFG_ERR_UNSUPPORTED;
        }
        auto kern = k_elasticity_hex8_gather<THREADS>;
        PFG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)d.nchunks, THREADS, smem, st>>>(mv, prm, out);
    }
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_assemble_nlpoisson(pfg_mesh* mesh, const double* xdv_host, int nxdv, const double* u_dev,
                                      double* K_vals_dev, double* res_dev, int mode, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != 1 || d.nne != 4) {
        set_error("pfg_assemble_nlpoisson: quad4 with ndof_per_node == 1 only (NonlinearPoisson2D)");
        return PFG_ERR_UNSUPPORTED;
    }
    if (!xdv_host || nxdv < 1 || nxdv > kMaxXdv || !u_dev || (!K_vals_dev && !res_dev)) {
        set_error("pfg_assemble_nlpoisson: need 1..%d design variables, u and at least one output", kMaxXdv);
        return PFG_ERR_INVALID;
    }
    bool gather;
    PFG_TRY(resolve_mode(d, mode, &gather));
    NlPoissonQuad4Op::Params prm;
    prm.u = u_dev;
    prm.nxdv = nxdv;
    // binomial(nxdv-1, k) by the multiplicative recurrence (exact in double for nxdv <= 32)
    double binom = 1.0;
    for (int k = 0; k < kMaxXdv; ++k) {
        prm.coef[k] = (k < nxdv) ? xdv_host[k] * binom : 0.0;
        if (k < nxdv - 1) binom = binom * (double)(nxdv - 1 - k) / (double)(k + 1);
    }
    Outputs out{{K_vals_dev, nullptr}, res_dev};
    return launch<NlPoissonQuad4Op, 128, 3>(d, prm, out, gather, (cudaStream_t)stream);
}

// K(rho) for a COMPLEX nodal density (the reference's complex-step checks, pyfem.py:1018-1020 / 1783-1785 with
// tests/test_linear_poisson.py:57-89, tests/test_elasticity.py:68-104): the element matrices are real multiples of the
// complex RAMP factor, so the real and imaginary parts of K are two real assemblies, each with one part of the factor.
template <class Op>
static int launch_complex(const MeshDev& d, typename Op::Params prm, double* vals_re, double* vals_im, cudaStream_t st) {
    const MeshView mv = view_of(d);
    const unsigned grid = (unsigned)((d.nelems + 127) / 128);
    double* outs[2] = {vals_re, vals_im};
    for (int part = 0; part < 2; ++part) {
        if (!outs[part]) continue;
        Outputs out{{outs[part], nullptr}, nullptr};
        PFG_TRY(zero_outputs(d, out, st));
        prm.mat.part = part;
        k_assemble_atomic_complex<Op><<<grid, 128, 0, st>>>(mv, prm, out);
    }
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_assemble_poisson_complex(pfg_mesh* mesh, const double* rho_re_dev, const double* rho_im_dev, double p,
                                            double* vals_re_dev, double* vals_im_dev, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != 1 || !rho_re_dev || !rho_im_dev || (!vals_re_dev && !vals_im_dev)) {
        set_error("pfg_assemble_poisson_complex: needs ndof_per_node == 1, both parts of rho and an output");
        return PFG_ERR_INVALID;
    }
    Material mat = material_of(rho_re_dev, 0.0, p);
    mat.rho_im = rho_im_dev;
    if (d.nne == 4) return launch_complex<PoissonOp<4>>(d, {mat}, vals_re_dev, vals_im_dev, (cudaStream_t)stream);
    return launch_complex<PoissonOp<8>>(d, {mat}, vals_re_dev, vals_im_dev, (cudaStream_t)stream);
}

extern "C" int pfg_assemble_elasticity_complex(pfg_mesh* mesh, const double* rho_re_dev, const double* rho_im_dev,
                                               double p, double E, double nu, double* vals_re_dev, double* vals_im_dev,
                                               void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != d.ndims || !rho_re_dev || !rho_im_dev || (!vals_re_dev && !vals_im_dev)) {
        set_error("pfg_assemble_elasticity_complex: needs ndof_per_node == ndims, both parts of rho and an output");
        return PFG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Material mat = material_of(rho_re_dev, 0.0, p);
    mat.rho_im = rho_im_dev;
    if (d.nne == 4) {
        const double f = E / (1.0 - nu * nu);
        return launch_complex<ElasticityQuad4Op>(d, {mat, f, f * nu, f * 0.5 * (1.0 - nu)}, vals_re_dev, vals_im_dev, st);
    }
    const double f = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
    ElasticityHex8Params prm{mat, f * (1.0 - nu), f * nu, f * (0.5 - nu)};
    const MeshView mv = view_of(d);
    double* outs[2] = {vals_re_dev, vals_im_dev};
    for (int part = 0; part < 2; ++part) {
        if (!outs[part]) continue;
        Outputs out{{outs[part], nullptr}, nullptr};
        PFG_TRY(zero_outputs(d, out, st));
        prm.mat.part = part;
        k_elasticity_hex8_atomic<<<(unsigned)((d.nelems + 15) / 16), 128, 0, st>>>(mv, prm, out);
    }
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_quad_points(pfg_mesh* mesh, double* Xq_dev, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (!Xq_dev) {
        set_error("pfg_quad_points: output is NULL");
        return PFG_ERR_INVALID;
    }
    const MeshView mv = view_of(d);
    const unsigned grid = (unsigned)((d.nelems + 127) / 128);
    if (d.nne == 4) k_quad_points<4><<<grid, 128, 0, (cudaStream_t)stream>>>(mv, Xq_dev);
    else k_quad_points<8><<<grid, 128, 0, (cudaStream_t)stream>>>(mv, Xq_dev);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_poisson_rhs(pfg_mesh* mesh, const double* gq_dev, double* rhs_dev, int mode, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != 1 || !gq_dev || !rhs_dev) {
        set_error("pfg_poisson_rhs: needs ndof_per_node == 1, g at the quadrature points and an output");
        return PFG_ERR_INVALID;
    }
    bool gather;
    PFG_TRY(resolve_mode(d, mode, &gather));
    Outputs out{{nullptr, nullptr}, rhs_dev};
    if (d.nne == 4) {
        PoissonRhsOp<4>::Params prm{gq_dev};
        return launch<PoissonRhsOp<4>, 128, 4>(d, prm, out, gather, (cudaStream_t)stream);
    }
    PoissonRhsOp<8>::Params prm{gq_dev};
    return launch<PoissonRhsOp<8>, 128, 2>(d, prm, out, gather, (cudaStream_t)stream);
}

template <class Op>
static int launch_elements(const MeshDev& d, const typename Op::Params& prm, double* Ke0, double* Ke1, double* fe,
                           cudaStream_t st) {
    const MeshView mv = view_of(d);
    k_element_matrices<Op><<<(unsigned)((d.nelems + 127) / 128), 128, 0, st>>>(mv, prm, Ke0, Ke1, fe);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_element_matrices(pfg_mesh* mesh, int physics, const double* field_dev, double field_const,
                                    const double* params_host, int nparams, double* Ke_dev, double* Ke2_dev,
                                    double* fe_dev, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    cudaStream_t st = (cudaStream_t)stream;
    auto par = [&](int i, double dflt) { return (params_host && i < nparams) ? params_host[i] : dflt; };
    switch (physics) {
        case PFG_PHYS_POISSON: {
            if (d.m != 1 || !Ke_dev) break;
            if (d.nne == 4) {
                PoissonOp<4>::Params prm{material_of(field_dev, field_const, par(0, 0.0))};
                return launch_elements<PoissonOp<4>>(d, prm, Ke_dev, nullptr, nullptr, st);
            }
            PoissonOp<8>::Params prm{material_of(field_dev, field_const, par(0, 0.0))};
            return launch_elements<PoissonOp<8>>(d, prm, Ke_dev, nullptr, nullptr, st);
        }
        case PFG_PHYS_ELASTICITY: {
            if (d.m != d.ndims || !Ke_dev) break;
            const double p = par(0, 0.0), E = par(1, 10.0), nu = par(2, 0.3);
            if (d.nne == 4) {
                const double f = E / (1.0 - nu * nu);
                ElasticityQuad4Op::Params prm{material_of(field_dev, field_const, p), f, f * nu, f * 0.5 * (1.0 - nu)};
                return launch_elements<ElasticityQuad4Op>(d, prm, Ke_dev, nullptr, nullptr, st);
            }
            const double f = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
            ElasticityHex8Params prm{material_of(field_dev, field_const, p), f * (1.0 - nu), f * nu, f * (0.5 - nu)};
            k_elasticity_hex8_matrices<<<(unsigned)((d.nelems + 15) / 16), 128, 0, st>>>(view_of(d), prm, Ke_dev);
            PFG_CUDA_TRY(cudaGetLastError());
            return PFG_OK;
        }
        case PFG_PHYS_HELMHOLTZ: {
            if (d.m != 1 || (!Ke_dev && !Ke2_dev)) break;
            const double r0 = par(0, 0.0);
            if (d.nne == 4) {
                HelmholtzOp<4>::Params prm{r0 * r0};
                return launch_elements<HelmholtzOp<4>>(d, prm, Ke_dev, Ke2_dev, nullptr, st);
            }
            HelmholtzOp<8>::Params prm{r0 * r0};
            return launch_elements<HelmholtzOp<8>>(d, prm, Ke_dev, Ke2_dev, nullptr, st);
        }
        case PFG_PHYS_NLPOISSON: {
            if (d.m != 1 || d.nne != 4 || !field_dev || nparams < 1 || nparams > kMaxXdv || (!Ke_dev && !fe_dev)) break;
            NlPoissonQuad4Op::Params prm;
            prm.u = field_dev;
            prm.nxdv = nparams;
            double binom = 1.0;
            for (int k = 0; k < kMaxXdv; ++k) {
                prm.coef[k] = (k < nparams) ? params_host[k] * binom : 0.0;
                if (k < nparams - 1) binom = binom * (double)(nparams - 1 - k) / (double)(k + 1);
            }
            return launch_elements<NlPoissonQuad4Op>(d, prm, Ke_dev, nullptr, fe_dev, st);
        }
        default:
            set_error("pfg_element_matrices: unknown physics %d", physics);
            return PFG_ERR_INVALID;
    }
    set_error("pfg_element_matrices: physics %d does not fit this mesh handle or an output is missing", physics);
    return PFG_ERR_INVALID;
}

static int sens_params(const MeshDev& d, int physics, const double* rho_dev, double rho_const, double p,
                       const double* params_host, int nparams, const double* phi_dev, const double* psi_dev,
                       const double* out_dev, const char* who, SensParams* prm, int* mf) {
    if (!phi_dev || !psi_dev || !out_dev) {
        set_error("%s: NULL argument", who);
        return PFG_ERR_INVALID;
    }
    // elasticity: phi / psi carry ndims entries per node whatever the handle's own dof count
    if ((physics == PFG_PHYS_POISSON && d.m != 1) || (physics == PFG_PHYS_ELASTICITY && d.m != d.ndims && d.m != 1) ||
        (physics != PFG_PHYS_POISSON && physics != PFG_PHYS_ELASTICITY)) {
        set_error("%s: physics %d does not fit a handle with %d dofs per node", who, physics, d.m);
        return PFG_ERR_INVALID;
    }
    *mf = (physics == PFG_PHYS_POISSON) ? 1 : d.ndims;  // entries of phi / psi per node
    *prm = SensParams{material_of(rho_dev, rho_const, p), phi_dev, psi_dev, 0.0, 0.0, 0.0};
    if (physics == PFG_PHYS_ELASTICITY) {
        const double E = (params_host && nparams > 0) ? params_host[0] : 10.0;
        const double nu = (params_host && nparams > 1) ? params_host[1] : 0.3;
        if (d.ndims == 2) {  // plane stress (pyfem.py:1746-1750)
            const double f = E / (1.0 - nu * nu);
            prm->c11 = f, prm->c12 = f * nu, prm->c33 = f * 0.5 * (1.0 - nu);
        } else {  // pyfem.py:1752-1757
            const double f = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
            prm->c11 = f * (1.0 - nu), prm->c12 = f * nu, prm->c33 = f * (0.5 - nu);
        }
    }
    return PFG_OK;
}

extern "C" int pfg_k_dv_sens(pfg_mesh* mesh, int physics, const double* rho_dev, double rho_const, double p,
                             const double* params_host, int nparams, const double* phi_dev, const double* psi_dev,
                             double* out_dev, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    SensParams prm;
    int mf = 1;
    PFG_TRY(sens_params(d, physics, rho_dev, rho_const, p, params_host, nparams, phi_dev, psi_dev, out_dev,
                        "pfg_k_dv_sens", &prm, &mf));
    cudaStream_t st = (cudaStream_t)stream;
    const MeshView mv = view_of(d);
    PFG_CUDA_TRY(cudaMemsetAsync(out_dev, 0, (d.own_end - d.own_begin) * sizeof(double), st));
    const unsigned grid = (unsigned)((d.nelems + 127) / 128);
    if (d.nne == 4 && mf == 1) k_dv_sens<4, 1><<<grid, 128, 0, st>>>(mv, prm, out_dev);
    else if (d.nne == 4) k_dv_sens<4, 2><<<grid, 128, 0, st>>>(mv, prm, out_dev);
    else if (mf == 1) k_dv_sens<8, 1><<<grid, 128, 0, st>>>(mv, prm, out_dev);
    else k_dv_sens<8, 3><<<grid, 128, 0, st>>>(mv, prm, out_dev);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_k_dv_sens_ordered(pfg_mesh* mesh, int physics, const double* rho_dev, double rho_const, double p,
                                     const double* params_host, int nparams, const double* phi_dev,
                                     const double* psi_dev, double* out_dev, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    SensParams prm;
    int mf = 1;
    PFG_TRY(sens_params(d, physics, rho_dev, rho_const, p, params_host, nparams, phi_dev, psi_dev, out_dev,
                        "pfg_k_dv_sens_ordered", &prm, &mf));
    if (d.m != 1 || d.tile_dir == nullptr) {
        set_error("pfg_k_dv_sens_ordered: needs a scalar handle (ndof_per_node == 1) with a gather plan: its tile plan "
                  "carries the nodal-vector codes of the plan-ordered sums");
        return PFG_ERR_UNSUPPORTED;
    }
    const MeshView mv = view_of(d);
    Outputs out{{nullptr, nullptr}, out_dev};
    MeshDev& dm = const_cast<MeshDev&>(d);
    cudaStream_t st = (cudaStream_t)stream;
    if (d.nne == 4 && mf == 1) return launch_tile<SensOp<4, 1>, 128, 4>(dm, mv, prm, out, st);
    if (d.nne == 4) return launch_tile<SensOp<4, 2>, 128, 3>(dm, mv, prm, out, st);
    if (mf == 1) return launch_tile<SensOp<8, 1>, 128, 2>(dm, mv, prm, out, st);
    return launch_tile<SensOp<8, 3>, 128, 2>(dm, mv, prm, out, st);
}

extern "C" int pfg_scatter_matrix(pfg_mesh* mesh, const double* Ke_dev, double* vals_dev, int mode, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (!Ke_dev || !vals_dev) {
        set_error("pfg_scatter_matrix: NULL argument");
        return PFG_ERR_INVALID;
    }
    bool gather;
    PFG_TRY(resolve_mode(d, mode, &gather));
    cudaStream_t st = (cudaStream_t)stream;
    Outputs out{{vals_dev, nullptr}, nullptr};
    if (d.nne == 4 && d.m == 1) return launch<ScatterMatOp<4, 1>, 128, 4>(d, {Ke_dev}, out, gather, st);
    if (d.nne == 4 && d.m == 2) return launch<ScatterMatOp<4, 2>, 128, 2>(d, {Ke_dev}, out, gather, st);
    if (d.nne == 8 && d.m == 1) return launch<ScatterMatOp<8, 1>, 128, 2>(d, {Ke_dev}, out, gather, st);
    // hex8 with three dofs per node: slot-indexed atomic scatter (no tile plan for 24 x 24 element matrices)
    if (mode == PFG_MODE_GATHER) {
        set_error("pfg_scatter_matrix: hex8 elasticity handles scatter supplied matrices with atomics only");
        return PFG_ERR_UNSUPPORTED;
    }
    return launch<ScatterMatOp<8, 3>, 128, 1>(d, {Ke_dev}, out, false, st);
}

extern "C" int pfg_scatter_vector(pfg_mesh* mesh, const double* fe_dev, double* rhs_dev, int mode, void* stream) {
    PFG_CHECK_MESH(mesh);
    const MeshDev& d = mesh->d;
    if (d.m != 1 || !fe_dev || !rhs_dev) {
        set_error("pfg_scatter_vector: needs ndof_per_node == 1, element vectors and an output");
        return PFG_ERR_INVALID;
    }
    bool gather;
    PFG_TRY(resolve_mode(d, mode, &gather));
    Outputs out{{nullptr, nullptr}, rhs_dev};
    if (d.nne == 4) return launch<ScatterVecOp<4>, 128, 4>(d, {fe_dev}, out, gather, (cudaStream_t)stream);
    return launch<ScatterVecOp<8>, 128, 4>(d, {fe_dev}, out, gather, (cudaStream_t)stream);
}

__global__ void k_add_indexed(double* __restrict__ vals, const int64_t* __restrict__ idx, const double* __restrict__ src,
                              int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) vals[idx[i]] += src[i];
}

extern "C" int pfg_add_indexed(double* vals_dev, const int64_t* idx_dev, const double* src_dev, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!vals_dev || !idx_dev || !src_dev))) {
        set_error("pfg_add_indexed: invalid argument");
        return PFG_ERR_INVALID;
    }
    if (n) k_add_indexed<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(vals_dev, idx_dev, src_dev, n);
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}
