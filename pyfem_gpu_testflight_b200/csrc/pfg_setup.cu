// Once-per-mesh setup on the device: CSR pattern builder, element->slot rank map and the
// row-chunk gather plans.  Replaces ModelBase.__init__'s pattern work (reference pyfem.py:640-757,
// :837-858) and the sort/unique that scipy's coo->csr redoes on every assembly (pyfem.py:930-931).
//
// Everything here is integer work on device arrays; sorts and scans use CUB (the CUDA toolkit's
// header-only primitives), the rest are small hand-written kernels.
#include <cub/cub.cuh>
#include <stdarg.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "pfg_internal.cuh"

namespace pfg {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

// ---------------------------------------------------------------------------------------------
// small RAII device buffer + CUB temp storage
// ---------------------------------------------------------------------------------------------
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    }
    T* take() {
        T* r = p;
        p = nullptr;
        n = 0;
        return r;
    }
};

struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
    ~Scratch() {
        if (p) cudaFree(p);
    }
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
};

#define PFG_CUB(scratch, stream, call_with_temp)                                      \
    do {                                                                              \
        void* d_temp_storage = nullptr;                                               \
        size_t temp_storage_bytes = 0;                                                \
        PFG_CUDA_TRY(call_with_temp);                                                 \
        PFG_CUDA_TRY((scratch).reserve(temp_storage_bytes + 16));                     \
        d_temp_storage = (scratch).p;                                                 \
        PFG_CUDA_TRY(call_with_temp);                                                 \
    } while (0)

static inline int bits_for(uint64_t maxval) {
    int b = 1;
    while (b < 64 && (maxval >> b)) ++b;
    return b;
}

constexpr int kThreads = 256;
static inline unsigned grid_for(int64_t n, int threads = kThreads) {
    return (unsigned)std::max<int64_t>(1, (n + threads - 1) / threads);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_conn_to_i32(const int64_t* __restrict__ conn, int32_t* __restrict__ out, int64_t n,
                              long long* __restrict__ minmax) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    if (i < n) {
        long long v = conn[i];
        out[i] = (int32_t)v;
        lo = hi = v;
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
}

__global__ void k_iota_u32(uint32_t* out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

__global__ void k_count_nodes(const int32_t* __restrict__ conn, unsigned* __restrict__ counts, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[conn[i]], 1u);
}

// Sorted-unique neighbour list of one node in thread-local memory.
struct NbrList {
    int32_t v[kMaxRowBlocks + 1];
    int n = 0;
    bool overflow = false;
    __device__ void insert(int32_t c) {
        int lo = 0, hi = n;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (v[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (lo < n && v[lo] == c) return;
        if (n >= kMaxRowBlocks) { overflow = true; return; }
        for (int s = n; s > lo; --s) v[s] = v[s - 1];
        v[lo] = c;
        ++n;
    }
};

// pass 0: count neighbours of every owned node; pass 1: write them.
template <int NNE>
__global__ void k_node_neighbours(const int32_t* __restrict__ conn, const int64_t* __restrict__ inc_ptr,
                                  const uint32_t* __restrict__ inc_list, int64_t own_begin, int64_t nown,
                                  int* __restrict__ kcount, const int64_t* __restrict__ blk_ptr,
                                  int32_t* __restrict__ nbr, int* __restrict__ err) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nown) return;
    int64_t node = own_begin + r;
    NbrList L;
    for (int64_t s = inc_ptr[node]; s < inc_ptr[node + 1]; ++s) {
        int64_t e = inc_list[s] / NNE;
        const int32_t* c = conn + e * NNE;
#pragma unroll
        for (int b = 0; b < NNE; ++b) L.insert(c[b]);
    }
    if (L.overflow) atomicExch(err, 1);
    if (nbr == nullptr) {
        kcount[r] = L.n;
    } else {
        int64_t base = blk_ptr[r];
        for (int t = 0; t < L.n; ++t) nbr[base + t] = L.v[t];
    }
}

template <int NNE>
__global__ void k_rank_map(const int32_t* __restrict__ conn, const int64_t* __restrict__ blk_ptr,
                           const int32_t* __restrict__ nbr, int64_t own_begin, int64_t own_end, int64_t ninc,
                           uint8_t* __restrict__ rank) {
    int64_t ia = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // e*NNE + a
    if (ia >= ninc) return;
    int64_t e = ia / NNE;
    int32_t node = conn[ia];
    if (node < own_begin || node >= own_end) return;
    int64_t lo0 = blk_ptr[node - own_begin], hi0 = blk_ptr[node - own_begin + 1];
    const int32_t* c = conn + e * NNE;
#pragma unroll
    for (int b = 0; b < NNE; ++b) {
        int32_t col = c[b];
        int64_t lo = lo0, hi = hi0;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (nbr[mid] < col) lo = mid + 1; else hi = mid;
        }
        rank[ia * NNE + b] = (uint8_t)(lo - lo0);
    }
}

template <class IdxT>
__global__ void k_write_indptr(const int64_t* __restrict__ blk_ptr, int64_t nown, int m, IdxT* __restrict__ indptr) {
    int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // dof row, inclusive of the last+1
    int64_t nrows = nown * m;
    if (row > nrows) return;
    if (row == nrows) {
        indptr[row] = (IdxT)(blk_ptr[nown] * m * m);
        return;
    }
    int64_t r = row / m;
    int alpha = (int)(row - r * m);
    int64_t k = blk_ptr[r + 1] - blk_ptr[r];
    indptr[row] = (IdxT)(blk_ptr[r] * m * m + alpha * k * m);
}

template <class IdxT>
__global__ void k_write_indices(const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ nbr,
                                const int64_t* __restrict__ gid, int64_t nown, int m, IdxT* __restrict__ indices) {
    // one thread per (owned node, neighbour) block; writes m rows x m columns of column indices
    int64_t r = blockIdx.y * (int64_t)gridDim.x * blockDim.x + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // r indexes node blocks; find the row by binary search on blk_ptr
    int64_t nblocks = blk_ptr[nown];
    if (r >= nblocks) return;
    int64_t lo = 0, hi = nown;  // last row with blk_ptr[row] <= r
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (blk_ptr[mid] <= r) lo = mid; else hi = mid;
    }
    int64_t row = lo;
    int64_t k = blk_ptr[row + 1] - blk_ptr[row];
    int64_t t = r - blk_ptr[row];
    int64_t col = nbr[r];
    if (gid) col = gid[col];
    int64_t base = blk_ptr[row] * m * m;
    for (int alpha = 0; alpha < m; ++alpha)
        for (int beta = 0; beta < m; ++beta)
            indices[base + alpha * k * m + t * m + beta] = (IdxT)(col * m + beta);
}

// ---- chunking (sort-tile-recursive with tie-aware cuts) ---------------------------------------
__device__ __forceinline__ uint64_t sortable_f64(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x + 0.0);  // +0.0 folds -0 into +0
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_coord_keys(const double* __restrict__ X, int ndims, int axis, int64_t own_begin, int64_t nown,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nown) return;
    keys[r] = sortable_f64(X[(own_begin + r) * ndims + axis]);
    vals[r] = (uint32_t)r;
}

// which coordinate changes between consecutive node ids (counts[axis]), to find the direction node ids run along
__global__ void k_id_direction(const double* __restrict__ X, int ndims, int64_t own_begin, int64_t nown,
                               unsigned long long* __restrict__ counts) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r + 1 >= nown) return;
    int changed = 0, axis = -1;
    for (int l = 0; l < ndims; ++l)
        if (X[(own_begin + r) * ndims + l] != X[(own_begin + r + 1) * ndims + l]) ++changed, axis = l;
    if (changed == 1) atomicAdd(&counts[axis], 1ull);
}

__global__ void k_count_distinct(const uint64_t* __restrict__ sorted_keys, int64_t n, unsigned long long* __restrict__ count) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (r == 0 || sorted_keys[r] != sorted_keys[r - 1]) atomicAdd(count, 1ull);
}

__global__ void k_gather_u32(const uint32_t* __restrict__ table, const uint32_t* __restrict__ idx, int64_t n,
                             uint32_t* __restrict__ out) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) out[r] = table[idx[r]];
}

// Along the order (group, coord, id): mark the start position of each group and of each run of
// equal coordinate inside a group (as "r+1", 0 elsewhere) for the two max-scans.
__global__ void k_mark_starts(const uint32_t* __restrict__ group_sorted, const uint32_t* __restrict__ order,
                              const double* __restrict__ X, int ndims, int axis, int64_t own_begin, int64_t n,
                              int64_t* __restrict__ gstart, int64_t* __restrict__ vstart) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    bool newg = (r == 0) || (group_sorted[r] != group_sorted[r - 1]);
    bool newv = newg;
    if (!newv) {
        double a = X[(own_begin + order[r]) * ndims + axis];
        double b = X[(own_begin + order[r - 1]) * ndims + axis];
        newv = !(a == b);
    }
    gstart[r] = newg ? r + 1 : 0;
    vstart[r] = newv ? r + 1 : 0;
}

// cut id inside the group; then flag where (group, cut) changes.
__global__ void k_cut_flags(const uint32_t* __restrict__ group_sorted, const int64_t* __restrict__ gstart,
                            const int64_t* __restrict__ vstart, int64_t n, int64_t target,
                            uint32_t* __restrict__ cut, uint32_t* __restrict__ flag) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    auto cut_of = [&](int64_t q) -> uint32_t {
        int64_t g0 = gstart[q] - 1, v0 = vstart[q] - 1;
        int64_t pos = q - g0, vpos = v0 - g0;
        // keep runs of equal coordinate together unless the run itself is long
        int64_t basis = (pos - vpos >= (target + 3) / 4) ? pos : vpos;
        return (uint32_t)(basis / target);
    };
    uint32_t c = cut_of(r);
    cut[r] = c;
    bool f = (r == 0) || (group_sorted[r] != group_sorted[r - 1]) || (cut_of(r - 1) != c);
    flag[r] = f ? 1u : 0u;
}

__global__ void k_scatter_group(const uint32_t* __restrict__ order, const uint32_t* __restrict__ dense_incl,
                                int64_t n, uint32_t* __restrict__ group_of_node) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) group_of_node[order[r]] = dense_incl[r] - 1;
}

__global__ void k_chunk_by_id(int64_t n, int64_t target, uint32_t* __restrict__ group_of_node) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) group_of_node[r] = (uint32_t)(r / target);
}

// chunk-ordered node slots: per-slot valence / k, chunk boundaries
__global__ void k_slot_counts(const uint32_t* __restrict__ slot_node, const int64_t* __restrict__ inc_ptr,
                              const int64_t* __restrict__ blk_ptr, int64_t own_begin, int64_t n,
                              uint32_t* __restrict__ valence, uint32_t* __restrict__ kk) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    int64_t r = slot_node[p];
    valence[p] = (uint32_t)(inc_ptr[own_begin + r + 1] - inc_ptr[own_begin + r]);
    kk[p] = (uint32_t)(blk_ptr[r + 1] - blk_ptr[r]);
}

__global__ void k_chunk_node_begin(const uint32_t* __restrict__ slot_chunk, int64_t n, int64_t nchunks,
                                   ChunkHdr* __restrict__ chunks) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t c = slot_chunk[p];
    if (p == 0 || slot_chunk[p - 1] != c) chunks[c].node_begin = (uint32_t)p;
    if (p == n - 1 || slot_chunk[p + 1] != c) chunks[c].n_nodes = (uint32_t)(p + 1);  // fixed up below
}

__global__ void k_chunk_finish(int64_t nchunks, const int64_t* __restrict__ inc_excl, const uint32_t* __restrict__ kk,
                               const int64_t* __restrict__ plan_off, ChunkHdr* __restrict__ chunks, int64_t nslots,
                               int* __restrict__ maxima) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    ChunkHdr h = chunks[c];
    uint32_t end = h.n_nodes;  // holds end slot
    h.n_nodes = end - h.node_begin;
    int64_t inc0 = inc_excl[h.node_begin];
    int64_t inc1 = inc_excl[end];
    h.n_inc = (uint32_t)(inc1 - inc0);
    uint32_t kp = 0;
    for (uint32_t p = h.node_begin; p < end; ++p) kp = max(kp, kk[p]);
    h.kpad = kp;
    h.plan_begin = (uint32_t)plan_off[h.node_begin];
    h.plan_words = (uint32_t)(plan_off[end] - plan_off[h.node_begin]);
    chunks[c] = h;
    atomicMax(&maxima[4], (int)h.plan_words);
    atomicMax(&maxima[0], (int)h.n_inc);
    atomicMax(&maxima[1], (int)h.n_nodes);
    atomicMax(&maxima[2], (int)kp);
}

// one key per owned incidence: (chunk << 32) | element
template <int NNE>
__global__ void k_inc_keys(const uint32_t* __restrict__ slot_node, const uint32_t* __restrict__ slot_chunk,
                           const int64_t* __restrict__ inc_ptr, const uint32_t* __restrict__ inc_list,
                           const int64_t* __restrict__ inc_excl, int64_t own_begin, int64_t nslots,
                           uint64_t* __restrict__ keys) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    int64_t node = own_begin + slot_node[p];
    int64_t o = inc_excl[p];
    uint64_t c = slot_chunk[p];
    for (int64_t s = inc_ptr[node]; s < inc_ptr[node + 1]; ++s) keys[o++] = (c << 32) | (uint64_t)(inc_list[s] / NNE);
}

template <int NNE>
__global__ void k_fill_records(const uint64_t* __restrict__ rec_keys, int64_t nrecs, const int32_t* __restrict__ conn,
                               int32_t* __restrict__ rec_nodes, uint16_t* __restrict__ rec_dst,
                               int32_t* __restrict__ rec_elem, ChunkHdr* __restrict__ chunks, int* __restrict__ maxima) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nrecs) return;
    uint64_t key = rec_keys[r];
    uint32_t c = (uint32_t)(key >> 32);
    int64_t e = (int64_t)(key & 0xffffffffull);
    const int64_t at = r;
    rec_elem[at] = (int32_t)e;
#pragma unroll
    for (int a = 0; a < NNE; ++a) {
        rec_nodes[at * NNE + a] = conn[e * NNE + a];
        rec_dst[at * NNE + a] = kNoDst;
    }
    bool first = (r == 0) || ((uint32_t)(rec_keys[r - 1] >> 32) != c);
    bool last = (r == nrecs - 1) || ((uint32_t)(rec_keys[r + 1] >> 32) != c);
    if (first) chunks[c].rec_begin = r;
    if (last) chunks[c].n_recs = (uint32_t)(r + 1);  // end index (fits: checked on host), fixed up below
}

__global__ void k_chunk_rec_finish(int64_t nchunks, ChunkHdr* __restrict__ chunks, int* __restrict__ maxima) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    uint32_t n = (uint32_t)((int64_t)chunks[c].n_recs - chunks[c].rec_begin);
    chunks[c].n_recs = n;
    atomicMax(&maxima[3], (int)n);
}

template <int NNE>
__global__ void k_fill_dst(const uint32_t* __restrict__ slot_node, const uint32_t* __restrict__ slot_chunk,
                           const int64_t* __restrict__ inc_ptr, const uint32_t* __restrict__ inc_list,
                           const int64_t* __restrict__ inc_excl, const ChunkHdr* __restrict__ chunks,
                           const uint64_t* __restrict__ rec_keys, int64_t own_begin, int64_t nslots,
                           uint16_t* __restrict__ rec_dst) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    uint32_t c = slot_chunk[p];
    ChunkHdr h = chunks[c];
    int64_t node = own_begin + slot_node[p];
    int64_t inc_local = inc_excl[p] - inc_excl[h.node_begin];
    int j = 0;
    for (int64_t s = inc_ptr[node]; s < inc_ptr[node + 1]; ++s, ++j) {
        uint32_t ia = inc_list[s];
        uint64_t key = ((uint64_t)c << 32) | (uint64_t)(ia / NNE);
        int64_t lo = h.rec_begin, hi = h.rec_begin + h.n_recs;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (rec_keys[mid] < key) lo = mid + 1; else hi = mid;
        }
        rec_dst[lo * NNE + (ia % NNE)] = (uint16_t)(inc_local + j);
    }
}

// hex8 chunk-row pass: per chunk-ordered node slot and incidence, the element's record inside the chunk, the node's
// local index in it and the eight ranks of the element's nodes in the node's row (one load each in k_hex8_chunk_rows)
__global__ void k_fill_inc8(const uint32_t* __restrict__ slot_node, const uint32_t* __restrict__ slot_chunk,
                            const int64_t* __restrict__ inc_ptr, const uint32_t* __restrict__ inc_list,
                            const ChunkHdr* __restrict__ chunks, const uint64_t* __restrict__ rec_keys,
                            const uint8_t* __restrict__ rank, int64_t own_begin, int64_t nslots,
                            uint32_t* __restrict__ inc_rec8, uint64_t* __restrict__ inc_ranks8) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    const uint32_t c = slot_chunk[p];
    const ChunkHdr h = chunks[c];
    const int64_t node = own_begin + slot_node[p];
    int j = 0;
    for (int64_t s = inc_ptr[node]; s < inc_ptr[node + 1] && j < 8; ++s, ++j) {
        const uint32_t ia = inc_list[s];
        const uint64_t key = ((uint64_t)c << 32) | (uint64_t)(ia / 8);
        int64_t lo = h.rec_begin, hi = h.rec_begin + h.n_recs;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (rec_keys[mid] < key) lo = mid + 1; else hi = mid;
        }
        inc_rec8[p * 8 + j] = (uint32_t)((lo - h.rec_begin) * 8 + (ia % 8));
        inc_ranks8[p * 8 + j] = *reinterpret_cast<const uint64_t*>(rank + (size_t)ia * 8);
    }
    for (; j < 8; ++j) {
        inc_rec8[p * 8 + j] = 0xFFFF;
        inc_ranks8[p * 8 + j] = 0;
    }
}

__global__ void k_plan_sizes(const uint32_t* __restrict__ valence, const uint32_t* __restrict__ kk, int nne,
                             int64_t nslots, uint32_t* __restrict__ words) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    uint32_t bytes = kk[p] + 1 + valence[p] * nne;
    words[p] = (bytes + 3) / 4;
}

template <int NNE>
__global__ void k_fill_plans(const uint32_t* __restrict__ slot_node, const int64_t* __restrict__ inc_ptr,
                             const uint32_t* __restrict__ inc_list, const uint8_t* __restrict__ rank,
                             const int64_t* __restrict__ blk_ptr, const int64_t* __restrict__ inc_excl,
                             const ChunkHdr* __restrict__ chunks, const uint32_t* __restrict__ slot_chunk,
                             const int64_t* __restrict__ plan_off, int64_t own_begin, int64_t nslots, int m,
                             ChunkNode* __restrict__ cnodes, int32_t* __restrict__ cnode_id,
                             uint8_t* __restrict__ pool) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    int64_t r = slot_node[p];
    int64_t node = own_begin + r;
    int k = (int)(blk_ptr[r + 1] - blk_ptr[r]);
    int64_t s0 = inc_ptr[node], s1 = inc_ptr[node + 1];
    int val = (int)(s1 - s0);
    if (pool != nullptr) {  // (no plan records for handles served by the hex8 chunk-row pass: node table only)
        uint8_t* rec = pool + plan_off[p] * 4;
        uint8_t* start = rec;
        uint8_t* src = rec + k + 1;
        // counting sort of the valence*NNE contributions by neighbour rank
        uint8_t cnt[kMaxRowBlocks + 1];
        for (int t = 0; t <= k; ++t) cnt[t] = 0;
        for (int64_t s = s0; s < s1; ++s) {
            const uint8_t* rk = rank + (int64_t)inc_list[s] * NNE;
#pragma unroll
            for (int b = 0; b < NNE; ++b) cnt[rk[b] + 1]++;
        }
        for (int t = 0; t < k; ++t) cnt[t + 1] += cnt[t];
        for (int t = 0; t <= k; ++t) start[t] = cnt[t];
        int j = 0;
        for (int64_t s = s0; s < s1; ++s, ++j) {
            const uint8_t* rk = rank + (int64_t)inc_list[s] * NNE;
#pragma unroll
            for (int b = 0; b < NNE; ++b) src[cnt[rk[b]]++] = (uint8_t)((j << 3) | b);
        }
    }
    ChunkNode cn;
    cn.gslot = blk_ptr[r] * m * m;
    cn.plan = (uint32_t)plan_off[p];
    cn.inc_base = (uint16_t)(inc_excl[p] - inc_excl[chunks[slot_chunk[p]].node_begin]);
    cn.k = (uint8_t)k;
    cn.valence = (uint8_t)val;
    cnodes[p] = cn;
    cnode_id[p] = (int32_t)node;
}


// ---- tile plan (second format) -----------------------------------------------------------------
// per node: code groups (8 codes) of the matrix blocks and (scalar handles) groups of 4 vector codes
__global__ void k_tile_node_sizes(const uint32_t* __restrict__ valence, const uint32_t* __restrict__ slot_chunk, int nne,
                                  int64_t nslots, int* __restrict__ chunk_gmax, int* __restrict__ chunk_gvmax) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    const uint32_t n = valence[p] * nne;
    atomicMax(&chunk_gmax[slot_chunk[p]], (int)((n + 7) / 8));
    atomicMax(&chunk_gvmax[slot_chunk[p]], (int)((valence[p] + 3) / 4));
}

// a run starts where the chunk starts or the node ids stop being consecutive
__global__ void k_tile_run_flags(const uint32_t* __restrict__ slot_node, const uint32_t* __restrict__ slot_chunk,
                                 int64_t nslots, uint32_t* __restrict__ flag) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    flag[p] = (p == 0 || slot_chunk[p] != slot_chunk[p - 1] || slot_node[p] != slot_node[p - 1] + 1) ? 1u : 0u;
}

__global__ void k_tile_chunk_sizes(int64_t nchunks, const ChunkHdr* __restrict__ chunks,
                                   const uint32_t* __restrict__ run_id, const int* __restrict__ chunk_gmax,
                                   const int* __restrict__ chunk_gvmax, int m, uint32_t* __restrict__ blob_len16,
                                   uint32_t* __restrict__ code_len16, int* __restrict__ maxima) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const ChunkHdr h = chunks[c];
    const int64_t p0 = h.node_begin, p1 = p0 + h.n_nodes;
    const int64_t nruns = (int64_t)run_id[p1 - 1] - run_id[p0] + 1;
    const int64_t blob = tile_blob_tables(h.n_nodes, m) + (int64_t)sizeof(TileRun) * nruns;
    // group-major codes: [group][node] 8 matrix codes, then (scalar handles) [group][node] 4 vector codes
    const int64_t codes = 16 * (int64_t)chunk_gmax[c] * h.n_nodes + (m == 1 ? 8 * (int64_t)chunk_gvmax[c] * h.n_nodes : 0);
    blob_len16[c] = (uint32_t)((blob + 15) / 16);
    code_len16[c] = (uint32_t)((codes + 15) / 16);
    atomicMax(&maxima[5], (int)min((int64_t)INT_MAX, (blob + 15) / 16 * 16));
    atomicMax(&maxima[6], (int)min((int64_t)INT_MAX, (codes + 15) / 16 * 16));
}

__global__ void k_tile_dir(int64_t nchunks, const ChunkHdr* __restrict__ chunks, const int64_t* __restrict__ blob_off,
                           const int64_t* __restrict__ code_off, const uint32_t* __restrict__ blob_len16,
                           const uint32_t* __restrict__ code_len16, const uint32_t* __restrict__ win_begin,
                           const uint32_t* __restrict__ win_nodes_abs, const uint32_t* __restrict__ slot_node,
                           const int64_t* __restrict__ blk_ptr, int m, int64_t nwin, TileDir* __restrict__ dir,
                           int* __restrict__ err) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c > nchunks) return;
    TileDir t;
    memset(&t, 0, sizeof(t));
    if (c < nchunks) {  // entry nchunks is a zeroed sentinel (bulk copies of the directory may read it)
        const ChunkHdr h = chunks[c];
        const int64_t r_first = slot_node[h.node_begin];
        const int64_t wend = (c + 1 < nchunks) ? win_begin[c + 1] : nwin;
        t.gbase = blk_ptr[r_first] * m * m;
        t.blob_off16 = (uint32_t)blob_off[c];
        t.code_off16 = (uint32_t)code_off[c];
        t.loc_off = (uint32_t)h.rec_begin;
        t.win_off = win_begin[c];
        t.rec_begin = (uint32_t)h.rec_begin;
        t.node_base = win_nodes_abs[win_begin[c]];  // windows are sorted: the first id is the smallest
        t.row_base = (uint32_t)r_first;
        t.tmpl = (uint32_t)c;
        t.blob_len16 = (uint16_t)blob_len16[c];
        t.code_len16 = (uint16_t)code_len16[c];
        t.n_recs = (uint16_t)h.n_recs;
        t.n_win = (uint16_t)(wend - win_begin[c]);
        if (blob_len16[c] > 0xFFFFu || code_len16[c] > 0xFFFFu || h.n_recs > 0xFFFFu || wend - win_begin[c] > 0xFFFF)
            atomicExch(err, 1);
    }
    dir[c] = t;
}

// node ids of a window -> relative to the window's first (smallest) id
__global__ void k_win_relative(int64_t nchunks, const TileDir* __restrict__ dir, int64_t nwin,
                               const uint32_t* __restrict__ win_chunk, uint32_t* __restrict__ win_nodes) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nwin) return;
    win_nodes[i] -= dir[win_chunk[i]].node_base;
}

// ---- chunk templates: hash, group, verify -------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

struct TileTables {  // the four tables of a chunk as 32-bit words (all offsets / lengths are multiples of 4 bytes)
    const uint32_t* seg[4];
    uint32_t words[4];
};
__device__ __forceinline__ TileTables tile_tables(const TileDir& t, const uint8_t* blob, const uint16_t* codes_neutral,
                                                  const uint32_t* win, const uint16_t* loc, int nne) {
    TileTables T;
    T.seg[0] = reinterpret_cast<const uint32_t*>(blob + (size_t)t.blob_off16 * 16), T.words[0] = t.blob_len16 * 4u;
    T.seg[1] = reinterpret_cast<const uint32_t*>(codes_neutral + (size_t)t.code_off16 * 8), T.words[1] = t.code_len16 * 4u;
    T.seg[2] = win + t.win_off, T.words[2] = t.n_win;
    T.seg[3] = reinterpret_cast<const uint32_t*>(loc + (size_t)t.loc_off * nne), T.words[3] = t.n_recs * (uint32_t)nne / 2u;
    return T;
}

// one warp per chunk: position-aware 64-bit hash of its tables (sum of mixed words: lanes add in any order)
__global__ void k_tile_hash(int64_t nchunks, const TileDir* __restrict__ dir, const uint8_t* __restrict__ blob,
                            const uint16_t* __restrict__ codes_neutral, const uint32_t* __restrict__ win,
                            const uint16_t* __restrict__ loc, int nne, uint64_t* __restrict__ hash,
                            uint32_t* __restrict__ ids) {
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nchunks) return;
    const TileTables T = tile_tables(dir[c], blob, codes_neutral, win, loc, nne);
    uint64_t h = 0;
    for (int sgm = 0; sgm < 4; ++sgm) {
        for (uint32_t w = lane; w < T.words[sgm]; w += 32)
            h += mix64(((uint64_t)T.seg[sgm][w] << 32) | ((uint64_t)(sgm + 1) << 28) | w);
        h += (lane == 0) ? mix64(0x9e3779b97f4a7c15ull * (sgm + 1) + T.words[sgm]) : 0ull;
    }
    for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) hash[c] = h, ids[c] = (uint32_t)c;
}

// along the hash-sorted order: position of the first chunk of every run of equal hashes (0 elsewhere, for a max-scan)
__global__ void k_tile_group_heads(int64_t n, const uint64_t* __restrict__ sorted_hash, int64_t* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || sorted_hash[i] != sorted_hash[i - 1]) ? i : 0;
}

__global__ void k_tile_representative(int64_t n, const uint32_t* __restrict__ sorted_ids, const int64_t* __restrict__ head,
                                      uint32_t* __restrict__ rep) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) rep[sorted_ids[i]] = sorted_ids[head[i]];  // stable sort: the run's first entry is its smallest chunk id
}

// one warp per chunk: the chunk's tables must equal its representative's word for word (a hash collision clears `same`)
__global__ void k_tile_verify(int64_t nchunks, const TileDir* __restrict__ dir, const uint32_t* __restrict__ rep,
                              const uint8_t* __restrict__ blob, const uint16_t* __restrict__ codes_neutral,
                              const uint32_t* __restrict__ win, const uint16_t* __restrict__ loc, int nne,
                              int* __restrict__ mismatch) {
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nchunks || rep[c] == (uint32_t)c) return;
    const TileTables A = tile_tables(dir[c], blob, codes_neutral, win, loc, nne);
    const TileTables B = tile_tables(dir[rep[c]], blob, codes_neutral, win, loc, nne);
    bool bad = false;
    for (int sgm = 0; sgm < 4; ++sgm) {
        if (A.words[sgm] != B.words[sgm]) { bad = true; break; }
        for (uint32_t w = lane; w < A.words[sgm]; w += 32) bad |= A.seg[sgm][w] != B.seg[sgm][w];
    }
    if (bad) atomicExch(mismatch, 1);
}

// point every chunk at its template's tables; count templates and the bytes an assembly reads of the plan
__global__ void k_tile_apply_templates(int64_t nchunks, const uint32_t* __restrict__ rep, TileDir* __restrict__ dir,
                                       const TileDir* __restrict__ dir_in, int nne, unsigned long long* __restrict__ stats) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const uint32_t r = rep[c];
    TileDir t = dir_in[c];
    if (r == (uint32_t)c) {
        atomicAdd(&stats[0], 1ull);
        atomicAdd(&stats[1], (unsigned long long)(t.blob_len16 * 16u + t.code_len16 * 16u + t.n_win * 4u + t.n_recs * nne * 2u));
    } else {
        const TileDir tr = dir_in[r];
        t.blob_off16 = tr.blob_off16, t.code_off16 = tr.code_off16, t.loc_off = tr.loc_off, t.win_off = tr.win_off;
        t.tmpl = r;
    }
    dir[c] = t;
}

// ---- node windows: sorted unique node ids per chunk, and every record corner's index into its window
template <int NNE>
__global__ void k_win_keys(const uint64_t* __restrict__ rec_keys, const int32_t* __restrict__ rec_nodes, int64_t ncorners,
                           uint64_t* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ncorners) return;
    keys[i] = (rec_keys[i / NNE] & 0xffffffff00000000ull) | (uint32_t)rec_nodes[i];
}

__global__ void k_win_fill(const uint64_t* __restrict__ win_keys, int64_t nwin, uint32_t* __restrict__ win_nodes,
                           uint32_t* __restrict__ win_begin, uint32_t* __restrict__ win_chunk) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nwin) return;
    const uint64_t k = win_keys[i];
    win_nodes[i] = (uint32_t)k;
    win_chunk[i] = (uint32_t)(k >> 32);
    if (i == 0 || (win_keys[i - 1] >> 32) != (k >> 32)) win_begin[k >> 32] = (uint32_t)i;
}

__global__ void k_win_max(int64_t nchunks, const uint32_t* __restrict__ win_begin, int64_t nwin, int* __restrict__ maxima) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const int64_t e = (c + 1 < nchunks) ? win_begin[c + 1] : nwin;
    atomicMax(&maxima[8], (int)(e - win_begin[c]));
}

template <int NNE>
__global__ void k_rec_local(const uint64_t* __restrict__ rec_keys, const int32_t* __restrict__ rec_nodes,
                            const uint64_t* __restrict__ win_keys, const uint32_t* __restrict__ win_begin, int64_t nchunks,
                            int64_t nwin, int64_t ncorners, uint16_t* __restrict__ rec_local, int* __restrict__ err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ncorners) return;
    const uint64_t ck = rec_keys[i / NNE] & 0xffffffff00000000ull;
    const uint64_t c = ck >> 32;
    const uint64_t key = ck | (uint32_t)rec_nodes[i];
    int64_t lo = win_begin[c], hi = ((int64_t)c + 1 < nchunks) ? win_begin[c + 1] : nwin;
    const int64_t base = lo;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (win_keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo - base > 0xFFFF) atomicExch(err, 1);
    rec_local[i] = (uint16_t)(lo - base);
}

struct TileFillArgs {
    const uint32_t *slot_node, *slot_chunk;
    const int64_t* inc_ptr;
    const uint32_t* inc_list;
    const uint8_t* rank;
    const int64_t* blk_ptr;
    const ChunkHdr* chunks;
    const TileDir* dir;
    const uint64_t* rec_keys;
    const uint32_t *run_flag, *run_id;
    int64_t own_begin, nslots;
    int m;
    uint8_t* blob_pool;
    uint16_t* codes_neutral;
    int32_t* cnode_id;
    int* err;
    const int *chunk_gmax, *chunk_gvmax;
    int* maxima;
};

// neutral code: end << 15 | vec << 14 | record << 2*LB | a << LB | b; 0xFFFF = padding
template <int NNE>
__global__ void k_tile_fill(TileFillArgs A) {
    constexpr int LB = (NNE == 4) ? 2 : 3;
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= A.nslots) return;
    const uint32_t c = A.slot_chunk[p];
    const ChunkHdr h = A.chunks[c];
    const TileDir td = A.dir[c];
    const int64_t r = A.slot_node[p];
    const int64_t node = A.own_begin + r;
    const int64_t r_first = A.slot_node[h.node_begin];
    const int64_t p_end = (int64_t)h.node_begin + h.n_nodes;
    uint8_t* blob = A.blob_pool + (size_t)td.blob_off16 * 16;
    const int k = (int)(A.blk_ptr[r + 1] - A.blk_ptr[r]);
    const int64_t s0 = A.inc_ptr[node], s1 = A.inc_ptr[node + 1];
    const int64_t grel = (A.blk_ptr[r] - A.blk_ptr[r_first]) * A.m * A.m;
    const int64_t nruns = (int64_t)A.run_id[p_end - 1] - A.run_id[h.node_begin] + 1;
    const int64_t pl = p - h.node_begin;
    const int gmax = A.chunk_gmax[c], gvmax = A.chunk_gvmax[c];
    if (grel < 0 || grel > 0xFFFFFFFFll || h.n_recs >= (0x3FFFu >> (2 * LB)) || gmax > 0xFFFF || gvmax > 0xFFFF ||
        nruns > 128 || k > 0xFFFF) {
        atomicExch(A.err, 1);
        return;
    }
    if (p == h.node_begin) {
        TileHdr th;
        th.n_nodes = (uint16_t)h.n_nodes;
        th.n_recs = (uint16_t)h.n_recs;
        th.gmax = (uint16_t)gmax;
        th.gvmax = (uint16_t)gvmax;
        th.n_runs = (uint16_t)nruns;
        th.pad_[0] = th.pad_[1] = th.pad_[2] = 0;
        *reinterpret_cast<TileHdr*>(blob) = th;
    }
    TileNode* tn = reinterpret_cast<TileNode*>(blob + sizeof(TileHdr)) + pl;
    tn->gslot_rel = (uint32_t)grel;  // aux (image offset) is filled by k_tile_image_layout
    tn->k = (uint16_t)k;
    if (A.m == 1)
        reinterpret_cast<uint32_t*>(blob + sizeof(TileHdr) + sizeof(TileNode) * h.n_nodes)[pl] = (uint32_t)(r - r_first);
    // counting sort of the valence*NNE contributions by neighbour rank
    uint8_t cnt[kMaxRowBlocks + 1];
    for (int t = 0; t <= k; ++t) cnt[t] = 0;
    for (int64_t s = s0; s < s1; ++s) {
        const uint8_t* rk = A.rank + (int64_t)A.inc_list[s] * NNE;
#pragma unroll
        for (int b = 0; b < NNE; ++b) cnt[rk[b] + 1]++;
    }
    for (int t = 0; t < k; ++t) cnt[t + 1] += cnt[t];
    // matrix codes in block order, group-major: code s of this node sits at [s / 8][node][s % 8]; the last
    // contribution of each block carries the end flag.  Vector codes (scalar handles): [j / 4][node][j % 4].
    uint16_t* codes = A.codes_neutral + (size_t)td.code_off16 * 8;
    uint16_t* vcodes = codes + (size_t)gmax * h.n_nodes * 8;
    uint8_t end_at[kMaxRowBlocks + 1];  // last code index of block t = start of block t+1 minus one
    for (int t = 0; t < k; ++t) end_at[t] = (uint8_t)(cnt[t + 1] - 1);
    int j = 0;
    for (int64_t s = s0; s < s1; ++s, ++j) {
        const uint32_t ia = A.inc_list[s];
        const uint64_t key = ((uint64_t)c << 32) | (uint64_t)(ia / NNE);
        int64_t lo = h.rec_begin, hi = h.rec_begin + h.n_recs;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (A.rec_keys[mid] < key) lo = mid + 1; else hi = mid;
        }
        const uint32_t rloc = (uint32_t)(lo - h.rec_begin);
        const uint32_t a = ia % NNE;
        const uint8_t* rk = A.rank + (int64_t)ia * NNE;
#pragma unroll
        for (int b = 0; b < NNE; ++b) {
            const int t = rk[b];
            const int pos = cnt[t]++;
            const uint32_t end = (pos == end_at[t]) ? 0x8000u : 0u;
            codes[((int64_t)(pos >> 3) * h.n_nodes + pl) * 8 + (pos & 7)] = (uint16_t)(end | (rloc << (2 * LB)) | (a << LB) | b);
        }
        if (A.m == 1) {
            const uint32_t end = (s + 1 == s1) ? 0x8000u : 0u;
            vcodes[((int64_t)(j >> 2) * h.n_nodes + pl) * 4 + (j & 3)] = (uint16_t)(end | 0x4000u | (rloc << (2 * LB)) | (a << LB) | a);
        }
    }
    A.cnode_id[p] = (int32_t)node;
}

// one thread per chunk: place the nodes' rows in the chunk's CSR image and list the runs of consecutive node ids.
// Units are 16 bytes (2 dofs per node) or 8 bytes (scalar); a scalar run starts on the same 16-byte phase in the
// image as in the CSR values, so that its aligned middle part can leave as one bulk store.
__global__ void k_tile_image_layout(int64_t nchunks, const ChunkHdr* __restrict__ chunks, const TileDir* __restrict__ dir,
                                    const uint32_t* __restrict__ slot_node, const int64_t* __restrict__ blk_ptr,
                                    const uint32_t* __restrict__ run_flag, int m, uint8_t* __restrict__ blob_pool,
                                    int* __restrict__ maxima, int* __restrict__ err) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const ChunkHdr h = chunks[c];
    uint8_t* blob = blob_pool + (size_t)dir[c].blob_off16 * 16;
    TileNode* nodes = reinterpret_cast<TileNode*>(blob + sizeof(TileHdr));
    TileRun* runs = reinterpret_cast<TileRun*>(blob + tile_blob_tables(h.n_nodes, m));
    const int64_t gbase = dir[c].gbase;
    const int unit_doubles = (m == 2) ? 2 : 1;
    int64_t off = 0;  // image offset in units
    int nrun = -1;
    int64_t run_start = 0;
    for (int64_t q = 0; q < (int64_t)h.n_nodes; ++q) {
        const int64_t p = h.node_begin + q;
        const int64_t rq = slot_node[p];
        const int64_t len = (blk_ptr[rq + 1] - blk_ptr[rq]) * m * m / unit_doubles;
        if (run_flag[p]) {
            if (nrun >= 0) runs[nrun].len = (uint16_t)(off - run_start);
            if (unit_doubles == 1 && (((gbase + nodes[q].gslot_rel) ^ off) & 1)) ++off;  // same 16-byte phase
            ++nrun;
            run_start = off;
            runs[nrun].gslot_rel = nodes[q].gslot_rel;
            runs[nrun].out_off = (uint16_t)off;
        }
        nodes[q].aux = (uint16_t)off;
        off += len;
        if (off > 0xFFFF) {
            atomicExch(err, 1);
            return;
        }
    }
    if (nrun >= 0) runs[nrun].len = (uint16_t)(off - run_start);
    atomicMax(&maxima[7], (int)((off + 1) * unit_doubles * 8));
}

// ---- compaction: only the tables of the templates (the representative chunks) are kept -----------------------
// sizes of a chunk's four tables if it is a template, zero otherwise (inputs of four exclusive scans)
__global__ void k_tile_template_sizes(int64_t nchunks, const TileDir* __restrict__ dir, int nne, int64_t* __restrict__ blob16,
                                      int64_t* __restrict__ code16, int64_t* __restrict__ win, int64_t* __restrict__ recs) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c > nchunks) return;
    const bool keep = c < nchunks && dir[c].tmpl == (uint32_t)c;
    blob16[c] = keep ? dir[c].blob_len16 : 0;
    code16[c] = keep ? dir[c].code_len16 : 0;
    win[c] = keep ? dir[c].n_win : 0;  // (windows / corner tables may start anywhere: the kernel fetches them from the
    recs[c] = keep ? dir[c].n_recs : 0;  // enclosing 16-byte boundary)
}

// one warp per template: copy its tables to their compact places
__global__ void k_tile_compact_copy(int64_t nchunks, const TileDir* __restrict__ dir, int nne, const int64_t* __restrict__ blob16,
                                    const int64_t* __restrict__ code16, const int64_t* __restrict__ win,
                                    const int64_t* __restrict__ recs, const uint8_t* __restrict__ blob_in,
                                    const uint16_t* __restrict__ codes_in, const uint32_t* __restrict__ win_in,
                                    const uint16_t* __restrict__ loc_in, uint8_t* __restrict__ blob_out,
                                    uint16_t* __restrict__ codes_out, uint32_t* __restrict__ win_out,
                                    uint16_t* __restrict__ loc_out) {
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nchunks || dir[c].tmpl != (uint32_t)c) return;
    const TileDir t = dir[c];
    const uint32_t* bi = reinterpret_cast<const uint32_t*>(blob_in + (size_t)t.blob_off16 * 16);
    uint32_t* bo = reinterpret_cast<uint32_t*>(blob_out + (size_t)blob16[c] * 16);
    for (uint32_t w = lane; w < t.blob_len16 * 4u; w += 32) bo[w] = bi[w];
    const uint32_t* ci = reinterpret_cast<const uint32_t*>(codes_in + (size_t)t.code_off16 * 8);
    uint32_t* co = reinterpret_cast<uint32_t*>(codes_out + (size_t)code16[c] * 8);
    for (uint32_t w = lane; w < t.code_len16 * 4u; w += 32) co[w] = ci[w];
    for (uint32_t w = lane; w < t.n_win; w += 32) win_out[win[c] + w] = win_in[t.win_off + w];
    const uint16_t* li = loc_in + (size_t)t.loc_off * nne;
    uint16_t* lo = loc_out + (size_t)recs[c] * nne;
    for (uint32_t w = lane; w < t.n_recs * (uint32_t)nne; w += 32) lo[w] = li[w];
}

__global__ void k_tile_compact_dir(int64_t nchunks, const int64_t* __restrict__ blob16, const int64_t* __restrict__ code16,
                                   const int64_t* __restrict__ win, const int64_t* __restrict__ recs,
                                   TileDir* __restrict__ dir) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const uint32_t r = dir[c].tmpl;
    dir[c].blob_off16 = (uint32_t)blob16[r];
    dir[c].code_off16 = (uint32_t)code16[r];
    dir[c].win_off = (uint32_t)win[r];
    dir[c].loc_off = (uint32_t)recs[r];
}

// neutral codes -> staging offsets of one operator layout (record slot 0 is the zero record; padding -> 0)
__host__ __device__ inline uint32_t tile_encode(const TileLayout& L, uint32_t neutral) {
    if (neutral == 0xFFFFu) return 0;
    const int lb = (L.nne == 4) ? 2 : 3;
    const uint32_t end = (neutral >> 15) << 1;
    const bool vec = (neutral >> 14) & 1u;
    const uint32_t b = neutral & (L.nne - 1), a = (neutral >> lb) & (L.nne - 1), r = ((neutral & 0x3FFFu) >> (2 * lb)) + 1;
    if (vec) return L.vec_units >= 0 ? (((r * L.rec_units + L.vec_units + a) << 2) | end) : 0;
    if (!L.has_mat) return 0;
    if (!L.sym) return ((r * L.rec_units + (a * L.nne + b) * L.blk_units) << 2) | end;
    const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
    const uint32_t tri = lo * (2 * L.nne - 1 - lo) / 2 + hi;
    return ((r * L.rec_units + tri * L.blk_units) << 2) | end | (a > b ? 1u : 0u);
}

__global__ void k_tile_encode(const uint16_t* __restrict__ neutral, uint16_t* __restrict__ out, int64_t n, TileLayout L) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint16_t)tile_encode(L, neutral[i]);
}

// (re-)encode the working codes for an operator layout; no-op when they already match
int tile_prepare_layout(MeshDev& d, const TileLayout& L, cudaStream_t st) {
    if (d.tile_layout == L) return PFG_OK;
    const int64_t max_off = (int64_t)(d.max_chunk_recs + 1) * L.rec_units;
    if (max_off >= 16384) {
        set_error("chunk staging of %lld units exceeds the 14-bit code range", (long long)max_off);
        return PFG_ERR_UNSUPPORTED;
    }
    if (d.tile_ncodes)
        k_tile_encode<<<grid_for(d.tile_ncodes), kThreads, 0, st>>>(d.tile_codes_neutral, d.tile_codes, d.tile_ncodes, L);
    PFG_CUDA_TRY(cudaGetLastError());
    d.tile_layout = L;
    return PFG_OK;
}

// ---------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------
template <int NNE>
static int build_pattern(MeshDev& d, cudaStream_t st, Scratch& scratch) {
    const int64_t ninc = d.nelems * NNE;
    const int64_t nown = d.own_end - d.own_begin;

    // node -> incidence lists (stable radix sort of (node, e*NNE+a))
    DBuf<uint32_t> vals_in, keys_out;
    DBuf<unsigned> counts;
    PFG_CUDA_TRY(vals_in.alloc(ninc));
    PFG_CUDA_TRY(keys_out.alloc(ninc));
    PFG_CUDA_TRY(counts.alloc(d.nnodes + 1));
    PFG_CUDA_TRY(cudaMalloc(&d.inc_list, std::max<int64_t>(ninc, 1) * sizeof(uint32_t)));
    PFG_CUDA_TRY(cudaMalloc(&d.inc_ptr, (d.nnodes + 1) * sizeof(int64_t)));
    k_iota_u32<<<grid_for(ninc), kThreads, 0, st>>>(vals_in.p, ninc);
    PFG_CUB(scratch, st,
            cub::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, (const uint32_t*)d.conn, keys_out.p,
                                            (const uint32_t*)vals_in.p, d.inc_list, ninc, 0,
                                            bits_for((uint64_t)d.nnodes), st));
    PFG_CUDA_TRY(cudaMemsetAsync(counts.p, 0, (d.nnodes + 1) * sizeof(unsigned), st));
    k_count_nodes<<<grid_for(ninc), kThreads, 0, st>>>(d.conn, counts.p, ninc);
    PFG_CUB(scratch, st,
            cub::DeviceScan::ExclusiveSum(d_temp_storage, temp_storage_bytes, counts.p, d.inc_ptr, d.nnodes + 1, st));
    DBuf<unsigned> maxval;
    PFG_CUDA_TRY(maxval.alloc(1));
    PFG_CUB(scratch, st,
            cub::DeviceReduce::Max(d_temp_storage, temp_storage_bytes, counts.p, maxval.p, d.nnodes, st));
    unsigned h_maxval = 0;
    PFG_CUDA_TRY(cudaMemcpyAsync(&h_maxval, maxval.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));

    // neighbour counts -> blk_ptr -> neighbour lists
    DBuf<int> kcount, err;
    PFG_CUDA_TRY(kcount.alloc(nown + 1));
    PFG_CUDA_TRY(err.alloc(1));
    PFG_CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), st));
    PFG_CUDA_TRY(cudaMemsetAsync(kcount.p, 0, (nown + 1) * sizeof(int), st));
    PFG_CUDA_TRY(cudaMalloc(&d.blk_ptr, (nown + 1) * sizeof(int64_t)));
    k_node_neighbours<NNE><<<grid_for(nown, 128), 128, 0, st>>>(d.conn, d.inc_ptr, d.inc_list, d.own_begin, nown,
                                                               kcount.p, nullptr, nullptr, err.p);
    PFG_CUB(scratch, st,
            cub::DeviceScan::ExclusiveSum(d_temp_storage, temp_storage_bytes, kcount.p, d.blk_ptr, nown + 1, st));
    DBuf<int> maxk;
    PFG_CUDA_TRY(maxk.alloc(1));
    PFG_CUB(scratch, st, cub::DeviceReduce::Max(d_temp_storage, temp_storage_bytes, kcount.p, maxk.p, nown, st));
    int h_err = 0, h_maxk = 0;
    PFG_CUDA_TRY(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaMemcpyAsync(&h_maxk, maxk.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaMemcpyAsync(&d.nblocks, d.blk_ptr + nown, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_err) {
        set_error("a node has more than %d neighbour nodes; mesh valence too high for the device pattern builder",
                  kMaxRowBlocks);
        return PFG_ERR_UNSUPPORTED;
    }
    d.max_k = h_maxk;
    d.max_valence = (int)h_maxval;
    d.nnz = d.nblocks * d.m * d.m;
    PFG_CUDA_TRY(cudaMalloc(&d.nbr, std::max<int64_t>(d.nblocks, 1) * sizeof(int32_t)));
    k_node_neighbours<NNE><<<grid_for(nown, 128), 128, 0, st>>>(d.conn, d.inc_ptr, d.inc_list, d.own_begin, nown,
                                                               nullptr, d.blk_ptr, d.nbr, err.p);
    // element -> slot rank map
    PFG_CUDA_TRY(cudaMalloc(&d.rank, std::max<int64_t>(ninc * NNE, 1)));
    k_rank_map<NNE><<<grid_for(ninc), kThreads, 0, st>>>(d.conn, d.blk_ptr, d.nbr, d.own_begin, d.own_end, ninc,
                                                         d.rank);
    PFG_CUDA_TRY(cudaGetLastError());
    d.device_bytes += ninc * 4 + (d.nnodes + 1) * 8 + (nown + 1) * 8 + d.nblocks * 4 + ninc * NNE;
    return PFG_OK;
}

// hex8 chunk-row pass (k_hex8_chunk_rows): the lanes of a node add the block of local column index b at the same
// time, so for every owned node and every b the incident elements must reach pairwise different neighbours.
// True for any mesh in which two elements sharing a node do not hold it and a common second node at the same
// local positions; flags the first violation.
__global__ void k_hex_round_check(const int64_t* __restrict__ inc_ptr, const uint32_t* __restrict__ inc_list,
                                  const uint8_t* __restrict__ rank, int64_t own_begin, int64_t nown,
                                  int* __restrict__ bad) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nown) return;
    const int64_t i0 = inc_ptr[own_begin + r], i1 = inc_ptr[own_begin + r + 1];
    if (i1 - i0 > 8) {
        atomicExch(bad, 1);
        return;
    }
    for (int b = 0; b < 8; ++b) {
        int seen[8];
        for (int64_t i = i0; i < i1; ++i) {
            const int t = rank[(size_t)inc_list[i] * 8 + b];
            for (int64_t q = i0; q < i; ++q)
                if (seen[q - i0] == t) atomicExch(bad, 1);
            seen[i - i0] = t;
        }
    }
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// shared-memory doubles staged per incidence, worst case over the physics a handle with this
// (element, ndof_per_node) can run: m == 1 -> Helmholtz (two matrices) or nonlinear Poisson
// (matrix + residual); m > 1 -> elasticity.  Must match rb_stride() in pfg_assemble.cu.
static int worst_rb_doubles(int nne, int m) {
    if (m == 1) return 2 * nne + 2;
    return nne * m * m + 2;
}

template <int NNE>
static int build_gather_plan(MeshDev& d, cudaStream_t st, Scratch& scratch) {
    const int64_t nown = d.own_end - d.own_begin;
    if (nown == 0 || d.nelems == 0) return PFG_OK;
    if (d.max_valence > kMaxValence || d.max_k > kMaxRowBlocks) return PFG_OK;  // atomic path only

    // ---- chunk size
    d.tile_threads = 128;
    const bool tile = (d.m != 3);  // tile plan for the one-thread-per-element physics, first format for hex8 elasticity
    int64_t C;
    if (tile) {
        // aim at tile_threads element records per chunk: a d-dimensional block of n^d nodes is touched by
        // (n+1)^d elements; tie-aware cuts vary the block shape, so leave some slack
        const double side = std::pow((double)d.tile_threads, 1.0 / d.ndims) - 1.0;
        C = (int64_t)std::floor(std::pow(std::max(1.0, side), (double)d.ndims) * 0.86);
        if (d.m == 2) {
            // row format: three CTAs per SM need <= ~75 KB each.  Per node: staged records (42 doubles each,
            // (n+1)^2 / n^2 records per node), the two image rows (4 k doubles), codes, window and corner tables.
            const double recs_per_node = std::pow((side + 1.0) / side, (double)d.ndims);
            const double per_node = recs_per_node * (42 * 8 + NNE * 2 + 20) + 4.0 * 8 * std::max(1, d.max_k) +
                                    2.0 * NNE * std::max(1, d.max_valence) + 16;
            const int64_t c_smem = (int64_t)(env_int("PFG_TILE_SMEM_BYTES", 75 * 1024) * 0.89 / per_node);
            C = std::min(C, c_smem);
        }
    } else {
        int budget = env_int("PFG_CHUNK_SMEM_BYTES", 160 * 1024);
        int per_node = std::max(1, d.max_valence) * worst_rb_doubles(NNE, d.m) * 8;
        C = budget / per_node * 4 / 5;  // tie-aware cuts may overshoot the target by up to 1/4
        // chunk-row pass of hex8 elasticity (k_hex8_chunk_rows): seven consumer warps x four nodes per round
        if (d.hex_rows_ok) C = env_int("PFG_HEX_CHUNK_NODES", 27);
    }
    C = env_int("PFG_CHUNK_NODES", (int)C);
    C = std::max<int64_t>(4, std::min<int64_t>(C, 1024));

    // ---- chunk id per owned node
    DBuf<uint32_t> group;  // per owned node (index r = node - own_begin)
    PFG_CUDA_TRY(group.alloc(nown));
    PFG_CUDA_TRY(cudaMemsetAsync(group.p, 0, nown * sizeof(uint32_t), st));
    int64_t ngroups = 1;
    if (d.flags & PFG_CREATE_NO_REORDER) {
        k_chunk_by_id<<<grid_for(nown), kThreads, 0, st>>>(nown, C, group.p);
        ngroups = (nown + C - 1) / C;
    } else {
        DBuf<uint64_t> ckeys, ckeys2;
        DBuf<uint32_t> ord0, ord1, ord2, gsorted, gtmp, cut, flag, dense;
        DBuf<int64_t> gstart, vstart;
        PFG_CUDA_TRY(ckeys.alloc(nown));
        PFG_CUDA_TRY(ckeys2.alloc(nown));
        PFG_CUDA_TRY(ord0.alloc(nown));
        PFG_CUDA_TRY(ord1.alloc(nown));
        PFG_CUDA_TRY(ord2.alloc(nown));
        PFG_CUDA_TRY(gsorted.alloc(nown));
        PFG_CUDA_TRY(gtmp.alloc(nown));
        PFG_CUDA_TRY(cut.alloc(nown));
        PFG_CUDA_TRY(flag.alloc(nown));
        PFG_CUDA_TRY(dense.alloc(nown));
        PFG_CUDA_TRY(gstart.alloc(nown));
        PFG_CUDA_TRY(vstart.alloc(nown));
        double avg_group = (double)nown;
        // Row format on lattice-like 2-D meshes: make the chunk rows (nodes that are consecutive in id) a whole number
        // of quarter-warps wide.  Phase B maps consecutive lanes to consecutive chunk nodes; with rows of 8 (or 16)
        // nodes a quarter-warp never straddles a row, so each of its gather loads hits eight consecutive element
        // records -- bank-conflict free by the record stride.  Detected from the data (the direction ids run along
        // and how many nodes share a coordinate); any other mesh keeps the plain count-based tiling.
        int axis_order[3] = {0, 1, 2};
        int64_t level0_target = 0;
        const int tile_width = env_int("PFG_TILE_WIDTH", 8);
        if (tile && d.ndims == 2 && tile_width > 0 && nown > 64) {
            DBuf<unsigned long long> cnt;
            PFG_CUDA_TRY(cnt.alloc(4));
            PFG_CUDA_TRY(cudaMemsetAsync(cnt.p, 0, 4 * sizeof(unsigned long long), st));
            k_id_direction<<<grid_for(nown), kThreads, 0, st>>>(d.X, d.ndims, d.own_begin, nown, cnt.p);
            unsigned long long h_cnt[4] = {0, 0, 0, 0};
            PFG_CUDA_TRY(cudaMemcpyAsync(h_cnt, cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
            PFG_CUDA_TRY(cudaStreamSynchronize(st));
            const int fast = (h_cnt[1] > h_cnt[0]) ? 1 : 0;
            if ((double)h_cnt[fast] >= 0.8 * (double)(nown - 1)) {
                k_coord_keys<<<grid_for(nown), kThreads, 0, st>>>(d.X, d.ndims, fast, d.own_begin, nown, ckeys.p, ord0.p);
                PFG_CUB(scratch, st,
                        cub::DeviceRadixSort::SortKeys(d_temp_storage, temp_storage_bytes, (const uint64_t*)ckeys.p,
                                                       ckeys2.p, nown, 0, 64, st));
                k_count_distinct<<<grid_for(nown), kThreads, 0, st>>>(ckeys2.p, nown, cnt.p + 2);
                PFG_CUDA_TRY(cudaMemcpyAsync(h_cnt, cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
                PFG_CUDA_TRY(cudaStreamSynchronize(st));
                const double run = (double)nown / (double)std::max<unsigned long long>(1, h_cnt[2]);  // nodes per coordinate value
                if (run >= 8.0 && (double)h_cnt[2] >= 2.0 * tile_width) {
                    axis_order[0] = fast;
                    axis_order[1] = 1 - fast;
                    level0_target = (int64_t)std::llround(run * tile_width);
                    if (C > 2 * tile_width) C -= C % tile_width;  // whole rows
                }
            }
        }
        for (int level = 0; level < d.ndims; ++level) {
            int remaining = d.ndims - level;
            const int axis = axis_order[level];
            int64_t target;
            if (remaining == 1) {
                target = C;
            } else if (level == 0 && level0_target > 0) {
                target = level0_target;
            } else {
                double splits = std::ceil(std::pow(std::max(1.0, avg_group / (double)C), 1.0 / remaining) - 1e-9);
                target = (int64_t)std::ceil(avg_group / std::max(1.0, splits));
            }
            target = std::max<int64_t>(1, target);
            // order (group, coord[axis], id): stable sort by coord, then stable sort by group
            k_coord_keys<<<grid_for(nown), kThreads, 0, st>>>(d.X, d.ndims, axis, d.own_begin, nown, ckeys.p, ord0.p);
            PFG_CUB(scratch, st,
                    cub::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, (const uint64_t*)ckeys.p,
                                                    ckeys2.p, (const uint32_t*)ord0.p, ord1.p, nown, 0, 64, st));
            k_gather_u32<<<grid_for(nown), kThreads, 0, st>>>(group.p, ord1.p, nown, gtmp.p);
            PFG_CUB(scratch, st,
                    cub::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, (const uint32_t*)gtmp.p,
                                                    gsorted.p, (const uint32_t*)ord1.p, ord2.p, nown, 0,
                                                    bits_for((uint64_t)ngroups), st));
            k_mark_starts<<<grid_for(nown), kThreads, 0, st>>>(gsorted.p, ord2.p, d.X, d.ndims, axis, d.own_begin,
                                                              nown, gstart.p, vstart.p);
            PFG_CUB(scratch, st,
                    cub::DeviceScan::InclusiveScan(d_temp_storage, temp_storage_bytes, gstart.p, gstart.p,
                                                   cub::Max(), nown, st));
            PFG_CUB(scratch, st,
                    cub::DeviceScan::InclusiveScan(d_temp_storage, temp_storage_bytes, vstart.p, vstart.p,
                                                   cub::Max(), nown, st));
            k_cut_flags<<<grid_for(nown), kThreads, 0, st>>>(gsorted.p, gstart.p, vstart.p, nown, target, cut.p,
                                                            flag.p);
            PFG_CUB(scratch, st,
                    cub::DeviceScan::InclusiveSum(d_temp_storage, temp_storage_bytes, flag.p, dense.p, nown, st));
            k_scatter_group<<<grid_for(nown), kThreads, 0, st>>>(ord2.p, dense.p, nown, group.p);
            uint32_t h_ng = 0;
            PFG_CUDA_TRY(cudaMemcpyAsync(&h_ng, dense.p + (nown - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            PFG_CUDA_TRY(cudaStreamSynchronize(st));
            ngroups = h_ng;
            avg_group = (double)nown / (double)ngroups;
        }
    }
    d.nchunks = ngroups;

    // ---- chunk-ordered node slots: stable sort of id-ordered nodes by chunk id
    DBuf<uint32_t> ids, slot_node, slot_chunk, valence, kk, words;
    PFG_CUDA_TRY(ids.alloc(nown));
    PFG_CUDA_TRY(slot_node.alloc(nown));
    PFG_CUDA_TRY(slot_chunk.alloc(nown));
    PFG_CUDA_TRY(valence.alloc(nown + 1));
    PFG_CUDA_TRY(kk.alloc(nown + 1));
    PFG_CUDA_TRY(words.alloc(nown + 1));
    k_iota_u32<<<grid_for(nown), kThreads, 0, st>>>(ids.p, nown);
    PFG_CUB(scratch, st,
            cub::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, (const uint32_t*)group.p, slot_chunk.p,
                                            (const uint32_t*)ids.p, slot_node.p, nown, 0,
                                            bits_for((uint64_t)d.nchunks), st));
    PFG_CUDA_TRY(cudaMemsetAsync(valence.p, 0, (nown + 1) * sizeof(uint32_t), st));
    PFG_CUDA_TRY(cudaMemsetAsync(kk.p, 0, (nown + 1) * sizeof(uint32_t), st));
    PFG_CUDA_TRY(cudaMemsetAsync(words.p, 0, (nown + 1) * sizeof(uint32_t), st));
    k_slot_counts<<<grid_for(nown), kThreads, 0, st>>>(slot_node.p, d.inc_ptr, d.blk_ptr, d.own_begin, nown,
                                                      valence.p, kk.p);
    DBuf<int64_t> inc_excl, plan_off;
    PFG_CUDA_TRY(inc_excl.alloc(nown + 1));
    PFG_CUDA_TRY(plan_off.alloc(nown + 1));
    PFG_CUB(scratch, st,
            cub::DeviceScan::ExclusiveSum(d_temp_storage, temp_storage_bytes, valence.p, inc_excl.p, nown + 1, st));
    k_plan_sizes<<<grid_for(nown), kThreads, 0, st>>>(valence.p, kk.p, NNE, nown, words.p);
    PFG_CUB(scratch, st,
            cub::DeviceScan::ExclusiveSum(d_temp_storage, temp_storage_bytes, words.p, plan_off.p, nown + 1, st));

    PFG_CUDA_TRY(cudaMalloc(&d.chunks, d.nchunks * sizeof(ChunkHdr)));
    PFG_CUDA_TRY(cudaMemsetAsync(d.chunks, 0, d.nchunks * sizeof(ChunkHdr), st));
    DBuf<int> maxima;
    PFG_CUDA_TRY(maxima.alloc(12));
    PFG_CUDA_TRY(cudaMemsetAsync(maxima.p, 0, 12 * sizeof(int), st));
    k_chunk_node_begin<<<grid_for(nown), kThreads, 0, st>>>(slot_chunk.p, nown, d.nchunks, d.chunks);
    k_chunk_finish<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, inc_excl.p, kk.p, plan_off.p, d.chunks, nown,
                                                             maxima.p);

    // ---- element records: unique (chunk, element) over the owned incidences
    int64_t h_ninc_own = 0, h_plan_words = 0;
    PFG_CUDA_TRY(cudaMemcpyAsync(&h_ninc_own, inc_excl.p + nown, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaMemcpyAsync(&h_plan_words, plan_off.p + nown, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_plan_words >= (int64_t)0xffffffffll) {
        set_error("gather plan pool exceeds 16 GiB");
        return PFG_ERR_UNSUPPORTED;
    }
    DBuf<uint64_t> ikeys, ikeys_sorted, rec_keys;
    DBuf<int64_t> nsel;
    PFG_CUDA_TRY(ikeys.alloc(h_ninc_own));
    PFG_CUDA_TRY(ikeys_sorted.alloc(h_ninc_own));
    PFG_CUDA_TRY(rec_keys.alloc(h_ninc_own));
    PFG_CUDA_TRY(nsel.alloc(1));
    k_inc_keys<NNE><<<grid_for(nown), kThreads, 0, st>>>(slot_node.p, slot_chunk.p, d.inc_ptr, d.inc_list, inc_excl.p,
                                                        d.own_begin, nown, ikeys.p);
    PFG_CUB(scratch, st,
            cub::DeviceRadixSort::SortKeys(d_temp_storage, temp_storage_bytes, (const uint64_t*)ikeys.p,
                                           ikeys_sorted.p, h_ninc_own, 0, 32 + bits_for((uint64_t)d.nchunks), st));
    PFG_CUB(scratch, st,
            cub::DeviceSelect::Unique(d_temp_storage, temp_storage_bytes, ikeys_sorted.p, rec_keys.p, nsel.p,
                                      h_ninc_own, st));
    PFG_CUDA_TRY(cudaMemcpyAsync(&d.nrecs, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaStreamSynchronize(st));
    ikeys.release();
    ikeys_sorted.release();
    if (d.nrecs >= (int64_t)0xffffffffll) {
        set_error("too many element records for the gather plan");
        return PFG_ERR_UNSUPPORTED;
    }
    PFG_CUDA_TRY(cudaMalloc(&d.rec_nodes, d.nrecs * NNE * sizeof(int32_t) + 64));  // bulk copies may over-read
    PFG_CUDA_TRY(cudaMalloc(&d.rec_dst, d.nrecs * NNE * sizeof(uint16_t) + 32));
    PFG_CUDA_TRY(cudaMalloc(&d.rec_elem, d.nrecs * sizeof(int32_t)));
    k_fill_records<NNE><<<grid_for(d.nrecs), kThreads, 0, st>>>(rec_keys.p, d.nrecs, d.conn, d.rec_nodes, d.rec_dst,
                                                               d.rec_elem, d.chunks, maxima.p);
    k_chunk_rec_finish<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, d.chunks, maxima.p);
    if (tile) {
        // ---- tile plan: node windows, then blob (header, node table, runs) and contribution codes per chunk
        cudaFree(d.rec_dst);
        d.rec_dst = nullptr;
        DBuf<uint32_t> win_begin, win_chunk;
        DBuf<int> terr;
        PFG_CUDA_TRY(terr.alloc(1));
        PFG_CUDA_TRY(cudaMemsetAsync(terr.p, 0, sizeof(int), st));
        PFG_CUDA_TRY(win_begin.alloc(d.nchunks + 1));
        {   // node windows
            const int64_t ncorners = d.nrecs * NNE;
            DBuf<uint64_t> wkeys, wsorted, wuniq;
            DBuf<int64_t> nsel2;
            PFG_CUDA_TRY(wkeys.alloc(ncorners));
            PFG_CUDA_TRY(wsorted.alloc(ncorners));
            PFG_CUDA_TRY(nsel2.alloc(1));
            k_win_keys<NNE><<<grid_for(ncorners), kThreads, 0, st>>>(rec_keys.p, d.rec_nodes, ncorners, wkeys.p);
            PFG_CUB(scratch, st,
                    cub::DeviceRadixSort::SortKeys(d_temp_storage, temp_storage_bytes, (const uint64_t*)wkeys.p, wsorted.p,
                                                   ncorners, 0, 32 + bits_for((uint64_t)d.nchunks), st));
            wkeys.release();
            PFG_CUDA_TRY(wuniq.alloc(ncorners));
            PFG_CUB(scratch, st,
                    cub::DeviceSelect::Unique(d_temp_storage, temp_storage_bytes, wsorted.p, wuniq.p, nsel2.p, ncorners,
                                              st));
            PFG_CUDA_TRY(cudaMemcpyAsync(&d.nwin, nsel2.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            PFG_CUDA_TRY(cudaStreamSynchronize(st));
            wsorted.release();
            if (d.nwin >= (int64_t)0xffffffffll) {
                set_error("node windows exceed the 32-bit index range");
                return PFG_ERR_UNSUPPORTED;
            }
            PFG_CUDA_TRY(cudaMalloc(&d.win_nodes, d.nwin * sizeof(uint32_t) + 64));
            PFG_CUDA_TRY(cudaMalloc(&d.rec_local, ncorners * sizeof(uint16_t) + 64));
            PFG_CUDA_TRY(cudaMemsetAsync(win_begin.p, 0, (d.nchunks + 1) * sizeof(uint32_t), st));
            PFG_CUDA_TRY(win_chunk.alloc(d.nwin));
            k_win_fill<<<grid_for(d.nwin), kThreads, 0, st>>>(wuniq.p, d.nwin, d.win_nodes, win_begin.p, win_chunk.p);
            k_win_max<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, win_begin.p, d.nwin, maxima.p);
            k_rec_local<NNE><<<grid_for(ncorners), kThreads, 0, st>>>(rec_keys.p, d.rec_nodes, wuniq.p, win_begin.p,
                                                                     d.nchunks, d.nwin, ncorners, d.rec_local, terr.p);
            PFG_CUDA_TRY(cudaStreamSynchronize(st));
            PFG_CUDA_TRY(cudaGetLastError());
        }
        DBuf<uint32_t> blob_len16, code_len16, run_flag, run_id;
        DBuf<int64_t> blob_off, code_off;
        DBuf<int> chunk_gmax, chunk_gvmax;
        PFG_CUDA_TRY(chunk_gmax.alloc(d.nchunks));
        PFG_CUDA_TRY(chunk_gvmax.alloc(d.nchunks));
        PFG_CUDA_TRY(cudaMemsetAsync(chunk_gmax.p, 0, d.nchunks * sizeof(int), st));
        PFG_CUDA_TRY(cudaMemsetAsync(chunk_gvmax.p, 0, d.nchunks * sizeof(int), st));
        PFG_CUDA_TRY(run_flag.alloc(nown + 1));
        PFG_CUDA_TRY(run_id.alloc(nown + 1));
        PFG_CUDA_TRY(blob_len16.alloc(d.nchunks + 1));
        PFG_CUDA_TRY(code_len16.alloc(d.nchunks + 1));
        PFG_CUDA_TRY(blob_off.alloc(d.nchunks + 1));
        PFG_CUDA_TRY(code_off.alloc(d.nchunks + 1));
        PFG_CUDA_TRY(cudaMemsetAsync(run_flag.p, 0, (nown + 1) * sizeof(uint32_t), st));
        PFG_CUDA_TRY(cudaMemsetAsync(blob_len16.p, 0, (d.nchunks + 1) * sizeof(uint32_t), st));
        PFG_CUDA_TRY(cudaMemsetAsync(code_len16.p, 0, (d.nchunks + 1) * sizeof(uint32_t), st));
        k_tile_node_sizes<<<grid_for(nown), kThreads, 0, st>>>(valence.p, slot_chunk.p, NNE, nown, chunk_gmax.p,
                                                               chunk_gvmax.p);
        k_tile_run_flags<<<grid_for(nown), kThreads, 0, st>>>(slot_node.p, slot_chunk.p, nown, run_flag.p);
        PFG_CUB(scratch, st,
                cub::DeviceScan::InclusiveSum(d_temp_storage, temp_storage_bytes, run_flag.p, run_id.p, nown + 1, st));
        k_tile_chunk_sizes<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, d.chunks, run_id.p, chunk_gmax.p,
                                                                     chunk_gvmax.p, d.m, blob_len16.p, code_len16.p,
                                                                     maxima.p);
        PFG_CUB(scratch, st,
                cub::DeviceScan::ExclusiveSum(d_temp_storage, temp_storage_bytes, blob_len16.p, blob_off.p,
                                              d.nchunks + 1, st));
        PFG_CUB(scratch, st,
                cub::DeviceScan::ExclusiveSum(d_temp_storage, temp_storage_bytes, code_len16.p, code_off.p,
                                              d.nchunks + 1, st));
        int64_t h_blob16 = 0, h_code16 = 0;
        PFG_CUDA_TRY(cudaMemcpyAsync(&h_blob16, blob_off.p + d.nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaMemcpyAsync(&h_code16, code_off.p + d.nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        if (h_blob16 >= 0xffffffffll || h_code16 >= 0xffffffffll) {
            set_error("tile plan pools exceed 64 GiB");
            return PFG_ERR_UNSUPPORTED;
        }
        d.tile_blob_bytes = h_blob16 * 16;
        d.tile_ncodes = h_code16 * 8;
        PFG_CUDA_TRY(cudaMalloc(&d.tile_dir, (d.nchunks + 1) * sizeof(TileDir)));
        PFG_CUDA_TRY(cudaMalloc(&d.tile_blob, d.tile_blob_bytes + 64));
        PFG_CUDA_TRY(cudaMalloc(&d.tile_codes, d.tile_ncodes * 2 + 64));
        PFG_CUDA_TRY(cudaMalloc(&d.tile_codes_neutral, d.tile_ncodes * 2 + 64));
        PFG_CUDA_TRY(cudaMemsetAsync(d.tile_blob, 0, d.tile_blob_bytes + 64, st));
        PFG_CUDA_TRY(cudaMemsetAsync(d.tile_codes, 0, d.tile_ncodes * 2 + 64, st));
        PFG_CUDA_TRY(cudaMemsetAsync(d.tile_codes_neutral, 0xFF, d.tile_ncodes * 2 + 64, st));  // 0xFFFF = padding
        PFG_CUDA_TRY(cudaMalloc(&d.cnode_id, nown * sizeof(int32_t)));
        k_tile_dir<<<grid_for(d.nchunks + 1), kThreads, 0, st>>>(d.nchunks, d.chunks, blob_off.p, code_off.p, blob_len16.p,
                                                                 code_len16.p, win_begin.p, d.win_nodes, slot_node.p,
                                                                 d.blk_ptr, d.m, d.nwin, d.tile_dir, terr.p);
        k_win_relative<<<grid_for(d.nwin), kThreads, 0, st>>>(d.nchunks, d.tile_dir, d.nwin, win_chunk.p, d.win_nodes);
        TileFillArgs fa;
        fa.slot_node = slot_node.p, fa.slot_chunk = slot_chunk.p;
        fa.inc_ptr = d.inc_ptr, fa.inc_list = d.inc_list, fa.rank = d.rank, fa.blk_ptr = d.blk_ptr;
        fa.chunks = d.chunks, fa.dir = d.tile_dir, fa.rec_keys = rec_keys.p;
        fa.run_flag = run_flag.p, fa.run_id = run_id.p;
        fa.own_begin = d.own_begin, fa.nslots = nown, fa.m = d.m;
        fa.blob_pool = d.tile_blob, fa.codes_neutral = d.tile_codes_neutral, fa.cnode_id = d.cnode_id;
        fa.err = terr.p, fa.chunk_gmax = chunk_gmax.p, fa.chunk_gvmax = chunk_gvmax.p, fa.maxima = maxima.p;
        k_tile_fill<NNE><<<grid_for(nown, 128), 128, 0, st>>>(fa);
        k_tile_image_layout<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, d.chunks, d.tile_dir, slot_node.p,
                                                                      d.blk_ptr, run_flag.p, d.m, d.tile_blob, maxima.p,
                                                                      terr.p);
        // ---- chunk templates: chunks with byte-identical tables share the tables of the first of them
        int64_t compact_win = -1, compact_recs = -1;  // entries kept after compaction (-1: pools not compacted)
        d.ntemplates = d.nchunks;
        d.plan_read_bytes = 0;
        if (!env_int("PFG_NO_TEMPLATES", 0)) {
            DBuf<uint64_t> hash, hash_sorted;
            DBuf<uint32_t> ids, ids_sorted, rep;
            DBuf<int64_t> head;
            DBuf<int> mismatch;
            DBuf<unsigned long long> stats;
            DBuf<TileDir> dir2;
            PFG_CUDA_TRY(hash.alloc(d.nchunks));
            PFG_CUDA_TRY(hash_sorted.alloc(d.nchunks));
            PFG_CUDA_TRY(ids.alloc(d.nchunks));
            PFG_CUDA_TRY(ids_sorted.alloc(d.nchunks));
            PFG_CUDA_TRY(rep.alloc(d.nchunks));
            PFG_CUDA_TRY(head.alloc(d.nchunks));
            PFG_CUDA_TRY(mismatch.alloc(1));
            PFG_CUDA_TRY(stats.alloc(2));
            PFG_CUDA_TRY(dir2.alloc(d.nchunks + 1));
            PFG_CUDA_TRY(cudaMemsetAsync(mismatch.p, 0, sizeof(int), st));
            PFG_CUDA_TRY(cudaMemsetAsync(stats.p, 0, 2 * sizeof(unsigned long long), st));
            const unsigned warp_grid = grid_for(d.nchunks * 32);
            k_tile_hash<<<warp_grid, kThreads, 0, st>>>(d.nchunks, d.tile_dir, d.tile_blob, d.tile_codes_neutral, d.win_nodes,
                                                        d.rec_local, NNE, hash.p, ids.p);
            PFG_CUB(scratch, st,
                    cub::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, (const uint64_t*)hash.p,
                                                    hash_sorted.p, (const uint32_t*)ids.p, ids_sorted.p, d.nchunks, 0, 64,
                                                    st));
            k_tile_group_heads<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, hash_sorted.p, head.p);
            PFG_CUB(scratch, st,
                    cub::DeviceScan::InclusiveScan(d_temp_storage, temp_storage_bytes, head.p, head.p, cub::Max(),
                                                   d.nchunks, st));
            k_tile_representative<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, ids_sorted.p, head.p, rep.p);
            k_tile_verify<<<warp_grid, kThreads, 0, st>>>(d.nchunks, d.tile_dir, rep.p, d.tile_blob, d.tile_codes_neutral,
                                                          d.win_nodes, d.rec_local, NNE, mismatch.p);
            int h_mismatch = 0;
            PFG_CUDA_TRY(cudaMemcpyAsync(&h_mismatch, mismatch.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            PFG_CUDA_TRY(cudaStreamSynchronize(st));
            if (!h_mismatch) {  // (a 64-bit hash collision between different tables: keep every chunk its own template)
                PFG_CUDA_TRY(cudaMemcpyAsync(dir2.p, d.tile_dir, (d.nchunks + 1) * sizeof(TileDir),
                                             cudaMemcpyDeviceToDevice, st));
                k_tile_apply_templates<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, rep.p, d.tile_dir, dir2.p, NNE,
                                                                                 stats.p);
                unsigned long long h_stats[2] = {0, 0};
                PFG_CUDA_TRY(cudaMemcpyAsync(h_stats, stats.p, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
                PFG_CUDA_TRY(cudaStreamSynchronize(st));
                d.ntemplates = (int64_t)h_stats[0];
                d.plan_read_bytes = (int64_t)h_stats[1];
                // ---- keep only the templates' tables: compact pools, directory re-pointed, the rest freed
                if (d.ntemplates * 2 <= d.nchunks && !env_int("PFG_NO_COMPACT", 0)) {
                    DBuf<int64_t> sz[4], off[4];
                    for (int k = 0; k < 4; ++k) {
                        PFG_CUDA_TRY(sz[k].alloc(d.nchunks + 1));
                        PFG_CUDA_TRY(off[k].alloc(d.nchunks + 1));
                    }
                    k_tile_template_sizes<<<grid_for(d.nchunks + 1), kThreads, 0, st>>>(d.nchunks, d.tile_dir, NNE, sz[0].p,
                                                                                        sz[1].p, sz[2].p, sz[3].p);
                    int64_t total[4] = {0, 0, 0, 0};
                    for (int k = 0; k < 4; ++k) {
                        PFG_CUB(scratch, st,
                                cub::DeviceScan::ExclusiveSum(d_temp_storage, temp_storage_bytes, sz[k].p, off[k].p,
                                                              d.nchunks + 1, st));
                        PFG_CUDA_TRY(cudaMemcpyAsync(&total[k], off[k].p + d.nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
                    }
                    PFG_CUDA_TRY(cudaStreamSynchronize(st));
                    uint8_t* blob2 = nullptr;
                    uint16_t *codes2 = nullptr, *neutral2 = nullptr, *loc2 = nullptr;
                    uint32_t* win2 = nullptr;
                    PFG_CUDA_TRY(cudaMalloc(&blob2, total[0] * 16 + 64));
                    PFG_CUDA_TRY(cudaMalloc(&codes2, total[1] * 16 + 64));
                    PFG_CUDA_TRY(cudaMalloc(&neutral2, total[1] * 16 + 64));
                    PFG_CUDA_TRY(cudaMalloc(&win2, total[2] * 4 + 64));
                    PFG_CUDA_TRY(cudaMalloc(&loc2, total[3] * NNE * 2 + 64));
                    PFG_CUDA_TRY(cudaMemsetAsync(blob2, 0, total[0] * 16 + 64, st));
                    PFG_CUDA_TRY(cudaMemsetAsync(codes2, 0, total[1] * 16 + 64, st));
                    PFG_CUDA_TRY(cudaMemsetAsync(neutral2, 0xFF, total[1] * 16 + 64, st));
                    PFG_CUDA_TRY(cudaMemsetAsync(win2, 0, total[2] * 4 + 64, st));
                    PFG_CUDA_TRY(cudaMemsetAsync(loc2, 0, total[3] * NNE * 2 + 64, st));
                    k_tile_compact_copy<<<warp_grid, kThreads, 0, st>>>(d.nchunks, d.tile_dir, NNE, off[0].p, off[1].p, off[2].p,
                                                                        off[3].p, d.tile_blob, d.tile_codes_neutral, d.win_nodes,
                                                                        d.rec_local, blob2, neutral2, win2, loc2);
                    k_tile_compact_dir<<<grid_for(d.nchunks), kThreads, 0, st>>>(d.nchunks, off[0].p, off[1].p, off[2].p,
                                                                                 off[3].p, d.tile_dir);
                    PFG_CUDA_TRY(cudaStreamSynchronize(st));
                    PFG_CUDA_TRY(cudaGetLastError());
                    cudaFree(d.tile_blob), cudaFree(d.tile_codes), cudaFree(d.tile_codes_neutral);
                    cudaFree(d.win_nodes), cudaFree(d.rec_local);
                    d.tile_blob = blob2, d.tile_codes = codes2, d.tile_codes_neutral = neutral2;
                    d.win_nodes = win2, d.rec_local = loc2;
                    d.tile_blob_bytes = total[0] * 16;
                    d.tile_ncodes = total[1] * 8;
                    compact_win = total[2], compact_recs = total[3];
                }
            }
        }
        int h_terr = 0, h_max[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        PFG_CUDA_TRY(cudaMemcpyAsync(&h_terr, terr.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaMemcpyAsync(h_max, maxima.p, sizeof(h_max), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        PFG_CUDA_TRY(cudaGetLastError());
        d.max_chunk_inc = h_max[0];
        d.max_chunk_nodes = h_max[1];
        d.max_kpad = h_max[2];
        d.max_chunk_recs = h_max[3];
        d.max_blob_bytes = h_max[5];
        d.max_code_bytes = h_max[6];
        d.max_out_bytes = h_max[7];
        d.max_chunk_win = h_max[8];
        if (h_terr || d.max_blob_bytes > 0xFFFF * 16 || d.max_code_bytes > 0xFFFF * 16) {
            // chunk too irregular for the compact tile encoding: assemble with the atomic scatter instead
            cudaFree(d.tile_dir); d.tile_dir = nullptr;
            cudaFree(d.tile_blob); d.tile_blob = nullptr;
            cudaFree(d.tile_codes); d.tile_codes = nullptr;
            cudaFree(d.tile_codes_neutral); d.tile_codes_neutral = nullptr;
            cudaFree(d.win_nodes); d.win_nodes = nullptr;
            cudaFree(d.rec_local); d.rec_local = nullptr;
            d.nchunks = 0;
            return PFG_OK;
        }
        cudaFree(d.rec_nodes);  // the tile kernels address record corners through the node windows
        d.rec_nodes = nullptr;
        d.plan_bytes = d.tile_blob_bytes + d.tile_ncodes * 2;
        if (d.plan_read_bytes == 0)  // no template sharing: every chunk reads its own tables
            d.plan_read_bytes = d.plan_bytes + d.nrecs * (NNE * 2) + d.nwin * 4;
        d.plan_read_bytes += d.nchunks * (int64_t)sizeof(TileDir);
        d.device_bytes += d.nchunks * (sizeof(ChunkHdr) + sizeof(TileDir)) + nown * 4 + d.nrecs * 4 +
                          (compact_recs >= 0 ? compact_recs : d.nrecs) * (NNE * 2) +
                          (compact_win >= 0 ? compact_win : d.nwin) * 4 + d.tile_blob_bytes + d.tile_ncodes * 4;
        return PFG_OK;
    }
    k_fill_dst<NNE><<<grid_for(nown), kThreads, 0, st>>>(slot_node.p, slot_chunk.p, d.inc_ptr, d.inc_list, inc_excl.p,
                                                        d.chunks, rec_keys.p, d.own_begin, nown, d.rec_dst);
    if (NNE == 8 && d.hex_rows_ok) {
        PFG_CUDA_TRY(cudaMalloc(&d.inc_rec8, nown * 8 * sizeof(uint32_t)));
        PFG_CUDA_TRY(cudaMalloc(&d.inc_ranks8, nown * 8 * sizeof(uint64_t)));
        k_fill_inc8<<<grid_for(nown), kThreads, 0, st>>>(slot_node.p, slot_chunk.p, d.inc_ptr, d.inc_list, d.chunks,
                                                         rec_keys.p, d.rank, d.own_begin, nown, d.inc_rec8,
                                                         d.inc_ranks8);
        d.device_bytes += nown * 8 * (int64_t)(sizeof(uint32_t) + sizeof(uint64_t));
    }

    // ---- per-node plans (the hex8 chunk-row pass needs the node table only: no plan pool, no record corner tables)
    const bool rows_only = (NNE == 8 && d.hex_rows_ok);
    d.plan_bytes = rows_only ? 0 : h_plan_words * 4;
    if (!rows_only) {
        // +32 bytes: a chunk's plan is fetched with 16-byte aligned bulk copies that may over-read either end
        PFG_CUDA_TRY(cudaMalloc(&d.plan_pool, d.plan_bytes + 32));
        PFG_CUDA_TRY(cudaMemsetAsync(d.plan_pool, 0, d.plan_bytes + 32, st));
    }
    PFG_CUDA_TRY(cudaMalloc(&d.cnodes, nown * sizeof(ChunkNode)));
    PFG_CUDA_TRY(cudaMalloc(&d.cnode_id, nown * sizeof(int32_t)));
    k_fill_plans<NNE><<<grid_for(nown, 128), 128, 0, st>>>(slot_node.p, d.inc_ptr, d.inc_list, d.rank, d.blk_ptr,
                                                          inc_excl.p, d.chunks, slot_chunk.p, plan_off.p, d.own_begin,
                                                          nown, d.m, d.cnodes, d.cnode_id, d.plan_pool);
    int h_max[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    PFG_CUDA_TRY(cudaMemcpyAsync(h_max, maxima.p, sizeof(h_max), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaStreamSynchronize(st));
    PFG_CUDA_TRY(cudaGetLastError());
    d.max_chunk_inc = h_max[0];
    d.max_chunk_nodes = h_max[1];
    d.max_kpad = h_max[2];
    d.max_chunk_recs = h_max[3];
    d.max_chunk_plan_words = h_max[4];
    if (d.max_chunk_inc > 0xFFFE) {
        set_error("chunk with %d incidences exceeds the 16-bit slot range", d.max_chunk_inc);
        return PFG_ERR_UNSUPPORTED;
    }
    if (rows_only) {
        cudaFree(d.rec_nodes);
        cudaFree(d.rec_dst);
        d.rec_nodes = nullptr;
        d.rec_dst = nullptr;
        if (hex_rows_smem(d.max_chunk_recs, d.max_k).total > kMaxDynamicSmem) {
            // chunks with too many element records for the geometry ring (very irregular meshes): atomic scatter only
            cudaFree(d.inc_rec8);
            cudaFree(d.inc_ranks8);
            d.inc_rec8 = nullptr;
            d.inc_ranks8 = nullptr;
            d.hex_rows_ok = 0;
        }
        d.device_bytes += d.nchunks * sizeof(ChunkHdr) + nown * (sizeof(ChunkNode) + 4) + d.nrecs * 4;
        return PFG_OK;
    }
    d.device_bytes += d.nchunks * sizeof(ChunkHdr) + nown * (sizeof(ChunkNode) + 4) + d.nrecs * (NNE * 6 + 4) +
                      d.plan_bytes;
    return PFG_OK;
}

static void free_all(MeshDev& d) {
    void* ptrs[] = {d.X, d.conn, d.gid, d.inc_ptr, d.inc_list, d.blk_ptr, d.nbr, d.rank, d.chunks, d.cnodes,
                    d.cnode_id, d.rec_nodes, d.rec_dst, d.rec_elem, d.plan_pool, d.tile_dir, d.tile_blob,
                    d.tile_codes, d.tile_codes_neutral, d.win_nodes, d.rec_local, d.elem_skip, d.hex_geo, d.inc_rec8, d.inc_ranks8,
                    d.bc_fixed, d.bc_u0, d.cg_work, d.rec_skip, d.trank};
    for (void* p : ptrs)
        if (p) cudaFree(p);
}

template <int NNE>
static int create_impl(MeshDev& d, const double* X_dev, const int64_t* conn_dev, const int64_t* gid_dev,
                       cudaStream_t st) {
    Scratch scratch;
    const int64_t ninc = d.nelems * NNE;
    PFG_CUDA_TRY(cudaMalloc(&d.X, std::max<int64_t>(d.nnodes * d.ndims, 1) * sizeof(double)));
    PFG_CUDA_TRY(cudaMalloc(&d.conn, std::max<int64_t>(ninc, 1) * sizeof(int32_t)));
    PFG_CUDA_TRY(cudaMemcpyAsync(d.X, X_dev, d.nnodes * d.ndims * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (gid_dev) {
        PFG_CUDA_TRY(cudaMalloc(&d.gid, d.nnodes * sizeof(int64_t)));
        PFG_CUDA_TRY(cudaMemcpyAsync(d.gid, gid_dev, d.nnodes * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    }
    d.device_bytes += d.nnodes * d.ndims * 8 + ninc * 4 + (gid_dev ? d.nnodes * 8 : 0);
    DBuf<long long> minmax;
    PFG_CUDA_TRY(minmax.alloc(2));
    long long init[2] = {LLONG_MAX, LLONG_MIN};
    PFG_CUDA_TRY(cudaMemcpyAsync(minmax.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_conn_to_i32<<<grid_for(ninc), kThreads, 0, st>>>(conn_dev, d.conn, ninc, minmax.p);
    long long h_mm[2];
    PFG_CUDA_TRY(cudaMemcpyAsync(h_mm, minmax.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    PFG_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_mm[0] != 0 || h_mm[1] != d.nnodes - 1) {
        // the reference asserts exactly this (pyfem.py:680-681)
        set_error("conn.min() == %lld, conn.max() == %lld; expected 0 and nnodes-1 == %lld", h_mm[0], h_mm[1],
                  (long long)(d.nnodes - 1));
        return PFG_ERR_MESH;
    }
    PFG_TRY(build_pattern<NNE>(d, st, scratch));
    if (NNE == 8 && d.m == 3 && d.max_valence <= 8 && d.max_k <= 48 && d.own_end > d.own_begin) {
        DBuf<int> bad;
        PFG_CUDA_TRY(bad.alloc(1));
        PFG_CUDA_TRY(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
        const int64_t nown = d.own_end - d.own_begin;
        k_hex_round_check<<<grid_for(nown), kThreads, 0, st>>>(d.inc_ptr, d.inc_list, d.rank, d.own_begin, nown, bad.p);
        int h_bad = 1;
        PFG_CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        d.hex_rows_ok = h_bad ? 0 : 1;
    }
    if (!(d.flags & PFG_CREATE_NO_GATHER_PLAN)) PFG_TRY(build_gather_plan<NNE>(d, st, scratch));
    PFG_CUDA_TRY(cudaStreamSynchronize(st));
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

}  // namespace pfg

using namespace pfg;

extern "C" int pfg_abi_version(void) { return PFG_ABI_VERSION; }

extern "C" const char* pfg_last_error(void) { return g_last_error.c_str(); }

extern "C" int pfg_mesh_create(pfg_mesh** out, int elem_type, int ndof_per_node, int64_t nnodes, int64_t nelems,
                               const double* X_dev, const int64_t* conn_dev, int64_t own_begin, int64_t own_end,
                               const int64_t* node_gid_dev, int64_t ncols_global_nodes, int flags, void* stream) {
    if (!out) {
        set_error("pfg_mesh_create: out is NULL");
        return PFG_ERR_INVALID;
    }
    *out = nullptr;
    if (elem_type != PFG_QUAD4 && elem_type != PFG_HEX8) {
        set_error("unsupported element type %d (device path covers quad4 and hex8)", elem_type);
        return PFG_ERR_UNSUPPORTED;
    }
    int ndims = (elem_type == PFG_QUAD4) ? 2 : 3;
    if (ndof_per_node != 1 && ndof_per_node != ndims) {
        set_error("ndof_per_node must be 1 or %d for this element, got %d", ndims, ndof_per_node);
        return PFG_ERR_INVALID;
    }
    if (nnodes <= 0 || nelems <= 0 || !X_dev || !conn_dev) {
        set_error("empty mesh: nnodes=%lld nelems=%lld", (long long)nnodes, (long long)nelems);
        return PFG_ERR_INVALID;
    }
    if (nnodes > 0x7fffffffll || nelems * elem_type > 0xffffffffll) {
        set_error("mesh too large for 32-bit node / incidence ids");
        return PFG_ERR_UNSUPPORTED;
    }
    if (own_begin < 0 || own_end > nnodes || own_begin > own_end) {
        set_error("invalid owned row range [%lld, %lld)", (long long)own_begin, (long long)own_end);
        return PFG_ERR_INVALID;
    }
    pfg_mesh* mesh = new pfg_mesh();
    MeshDev& d = mesh->d;
    d.elem_type = elem_type;
    d.nne = elem_type;
    d.ndims = ndims;
    d.nquads = elem_type;
    d.m = ndof_per_node;
    d.nnodes = nnodes;
    d.nelems = nelems;
    d.own_begin = own_begin;
    d.own_end = own_end;
    d.ncols_nodes = node_gid_dev ? ncols_global_nodes : nnodes;
    d.flags = flags;
    // the handle lives on the device that owns X_dev (one handle per GPU / rank); this library's CUDA
    // runtime instance is separate from the caller's, so the device is selected explicitly
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, X_dev) != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
        cudaGetLastError();
        set_error("X_dev is not a device pointer");
        delete mesh;
        return PFG_ERR_INVALID;
    }
    d.device = attr.device;
    if (cudaSetDevice(d.device) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed", d.device);
        delete mesh;
        return PFG_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device);
    int rc = (elem_type == PFG_QUAD4) ? create_impl<4>(d, X_dev, conn_dev, node_gid_dev, (cudaStream_t)stream)
                                      : create_impl<8>(d, X_dev, conn_dev, node_gid_dev, (cudaStream_t)stream);
    if (rc != PFG_OK) {
        free_all(d);
        delete mesh;
        return rc;
    }
    // scipy's index-width rule for coo->csr: int32 iff max(coo nnz, ncols) fits (scipy/sparse/_coo.py:59-61,419)
    int64_t D = (int64_t)d.nne * d.m;
    int64_t coo_nnz = d.nelems * D * D;
    int64_t ncols = d.ncols_nodes * d.m;
    d.idx_bytes = (std::max(coo_nnz, ncols) <= 0x7fffffffll) ? 4 : 8;
    *out = mesh;
    return PFG_OK;
}

// tile plan: one skip flag per element record (chunk tables are shared between chunks, so the flag cannot ride in them)
__global__ void k_mask_records(const int32_t* __restrict__ rec_elem, const uint8_t* __restrict__ skip, int64_t nrecs,
                               uint8_t* __restrict__ rec_skip) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < nrecs) rec_skip[r] = skip[rec_elem[r]] ? 1 : 0;
}

extern "C" int pfg_mesh_set_element_mask(pfg_mesh* mesh, const uint8_t* elem_skip_dev, void* stream) {
    if (!mesh) {
        set_error("pfg_mesh_set_element_mask: mesh is NULL");
        return PFG_ERR_INVALID;
    }
    MeshDev& d = mesh->d;
    PFG_CUDA_TRY(cudaSetDevice(d.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (elem_skip_dev) {
        if (!d.elem_skip) PFG_CUDA_TRY(cudaMalloc(&d.elem_skip, std::max<int64_t>(d.nelems, 1)));
        PFG_CUDA_TRY(cudaMemcpyAsync(d.elem_skip, elem_skip_dev, d.nelems, cudaMemcpyDeviceToDevice, st));
        if (d.tile_dir && d.nrecs) {
            if (!d.rec_skip) PFG_CUDA_TRY(cudaMalloc(&d.rec_skip, d.nrecs + 256));  // the kernels prefetch past the end
            k_mask_records<<<grid_for(d.nrecs), kThreads, 0, st>>>(d.rec_elem, d.elem_skip, d.nrecs, d.rec_skip);
        }
    } else if (d.elem_skip) {
        PFG_CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(d.elem_skip);
        d.elem_skip = nullptr;
        if (d.rec_skip) cudaFree(d.rec_skip);
        d.rec_skip = nullptr;
    }
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}

extern "C" int pfg_mesh_destroy(pfg_mesh* mesh) {
    if (!mesh) return PFG_OK;
    free_all(mesh->d);
    delete mesh;
    return PFG_OK;
}

extern "C" int pfg_mesh_get(const pfg_mesh* mesh, int what, int64_t* value) {
    if (!mesh || !value) {
        set_error("pfg_mesh_get: NULL argument");
        return PFG_ERR_INVALID;
    }
    const MeshDev& d = mesh->d;
    switch (what) {
        case PFG_INFO_NNZ: *value = d.nnz; break;
        case PFG_INFO_NROWS: *value = (d.own_end - d.own_begin) * d.m; break;
        case PFG_INFO_NCOLS: *value = d.ncols_nodes * d.m; break;
        case PFG_INFO_IDX_BYTES: *value = d.idx_bytes; break;
        case PFG_INFO_NCHUNKS: *value = d.nchunks; break;
        case PFG_INFO_CHUNK_ELEMS: *value = d.nrecs; break;
        case PFG_INFO_PLAN_BYTES:
            if (d.tile_dir)  // what one assembly reads: the directory and the distinct chunk templates
                *value = d.plan_read_bytes;
            else if (d.inc_rec8)  // hex8 chunk-row pass: chunk headers, node table, incidence tables, record elements
                *value = (int64_t)(d.nchunks * sizeof(ChunkHdr) +
                                   (d.own_end - d.own_begin) * (sizeof(ChunkNode) + 8 * (4 + 8)) + d.nrecs * 4);
            else
                *value = d.nchunks ? (int64_t)(d.nchunks * sizeof(ChunkHdr) + (d.own_end - d.own_begin) * sizeof(ChunkNode) +
                                               d.nrecs * (d.nne * 6) + d.plan_bytes)
                                   : 0;
            break;
        case PFG_INFO_DEVICE_BYTES: *value = d.device_bytes; break;
        case PFG_INFO_MAX_ROW_BLOCKS: *value = d.max_k; break;
        case PFG_INFO_MAX_VALENCE: *value = d.max_valence; break;
        case PFG_INFO_HEX_ROWS: *value = (d.hex_rows_ok == 1 && d.nchunks > 0 && d.inc_rec8) ? 1 : 0; break;
        case PFG_INFO_TEMPLATES: *value = d.tile_dir ? d.ntemplates : 0; break;
        default:
            set_error("pfg_mesh_get: unknown query %d", what);
            return PFG_ERR_INVALID;
    }
    return PFG_OK;
}

extern "C" int pfg_mesh_pattern(const pfg_mesh* mesh, void* indptr_dev, void* indices_dev, int idx_bytes,
                                void* stream) {
    if (!mesh || !indptr_dev || !indices_dev || (idx_bytes != 4 && idx_bytes != 8)) {
        set_error("pfg_mesh_pattern: invalid argument");
        return PFG_ERR_INVALID;
    }
    const MeshDev& d = mesh->d;
    PFG_CUDA_TRY(cudaSetDevice(d.device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nown = d.own_end - d.own_begin;
    if (idx_bytes == 4 && (d.nnz > 0x7fffffffll || d.ncols_nodes * d.m > 0x7fffffffll)) {
        set_error("pattern does not fit 32-bit indices (nnz=%lld)", (long long)d.nnz);
        return PFG_ERR_INVALID;
    }
    int64_t nrows1 = nown * d.m + 1;
    unsigned gx = 65535u * 16u;
    int64_t nb = (d.nblocks + kThreads - 1) / kThreads;
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(nb, gx)), (unsigned)std::max<int64_t>(1, (nb + gx - 1) / gx));
    if (idx_bytes == 4) {
        k_write_indptr<int32_t><<<grid_for(nrows1), kThreads, 0, st>>>(d.blk_ptr, nown, d.m, (int32_t*)indptr_dev);
        if (d.nblocks) k_write_indices<int32_t><<<grid, kThreads, 0, st>>>(d.blk_ptr, d.nbr, d.gid, nown, d.m, (int32_t*)indices_dev);
    } else {
        k_write_indptr<int64_t><<<grid_for(nrows1), kThreads, 0, st>>>(d.blk_ptr, nown, d.m, (int64_t*)indptr_dev);
        if (d.nblocks) k_write_indices<int64_t><<<grid, kThreads, 0, st>>>(d.blk_ptr, d.nbr, d.gid, nown, d.m, (int64_t*)indices_dev);
    }
    PFG_CUDA_TRY(cudaGetLastError());
    return PFG_OK;
}
