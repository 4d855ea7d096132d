// Internal definitions shared by the setup (pattern / plan builder) and assembly kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pyfem_b200.h"

namespace pfg {

// ---------------------------------------------------------------------------------------------
// error plumbing: status codes out, message kept per host thread (pfg_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define PFG_CUDA_TRY(expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            pfg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? PFG_ERR_NOMEM : PFG_ERR_CUDA;            \
        }                                                                                      \
    } while (0)

#define PFG_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != PFG_OK) return _s; \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Gather-plan layout (see DESIGN.md "Data layout in HBM")
//
// Owned node rows are grouped into chunks; one CTA assembles one chunk:
//   phase A  every element touching the chunk is integrated once by one thread (hex8 elasticity:
//            8 threads) and the row blocks that belong to chunk nodes are staged in shared memory,
//            one "incidence" slot per (chunk node, adjacent element) pair, node-major;
//   phase B  one thread per (chunk node, neighbour node) block sums the <= valence contributions
//            listed in the node's plan in a fixed order and stores the m x m block to CSR.
// ---------------------------------------------------------------------------------------------
struct __align__(16) ChunkHdr {  // 48 bytes
    uint32_t node_begin; // first chunk-ordered node slot
    uint32_t n_nodes;
    uint32_t n_recs;
    uint32_t n_inc;      // incidences = sum of valence over chunk nodes
    int64_t rec_begin;   // first element record
    uint32_t plan_begin; // first plan word (4-byte units) of the chunk's nodes in the plan pool
    uint32_t plan_words; // plan words of the whole chunk (contiguous)
    uint32_t kpad;       // max neighbour count over chunk nodes
    uint32_t pad_[3];
};

struct __align__(16) ChunkNode { // 16 bytes per owned node, chunk-ordered
    int64_t gslot;       // offset of the node's first dof row in the owned CSR values
    uint32_t plan;       // offset of the node's plan record in the plan pool, 4-byte units
    uint16_t inc_base;   // first incidence slot of this node inside the chunk
    uint8_t k;           // neighbour nodes (row length in node blocks)
    uint8_t valence;     // adjacent elements
};
// plan record of a node: uint8 start[k+1] (prefix of contribution counts per neighbour),
// then uint8 src[valence*nne] sorted by neighbour rank, src = (adjacent-element ordinal << 3) | local column node.

// ---------------------------------------------------------------------------------------------
// Gather plan, second format ("tile plan"): used by every one-thread-per-element physics
// (quad4 all models, hex8 scalar models).  A chunk's element matrices are staged per element RECORD
// (upper triangle only for symmetric operators), so phase A needs no destination table, and
// phase B reads each contribution at a shared-memory offset that is stored ready-made in the plan.
//
// Per chunk, three global arrays feed one CTA iteration:
//   tile_dir[c]      48 B   the chunk's bases and where its (template's) blob / codes / window / corner tables live
//   blob             16 B aligned: TileHdr | TileNode[n_nodes] | format-specific tables (below)
//   codes            16 B aligned: per node uint16 contribution codes, sorted by neighbour rank
//   win_nodes        uint32 sorted unique node ids of the chunk's records, relative to TileDir::node_base (the "node
//                    window": coordinates and nodal fields are gathered once per window node into shared memory)
//   rec_local        (n_recs, nne) uint16 window index of every record corner
// A code is (offset << 2) | end << 1 | transpose: offset in units of the staging block (16 B for 2x2
// blocks, 8 B for scalars) from the start of the chunk's staging area, whose record slot 0 is all zeros
// (code 0 adds nothing); `end` marks the last contribution of a (node, neighbour) block.  Codes depend on
// the staging layout of the operator (record stride, symmetric or full storage); they are kept in the
// neutral form (end << 15 | vec << 14 | record << 2*lb | a << lb | b, 0xFFFF = padding) beside the working
// copy, which is re-encoded when an operator with another layout runs.
//
// Blob: TileHdr | TileNode[n_nodes] | (scalar handles) uint32 row[n_nodes] = owned-row index of each node relative to
// TileDir::row_base (vector outputs), padded to 8 B | TileRun[n_runs].
// Codes: group-major, [group][node][8 codes], every node padded with code 0 to the chunk's TileHdr::gmax groups, so
// that a warp's 16-byte code loads are contiguous; scalar handles append [group][node][4] vector codes (the node's
// incidences, for residual / right-hand-side outputs), padded to TileHdr::gvmax groups.
// Phase B runs one thread per chunk node: it walks the node's codes, sums every (node, neighbour) block in plan order
// and drops the node's dof rows into a shared-memory image of the CSR values; the image leaves through one TMA bulk
// store per run of consecutive node ids (scalar runs: the 16-byte aligned middle part; a leading / trailing odd value
// goes by a plain store).
// ---------------------------------------------------------------------------------------------
// Chunk TEMPLATES.  Everything a chunk's tables hold is relative to per-chunk bases kept in its directory entry (CSR
// slots relative to gbase, window node ids to node_base, owned-row indices to row_base, records and corner indices to
// the chunk itself), so chunks with the same local topology -- on a lattice-like mesh all interior chunks -- have
// byte-identical tables.  The plan builder hashes and compares them and points every chunk at the tables of the first
// chunk with the same content (its template).  A CTA keeps the tables of the running template in shared memory and
// reloads only when the template changes: on a structured mesh the 0.96 GB of plan reads per assembly of the 16.8 M
// quad case shrink to the directory (48 B per chunk) plus a few KB of L2-resident templates.
struct __align__(16) TileDir {  // 48 bytes
    int64_t gbase;         // CSR value slot of the chunk's first node (node slots are relative to it)
    uint32_t blob_off16;   // template blob in the blob pool, 16-byte units
    uint32_t code_off16;   // template codes in the code pool, 16-byte units
    uint32_t loc_off;      // first record of the template's corner-index table (rec_local), in records
    uint32_t win_off;      // first entry of the template's node window (win_nodes)
    uint32_t rec_begin;    // first record of THIS chunk (element ids for Op::NEEDS_ELEM, skip flags)
    uint32_t node_base;    // the window holds node ids relative to this
    uint32_t row_base;     // scalar handles: owned-row indices of the blob are relative to this
    uint32_t tmpl;         // template id = id of the first chunk with the same tables
    uint16_t blob_len16, code_len16;  // table lengths, 16-byte units
    uint16_t n_recs, n_win;
};
static_assert(sizeof(TileDir) == 48, "TileDir layout");

struct __align__(16) TileHdr {  // 16 bytes, first bytes of the blob
    uint16_t n_nodes, n_recs;
    uint16_t gmax;         // matrix code groups (8 codes) per node, padded to the chunk maximum
    uint16_t gvmax;        // vector code groups (4 codes) per node (scalar handles)
    uint16_t n_runs;       // runs of consecutive node ids
    uint16_t pad_[3];
};

struct __align__(8) TileNode {  // 8 bytes
    uint32_t gslot_rel;    // first value slot of the node's first dof row, relative to TileDir::gbase
    uint16_t aux;          // offset of the node's rows in the chunk's CSR image, in units (16 B for m = 2, 8 B for m = 1)
    uint16_t k;            // neighbour count: a dof row holds m * k values
};

struct __align__(8) TileRun {  // 8 bytes: one run of consecutive node ids, image -> CSR values
    uint32_t gslot_rel;    // first value slot, relative to TileDir::gbase
    uint16_t out_off;      // image offset, in units
    uint16_t len;          // length, in units
};

__host__ __device__ inline int64_t tile_blob_tables(int64_t n_nodes, int m) {  // offset of the run table in a blob
    return (int64_t)sizeof(TileHdr) + (int64_t)sizeof(TileNode) * n_nodes + (m == 1 ? ((4 * n_nodes + 7) / 8) * 8 : 0);
}

struct TileLayout {  // how an operator stages one record; decides the code encoding
    int nne = 0;       // nodes per element
    int unit_shift = 0;// log2 of bytes per code unit (3 or 4)
    int rec_units = 0; // record stride in code units
    int sym = 0;       // upper-triangle storage
    int blk_units = 0; // units per node-pair block
    int has_mat = 0;   // 0: vector-only operator
    int vec_units = -1;// offset of the vector entries inside a record, in code units (-1: the operator has none)
    bool operator==(const TileLayout& o) const {
        return nne == o.nne && unit_shift == o.unit_shift && rec_units == o.rec_units && sym == o.sym &&
               blk_units == o.blk_units && has_mat == o.has_mat && vec_units == o.vec_units;
    }
};

// hex8 chunk-row pass (k_hex8_chunk_rows): shared-memory layout shared by the plan builder (does a mesh fit?) and
// the launch
constexpr int kHexRowWarps = 7;                     // consumer warps, four chunk nodes each (+ one producer warp)
constexpr int kHexGeoRecordBytes = (8 * 10 + 2) * 8; // staged geometry of one element record: 640 B + 16 B bank shift
struct HexRowsSmem {
    int off_geo, geo_stage_bytes, off_image, image_stride, off_meta;
    size_t total;
};
inline HexRowsSmem hex_rows_smem(int max_chunk_recs, int max_k) {
    HexRowsSmem L;
    L.image_stride = 9 * max_k;  // doubles per chunk node in a warp's image
    L.off_geo = 128;
    L.geo_stage_bytes = (max_chunk_recs * kHexGeoRecordBytes + 15) & ~15;
    L.off_image = L.off_geo + 2 * L.geo_stage_bytes;
    L.off_meta = L.off_image + kHexRowWarps * 4 * L.image_stride * 8;
    L.total = (size_t)L.off_meta + (size_t)kHexRowWarps * (32 * 64 + 16);
    return L;
}
constexpr size_t kMaxDynamicSmem = 227 * 1024;

constexpr int kMaxValence = 31;      // 5-bit ordinal; start[] must fit uint8 (31*8 = 248)
constexpr int kMaxRowBlocks = 255;   // rank fits uint8
constexpr uint16_t kNoDst = 0xFFFF;

struct MeshDev {
    // dimensions
    int elem_type = 0, nne = 0, ndims = 0, m = 0, nquads = 0;
    int64_t nnodes = 0, nelems = 0, own_begin = 0, own_end = 0, ncols_nodes = 0;
    int64_t nblocks = 0;      // node-level nnz of owned rows
    int64_t nnz = 0;          // dof-level nnz of owned rows = m*m*nblocks
    int idx_bytes = 4;
    int max_k = 0, max_valence = 0;
    int flags = 0;

    // mesh (device)
    double* X = nullptr;        // (nnodes, ndims)
    int32_t* conn = nullptr;    // (nelems, nne)
    int64_t* gid = nullptr;     // optional local->global node id
    uint8_t* elem_skip = nullptr; // optional (nelems): elements kept in the pattern but not integrated here

    // node -> element incidences (all local nodes), sorted by (node, element, local index)
    int64_t* inc_ptr = nullptr;   // (nnodes+1)
    uint32_t* inc_list = nullptr; // (nelems*nne) values e*nne + a

    // node-level CSR pattern of the owned rows
    int64_t* blk_ptr = nullptr;   // (nown+1) prefix of neighbour counts
    int32_t* nbr = nullptr;       // (nblocks) sorted neighbour (local) node ids
    uint8_t* rank = nullptr;      // (nelems*nne*nne) rank of conn[e][b] in the row of conn[e][a] (garbage for non-owned rows)

    // gather plan
    int64_t nchunks = 0, nrecs = 0, plan_bytes = 0;
    ChunkHdr* chunks = nullptr;
    ChunkNode* cnodes = nullptr;  // (nown)
    int32_t* cnode_id = nullptr;  // (nown) local node id of each chunk-ordered slot
    int32_t* rec_nodes = nullptr; // (nrecs, nne)
    uint16_t* rec_dst = nullptr;  // (nrecs, nne)
    int32_t* rec_elem = nullptr;  // (nrecs)
    uint8_t* plan_pool = nullptr;
    int max_chunk_inc = 0, max_chunk_nodes = 0, max_chunk_recs = 0, max_kpad = 0, max_chunk_plan_words = 0;

    // tile plan (second format); present when tile_dir != nullptr
    TileDir* tile_dir = nullptr;       // (nchunks)
    uint8_t* tile_blob = nullptr;      // blob pool
    uint16_t* tile_codes = nullptr;    // working codes, encoded for tile_layout
    uint16_t* tile_codes_neutral = nullptr;  // (record << 2*lb | a << lb | b), lb = log2(nne)
    uint32_t* win_nodes = nullptr;     // node windows, chunk after chunk
    uint16_t* rec_local = nullptr;     // (nrecs, nne) window index of every record corner
    uint8_t* rec_skip = nullptr;       // (nrecs) records of masked elements (pfg_mesh_set_element_mask), else null
    int64_t ntemplates = 0;            // distinct chunk templates
    int64_t plan_read_bytes = 0;       // directory + template tables: what one assembly reads of the plan
    int64_t nwin = 0;
    int max_chunk_win = 0;
    int64_t tile_blob_bytes = 0, tile_ncodes = 0;
    int max_blob_bytes = 0, max_code_bytes = 0, max_out_bytes = 0;
    TileLayout tile_layout;            // encoding of tile_codes (nne == 0: not encoded yet)
    int tile_threads = 128;            // CTA size of the tile kernels = target element records per chunk

    // hex8 3-D elasticity, owner-computes passes (k_hex8_geometry / k_hex8_chunk_rows)
    int hex_rows_ok = 0;               // 1: every owned node has <= 8 elements, <= 48 neighbours and conflict-free rounds
    double* hex_geo = nullptr;         // scratch (nelems, 8, 10), allocated by the first assembly that needs it
    uint32_t* inc_rec8 = nullptr;      // (nown, 8) chunk-ordered node slot x incidence -> record * 8 + local node, 0xFFFF: none
    uint64_t* inc_ranks8 = nullptr;    // (nown, 8) the incidence's eight ranks, rank[(e*8 + a)*8 + 0..7]

    // scratch of the device consumers (pfg_solve.cu), allocated on first use
    uint8_t* bc_fixed = nullptr;       // (ncols) flags of the fixed dofs
    double* bc_u0 = nullptr;           // (ncols) prescribed values, defined at the fixed dofs
    double* cg_work = nullptr;         // r, z, p, Ap, 1/diag and the dot-product partial sums
    uint8_t* trank = nullptr;          // (nblocks) transposed-slot map of pfg_spmv_t, built on first use

    int64_t device_bytes = 0;
    int sm_count = 148;
    int device = 0;
};

// (re-)encode the working codes of the tile plan for an operator's staging layout
int tile_prepare_layout(MeshDev& d, const TileLayout& L, cudaStream_t st);

}  // namespace pfg

struct pfg_mesh {
    pfg::MeshDev d;
};
