"""Element-block / row-slab partitioning for one-process-per-GPU assembly.

Rank r owns a contiguous range of node rows (a slab of the global CSR matrix).  Its local mesh is its
block of elements plus the one layer of ghost elements that touch an owned node, so every owned row
is assembled completely on the owning GPU: the matrix and rhs need NO value exchange.  Local nodes are
renumbered in increasing global order, which keeps each row's sorted-column order identical to the
global pattern; `node_gid` maps local columns back to global ids.  The global CSR is the row-wise
concatenation of the rank slabs.  The only collective the path needs is the scalar all-reduce of the
Newton residual norm (reference pyfem.py:2344), see `global_norm`.
"""
import numpy as np


def split_range(n, size):
    """`size` balanced contiguous ranges of range(n): list of (begin, end)."""
    base, rem = divmod(int(n), int(size))
    out, b = [], 0
    for r in range(size):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


class LocalMesh:
    """One rank's piece: arrays to hand to DeviceMesh / the model constructors."""

    def __init__(self, X, conn, own_range, node_gid, nnodes_global, elem_gid, rank, size):
        self.X = X
        self.conn = conn
        self.own_range = own_range          # (begin, end) in LOCAL node numbering
        self.node_gid = node_gid            # (nnodes_local,) global node id of each local node, increasing
        self.nnodes_global = int(nnodes_global)
        self.elem_gid = elem_gid            # (nelems_local,) global element ids (block + ghosts)
        self.rank, self.size = rank, size

    @property
    def owned_global_range(self):
        return int(self.node_gid[self.own_range[0]]), int(self.node_gid[self.own_range[1] - 1]) + 1


def partition_mesh(X, conn, rank, size, node_ranges=None):
    """General meshes: split node ids into `size` contiguous ranges (or use `node_ranges`), keep every
    element that touches an owned node.  Works for any connectivity whose numbering has locality."""
    X = np.asarray(X)
    conn = np.asarray(conn)
    nnodes = X.shape[0]
    b, e = (node_ranges or split_range(nnodes, size))[rank]
    touches = ((conn >= b) & (conn < e)).any(axis=1)
    elem_gid = np.nonzero(touches)[0]
    sub = conn[elem_gid]
    node_gid = np.unique(sub)
    conn_local = np.searchsorted(node_gid, sub)
    lb, le = np.searchsorted(node_gid, [b, e])
    if le - lb != e - b:
        raise ValueError("an owned node is not referenced by any element")
    return LocalMesh(np.ascontiguousarray(X[node_gid]), conn_local.astype(np.int64), (int(lb), int(le)),
                     node_gid.astype(np.int64), nnodes, elem_gid, rank, size)


def structured_slab(nnodes_x, nnodes_y, nnodes_z, rank, size, Lx=None, Ly=None, Lz=None):
    """The rank's slab of a ProblemCreator mesh (pyfem.py:2469-2535) generated directly, without building
    the global arrays: slabs in y for quads (nnodes_z=None), in z for hex blocks.  Equal to
    partition_mesh(ProblemCreator(...)) with plane-aligned node ranges."""
    three_d = nnodes_z is not None
    nzz = nnodes_z if three_d else 1
    Lx = (nnodes_x - 1) / (nnodes_y - 1) if Lx is None else Lx
    Ly = 1.0 if Ly is None else Ly
    Lz = (nzz - 1) / (nnodes_y - 1) if Lz is None else Lz
    x = np.linspace(0, Lx, nnodes_x)
    y = np.linspace(0, Ly, nnodes_y)
    z = np.linspace(0, Lz, nzz)
    nslow = nzz if three_d else nnodes_y            # node layers along the slab axis
    plane = nnodes_x * nnodes_y if three_d else nnodes_x
    rb, re = split_range(nslow, size)[rank]          # owned node layers
    e0, e1 = max(rb - 1, 0), min(re, nslow - 1)      # element layers touching them
    l0, l1 = e0, e1                                  # local node layers [l0, l1] (inclusive)
    nl = l1 - l0 + 1
    if three_d:
        X = np.empty((nl, nnodes_y, nnodes_x, 3))
        X[..., 0] = x[None, None, :]
        X[..., 1] = y[None, :, None]
        X[..., 2] = z[l0:l1 + 1, None, None]
        X = X.reshape(-1, 3)
        ids = np.arange(nl * plane, dtype=np.int64).reshape(nl, nnodes_y, nnodes_x)
        lo, hi = ids[:-1], ids[1:]
        corners = [lo[:, :-1, :-1], lo[:, :-1, 1:], lo[:, 1:, 1:], lo[:, 1:, :-1],
                   hi[:, :-1, :-1], hi[:, :-1, 1:], hi[:, 1:, 1:], hi[:, 1:, :-1]]
        elems_per_layer = (nnodes_x - 1) * (nnodes_y - 1)
    else:
        X = np.empty((nl, nnodes_x, 2))
        X[..., 0] = x[None, :]
        X[..., 1] = y[l0:l1 + 1, None]
        X = X.reshape(-1, 2)
        ids = np.arange(nl * plane, dtype=np.int64).reshape(nl, nnodes_x)
        corners = [ids[:-1, :-1], ids[:-1, 1:], ids[1:, 1:], ids[1:, :-1]]
        elems_per_layer = nnodes_x - 1
    conn = np.stack([c.ravel() for c in corners], axis=1)
    node_gid = np.arange(l0 * plane, (l1 + 1) * plane, dtype=np.int64)
    elem_gid = np.arange(e0 * elems_per_layer, e1 * elems_per_layer, dtype=np.int64)
    own = ((rb - l0) * plane, (re - l0) * plane)
    return LocalMesh(X, conn, own, node_gid, nslow * plane, elem_gid, rank, size)


def slab_node_ranges(nnodes_x, nnodes_y, nnodes_z, size):
    """Global node-id ranges owned by the ranks of structured_slab(...): whole node layers along the slab axis."""
    three_d = nnodes_z is not None
    nslow = nnodes_z if three_d else nnodes_y
    plane = nnodes_x * nnodes_y if three_d else nnodes_x
    return [(b * plane, e * plane) for b, e in split_range(nslow, size)]


def concat_slabs(slabs, ncols):
    """Row-wise concatenation of per-rank CSR slabs (indptr, indices, data) into the global matrix."""
    from scipy import sparse
    indptr = [np.zeros(1, dtype=np.int64)]
    off = 0
    for ip, _, _ in slabs:
        indptr.append(np.asarray(ip[1:], dtype=np.int64) + off)
        off += int(ip[-1])
    indptr = np.concatenate(indptr)
    indices = np.concatenate([np.asarray(ix) for _, ix, _ in slabs])
    data = np.concatenate([np.asarray(d) for _, _, d in slabs])
    return sparse.csr_matrix((data, indices, indptr), shape=(len(indptr) - 1, ncols))


def global_norm(local_owned_vec):
    """sqrt of the sum over ranks of |v_owned|^2: the Newton residual norm of Assembler.solve_nonlinear
    (pyfem.py:2344) when the residual is distributed by row slabs.  NCCL (or gloo) all-reduce of one double."""
    import torch
    import torch.distributed as dist
    v = torch.as_tensor(local_owned_vec)
    s = (v.double() * v.double()).sum().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(torch.sqrt(s).item())
