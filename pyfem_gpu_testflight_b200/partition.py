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

    def __init__(self, X, conn, own_range, node_gid, nnodes_global, elem_gid, rank, size, nelems_global=None):
        self.nelems_global = nelems_global  # elements of the whole mesh (index-width rule of the global matrix)
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
                     node_gid.astype(np.int64), nnodes, elem_gid, rank, size, nelems_global=int(conn.shape[0]))


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
    return LocalMesh(X, conn, own, node_gid, nslow * plane, elem_gid, rank, size,
                     nelems_global=int(elems_per_layer) * (nslow - 1))


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


class SlabContext:
    """What a physics model needs to run as ONE RANK of a row-slab partition (models.py, `group=` / `partition=`):
    the rank's local mesh, the node ranges of all ranks, global <-> local nodal fields, and the gathers that rebuild
    the reference's global CSR / vectors for parity checks.  Host logic only (numpy + torch.distributed): runs under
    gloo on a CPU as well as under NCCL."""

    def __init__(self, X, conn, group=None, partition=None, ranges=None, rank=None, size=None):
        import torch.distributed as dist
        self.group = group
        if partition is not None:
            rank, size = partition.rank, partition.size
        elif rank is None or size is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("group= needs an initialised torch.distributed process group")
            rank, size = dist.get_rank(group), dist.get_world_size(group)
        self.rank, self.size = int(rank), int(size)
        self.part = partition if partition is not None else partition_mesh(X, conn, self.rank, self.size,
                                                                           node_ranges=ranges)
        self.ranges = list(ranges) if ranges is not None else split_range(self.part.nnodes_global, self.size)
        gb, ge = self.part.owned_global_range
        if (gb, ge) != tuple(self.ranges[self.rank]):
            raise ValueError(f"partition owns global nodes [{gb}, {ge}), the ranges say {self.ranges[self.rank]}")

    @property
    def owned_nodes(self):
        """Global node range [begin, end) whose rows this rank assembles."""
        return self.part.owned_global_range

    def local_field(self, f):
        """A nodal field given on the GLOBAL mesh (as the reference's callers pass it) -> the rank's local nodes;
        scalars (constant fields) pass through."""
        if f is None or np.ndim(f) == 0:
            return f
        f = np.asarray(f) if not hasattr(f, "device") else f
        if f.shape[0] == len(self.part.node_gid):
            return f  # already local (a single rank's local nodes are the global nodes)
        if f.shape[0] == self.part.nnodes_global:
            return f[self.part.node_gid] if not hasattr(f, "device") else f[_as_index(f, self.part.node_gid)]
        raise ValueError(f"nodal field has {f.shape[0]} entries, expected {self.part.nnodes_global} (global) or "
                         f"{len(self.part.node_gid)} (this rank's local nodes)")

    def gather_matrix(self, K_slab, dst=0):
        """Row-wise concatenation of every rank's slab (scipy CSR, global columns) on rank `dst`: the reference's
        global matrix.  For checks and small problems -- the slabs travel as pickled host arrays."""
        import torch.distributed as dist
        mine = (np.asarray(K_slab.indptr), np.asarray(K_slab.indices), np.asarray(K_slab.data))
        out = [None] * self.size if self.rank == dst else None
        dist.gather_object(mine, out, dst=_global_rank(self.group, dst), group=self.group)
        if self.rank != dst:
            return None
        K = concat_slabs(out, K_slab.shape[1])
        idx = K_slab.indices.dtype  # every rank used the global index-width rule
        return type(K)((K.data, K.indices.astype(idx), K.indptr.astype(idx)), shape=K.shape)

    def gather_vector(self, v_owned, dst=0):
        import torch.distributed as dist
        out = [None] * self.size if self.rank == dst else None
        dist.gather_object(np.asarray(v_owned), out, dst=_global_rank(self.group, dst), group=self.group)
        return np.concatenate(out) if self.rank == dst else None


def _global_rank(group, r):
    import torch.distributed as dist
    return r if group is None else dist.get_global_rank(group, r)


def _as_index(like, idx):
    import torch
    return torch.as_tensor(idx, device=like.device)
