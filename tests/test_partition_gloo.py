"""The multi-rank host logic under torch.distributed (gloo, world_size 2, CPU): every rank partitions the mesh
(element block + ghost layer, order-preserving renumbering), assembles its row slab -- here with the numpy oracle
standing in for the GPU, since this container has none -- and the gathered slabs must reproduce the global
CSR bit for bit; the distributed residual norm must equal the serial one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyfem_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, three_d, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyfem_gpu_testflight_b200.partition import concat_slabs, global_norm, partition_mesh, structured_slab
    dims = (6, 5, 7) if three_d else (11, 9, None)
    X, conn = orc.structured_mesh(*dims)
    rng = np.random.default_rng(3)
    X = X + rng.uniform(-0.01, 0.01, size=X.shape)
    m = X.shape[1]
    part = partition_mesh(X, conn, rank, world)
    slab = structured_slab(*dims, rank, world)
    assert slab.conn.shape[1] == conn.shape[1] and slab.nnodes_global == X.shape[0]
    # local assembly of the owned rows
    Kl = orc.assemble_elasticity(part.X, part.conn)
    lb, le = part.own_range
    rows = Kl[lb * m: le * m]
    gcols = m * part.node_gid[rows.indices // m] + rows.indices % m
    mine = (rows.indptr.astype(np.int64), gcols.astype(np.int64), rows.data)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # distributed norm of a row-distributed vector
    v = np.arange(X.shape[0] * m, dtype=float)
    gb, ge = part.owned_global_range
    nrm = global_norm(torch.from_numpy(v[gb * m: ge * m]))
    if rank == 0:
        Kg = orc.assemble_elasticity(X, conn)
        K = concat_slabs(gathered, Kg.shape[1])
        ok = (np.array_equal(K.indptr, Kg.indptr) and np.array_equal(K.indices, Kg.indices)
              and np.max(np.abs(K.data - Kg.data)) <= 1e-13 * np.max(np.abs(Kg.data))
              and abs(nrm - np.linalg.norm(v)) <= 1e-9 * np.linalg.norm(v))
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("three_d", [False, True])
def test_row_slab_partition_world2(tmp_path, three_d):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, three_d, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"


def test_structured_slab_equals_general_partition():
    from pyfem_gpu_testflight_b200.partition import partition_mesh, split_range, structured_slab
    for dims in [(7, 9, None), (5, 4, 6)]:
        nx, ny, nz = dims
        X, conn = orc.structured_mesh(nx, ny, nz)
        plane = nx if nz is None else nx * ny
        nslow = ny if nz is None else nz
        for size in (1, 2, 3, 4):
            ranges = [(b * plane, e * plane) for b, e in split_range(nslow, size)]
            for r in range(size):
                a = structured_slab(nx, ny, nz, r, size)
                b = partition_mesh(X, conn, r, size, ranges)
                assert np.array_equal(a.conn, b.conn) and np.allclose(a.X, b.X)
                assert a.own_range == b.own_range and np.array_equal(a.node_gid, b.node_gid)
                assert np.array_equal(a.elem_gid, b.elem_gid)


def _slab_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyfem_gpu_testflight_b200.partition import SlabContext
    from scipy import sparse
    X, conn = orc.structured_mesh(13, 10)
    X = X + np.random.default_rng(3).uniform(-0.01, 0.01, size=X.shape)
    ctx = SlabContext(X, conn, group=dist.group.WORLD)  # what ModelBase(..., group=...) builds
    part = ctx.part
    rho = 0.1 + np.random.default_rng(1).random(X.shape[0])
    assert np.array_equal(ctx.local_field(rho), rho[part.node_gid]) and ctx.local_field(2.5) == 2.5
    assert ctx.local_field(rho[part.node_gid]) is not None  # an already-local field passes
    # the rank's slab, with the oracle standing in for the device handle
    Kl = orc.assemble_poisson(part.X, part.conn, ctx.local_field(rho), 3.0)
    lb, le = part.own_range
    rows = Kl[lb:le]
    slab = sparse.csr_matrix((rows.data, part.node_gid[rows.indices].astype(np.int32), rows.indptr),
                             shape=(le - lb, X.shape[0]))
    K = ctx.gather_matrix(slab)
    gb, ge = ctx.owned_nodes
    v = ctx.gather_vector(np.arange(gb, ge, dtype=float))
    if rank == 0:
        Kg = orc.assemble_poisson(X, conn, rho, 3.0)
        ok = (K.indices.dtype == np.int32 and np.array_equal(K.indptr, Kg.indptr) and np.array_equal(K.indices, Kg.indices)
              and np.max(np.abs(K.data - Kg.data)) <= 1e-13 * np.max(np.abs(Kg.data))
              and np.array_equal(v, np.arange(X.shape[0], dtype=float)))
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    else:
        assert K is None and v is None
    dist.barrier()
    dist.destroy_process_group()


def test_slab_context_world2(tmp_path):
    """Host logic behind ModelBase(group=...): partition from the process group, global -> local fields, gathers."""
    port = _free_port()
    mp.spawn(_slab_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"


def test_index_bytes_rule():
    from pyfem_gpu_testflight_b200.engine import index_bytes_rule
    assert index_bytes_rule(16777216, 8, 33570818) == 4        # C2: COO nnz = 2^30 <= 2^31 - 1
    assert index_bytes_rule(16777216 * 2, 8, 2 * 33570818) == 8
    assert index_bytes_rule(16777216, 24, 50923779) == 8       # C5
    assert index_bytes_rule(10, 4, 2 ** 31) == 8               # many columns alone switch the width


def cg_reference(matvec, exchange, b_owned, row0, ncols, dinv, rtol=1e-8, atol=0.0, max_iter=1000):
    """The recurrence pfg_cg_dist runs, with torch tensors and a caller-supplied slab product: the CPU model of the
    distributed solve (test support for the HOST logic -- halo plan, exchange, collectives; the product path is
    slab_solve.SlabKrylov on the device)."""
    import torch
    n = b_owned.numel()
    x = torch.zeros_like(b_owned)
    p_full = torch.zeros(ncols, dtype=torch.float64, device=b_owned.device)
    p = p_full[row0:row0 + n]
    r = b_owned.clone()
    z = dinv * r
    p.copy_(z)
    rz = exchange.all_reduce(torch.dot(r, z).reshape(1))
    rr = exchange.all_reduce(torch.dot(r, r).reshape(1))
    bb = exchange.all_reduce(torch.dot(b_owned, b_owned).reshape(1))
    target = max(rtol * float(bb.sqrt()), atol)
    it = 0
    while float(rr.sqrt()) > target and it < max_iter:
        exchange.refresh(p_full)
        Ap = matvec(p_full)
        pAp = exchange.all_reduce(torch.dot(p, Ap).reshape(1))
        alpha = rz / pAp
        x += alpha * p
        r -= alpha * Ap
        z = dinv * r
        rz_new = exchange.all_reduce(torch.dot(r, z).reshape(1))
        rr = exchange.all_reduce(torch.dot(r, r).reshape(1))
        p.copy_(z + (rz_new / rz) * p)
        rz = rz_new
        it += 1
    return x, it, float(rr.sqrt())


def _cg_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyfem_gpu_testflight_b200.partition import SlabContext
    from pyfem_gpu_testflight_b200.slab_solve import HaloExchange
    from scipy import sparse
    from scipy.sparse.linalg import spsolve
    ok = True
    for dims, m in (((17, 12, None), 1), ((6, 5, 7), 3)):
        X, conn = orc.structured_mesh(*dims)
        X = X + np.random.default_rng(3).uniform(-0.01, 0.01, size=X.shape)
        N = X.shape[0]
        rho = 0.2 + np.random.default_rng(1).random(N)
        Kg = (orc.assemble_poisson(X, conn, rho, 3.0) if m == 1 else orc.assemble_elasticity(X, conn, rho, 3.0)).tolil()
        fixed = np.arange(m * dims[0])  # the first mesh line / its dofs
        b = np.random.default_rng(2).random(N * m)
        for f in fixed:  # symmetric elimination with zero values
            Kg[f, :] = 0.0
            Kg[:, f] = 0.0
            Kg[f, f] = 1.0
            b[f] = 0.0
        Kg = Kg.tocsr()
        ctx = SlabContext(X, conn, group=dist.group.WORLD)
        part = ctx.part
        gb, ge = ctx.owned_nodes
        lb, le = part.own_range
        gid = np.asarray(part.node_gid)
        ghost = np.concatenate([gid[:lb], gid[le:]])
        ex = HaloExchange(ghost, ctx.ranges, m, rank, group=dist.group.WORLD)
        rows = Kg[gb * m: ge * m]
        # the slab reads owned + ghost columns only
        used = np.unique(rows.indices // m)
        ok &= bool(np.all(np.isin(used, np.concatenate([np.arange(gb, ge), ghost]))))
        A = torch.sparse_csr_tensor(torch.from_numpy(rows.indptr.astype(np.int64)), torch.from_numpy(rows.indices.astype(np.int64)),
                                    torch.from_numpy(rows.data), size=rows.shape)
        dinv = torch.from_numpy(1.0 / Kg.diagonal()[gb * m: ge * m])

        def matvec(p_full):  # poison what this rank neither owns nor asked for: the product must not read it
            masked = torch.full_like(p_full, float("nan"))
            for nodes in (np.arange(gb, ge), ghost):
                idx = torch.from_numpy((nodes[:, None] * m + np.arange(m)).ravel())
                masked[idx] = p_full[idx]
            masked = torch.nan_to_num(masked, nan=1e300)
            return (A @ masked.unsqueeze(1)).squeeze(1)

        x, iters, resid = cg_reference(matvec, ex, torch.from_numpy(b[gb * m: ge * m]), gb * m, N * m, dinv, rtol=1e-12,
                                       max_iter=4000)
        xs = ctx.gather_vector(x.numpy())
        if rank == 0:
            ref = spsolve(Kg.tocsc(), b)
            ok &= iters > 3 and np.max(np.abs(xs - ref)) <= 1e-8 * np.max(np.abs(ref))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write(str(int(flag.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_cg_host_logic(tmp_path, world):
    """Halo plan + exchange + the conjugate-gradient recurrence of pfg_cg_dist over row slabs (gloo, CPU tensors, a
    scipy slab standing in for the device SpMV): the gathered solution equals the direct solve of the global system,
    and no rank reads an entry it neither owns nor requested."""
    port = _free_port()
    mp.spawn(_cg_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "ok").read() == "1"
