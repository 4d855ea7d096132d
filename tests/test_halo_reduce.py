"""The reduce variant of the multi-GPU assembly (pyfem_gpu_testflight_b200/halo.py): every element integrated on one
rank, interface rows completed by the neighbours' contributions.

CPU part (no GPU): the planning (element ownership, halo sub-meshes, slot matching) with the numpy oracle standing in
for the device handles -- the summed slabs must reproduce the global CSR.  GPU part: the same with the real handles
(masked ghost elements, halo handles, pfg_add_indexed), ranks emulated one after another on one GPU; the NCCL
exchange itself runs in `bench.py --halo reduce` under torchrun.
"""
import numpy as np
import pytest

import pyfem_oracle as orc
from parity import VAL_TOL, assert_values_close


def _case(three_d):
    if three_d:
        X, conn = orc.structured_mesh(7, 6, 9)
    else:
        X, conn = orc.structured_mesh(23, 31)
    rng = np.random.default_rng(7)
    X = X + rng.uniform(-0.004, 0.004, size=X.shape)
    return X, conn


def _slab_pattern(part, m):
    """Owner pattern of the rank's slab (all local elements, ghost layer included), global columns."""
    Kl = orc.assemble_elasticity(part.X, part.conn)
    lb, le = part.own_range
    rows = Kl[lb * m: le * m]
    gcols = m * part.node_gid[rows.indices // m] + rows.indices % m
    return rows.indptr.astype(np.int64), gcols.astype(np.int64)


def _aligned(indptr, gcols, K_rows, gcols_rows):
    """Values of a row-compatible matrix with a sub-pattern, laid out on (indptr, gcols)."""
    out = np.zeros(len(gcols))
    for r in range(len(indptr) - 1):
        seg = gcols[indptr[r]:indptr[r + 1]]
        a, b = K_rows.indptr[r], K_rows.indptr[r + 1]
        pos = np.searchsorted(seg, gcols_rows[a:b])
        assert np.array_equal(seg[pos], gcols_rows[a:b])
        out[indptr[r] + pos] = K_rows.data[a:b]
    return out


@pytest.mark.parametrize("three_d", [False, True])
@pytest.mark.parametrize("size", [2, 3])
def test_halo_plan_reproduces_global_matrix_cpu(three_d, size):
    from pyfem_gpu_testflight_b200.halo import HaloPlan, match_rows
    from pyfem_gpu_testflight_b200.partition import concat_slabs, partition_mesh, split_range
    X, conn = _case(three_d)
    m = X.shape[1]
    Kg = orc.assemble_elasticity(X, conn)
    ranges = split_range(X.shape[0], size)
    parts = [partition_mesh(X, conn, r, size) for r in range(size)]
    plans = [HaloPlan(p, ranges) for p in parts]
    assert sum(int(pl.mine.sum()) for pl in plans) == conn.shape[0]  # every element integrated exactly once
    slabs = []
    for r, (part, plan) in enumerate(zip(parts, plans)):
        indptr, gcols = _slab_pattern(part, m)
        lb, le = part.own_range
        # main handle: only my elements, on the owner pattern
        Km = orc.assemble_elasticity(part.X, part.conn[plan.mine])
        Km = Km[lb * m: le * m] if Km.shape[0] >= le * m else None
        vals = np.zeros(len(gcols))
        if Km is not None:
            gc = m * part.node_gid[Km.indices // m] + Km.indices % m
            vals = _aligned(indptr, gcols, Km, gc)
        # contributions shipped by the neighbours
        for q in plan.recv_from:
            (s,) = [s for s in plans[q].sends if s.dest == r]
            Kh = orc.assemble_elasticity(s.X, s.conn)
            hb, he = s.own_range
            rows = Kh[hb * m: he * m]
            h_cols = m * s.node_gid[rows.indices // m] + rows.indices % m
            slots = match_rows(int(part.node_gid[lb]), m, indptr, gcols, s.node_gid[hb:he], rows.indptr, h_cols)
            assert len(np.unique(slots)) == len(slots)
            vals[slots] += rows.data
        slabs.append((indptr, gcols, vals))
    K = concat_slabs(slabs, Kg.shape[1])
    assert np.array_equal(K.indptr, Kg.indptr) and np.array_equal(K.indices, Kg.indices)
    assert_values_close(K.data, Kg.data, VAL_TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("three_d", [False, True])
@pytest.mark.parametrize("mode", ["gather", "atomic"])
def test_halo_reduce_device_emulated_ranks(three_d, mode):
    import torch
    import pyfem_gpu_testflight_b200 as pf
    from pyfem_gpu_testflight_b200.halo import HaloPlan, match_rows
    from pyfem_gpu_testflight_b200.partition import concat_slabs, partition_mesh, split_range
    X, conn = _case(three_d)
    m = X.shape[1]
    size = 3
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    Kg = orc.assemble_elasticity(X, conn, rho, 3.0)
    ranges = split_range(X.shape[0], size)
    parts = [partition_mesh(X, conn, r, size) for r in range(size)]
    plans = [HaloPlan(p, ranges) for p in parts]
    halo_out = {}
    for q, plan in enumerate(plans):  # what every rank would send
        for s in plan.sends:
            hm = pf.DeviceMesh(s.X, s.conn, m, own_range=s.own_range, node_gid=s.node_gid, ncols_nodes=X.shape[0])
            v = hm.assemble_elasticity(rho[parts[q].node_gid][s.local_nodes], 3.0, mode=mode)
            ip, ix = hm.pattern_host()
            halo_out[(q, s.dest)] = (s.node_gid[s.own_range[0]:s.own_range[1]], ip, ix, v)
    slabs = []
    for r, (part, plan) in enumerate(zip(parts, plans)):
        mesh = pf.DeviceMesh(part.X, part.conn, m, own_range=part.own_range, node_gid=part.node_gid,
                             ncols_nodes=part.nnodes_global)
        mesh.set_element_mask(plan.skip_mask)
        vals = mesh.assemble_elasticity(rho[part.node_gid], 3.0, mode=mode)
        indptr, indices = mesh.pattern_host()
        for q in plan.recv_from:
            rows, ip, ix, v = halo_out[(q, r)]
            slots = match_rows(int(part.node_gid[part.own_range[0]]), m, indptr, indices, rows, ip, ix)
            mesh.add_indexed(vals, torch.as_tensor(slots, device="cuda"), v)
        # clearing the mask restores the ghost-layer variant on the same handle
        slabs.append((indptr, indices, vals.cpu().numpy()))
        mesh.set_element_mask(None)
        full = mesh.assemble_elasticity(rho[part.node_gid], 3.0, mode=mode)
        assert_values_close(full.cpu().numpy(), slabs[-1][2], 1e-13)
    K = concat_slabs(slabs, Kg.shape[1])
    assert np.array_equal(K.indptr, Kg.indptr) and np.array_equal(K.indices, Kg.indices)
    assert_values_close(K.data, Kg.data, VAL_TOL)


def test_p2p_inbox_layout():
    """Blocks of the symmetric-memory inboxes: disjoint, in sender order, one common size, every block (and its
    vector part) on a 16-byte boundary even for odd nnz / row counts (the halo handles write with bulk stores)."""
    from pyfem_gpu_testflight_b200.halo import _even, inbox_layout
    T = np.zeros((3, 3, 2), dtype=np.int64)  # [sender, dest] = (node rows, nnz)
    T[0, 1], T[1, 0], T[1, 2], T[2, 1] = (5, 91), (5, 80), (7, 121), (7, 133)
    for m in (1, 2, 3):
        n, off = inbox_layout(T, m)
        blk = lambda q, d: int(_even(_even(T[q, d, 1]) + T[q, d, 0] * m))
        for d in range(3):
            spans = sorted((off[q][d], off[q][d] + blk(q, d)) for q in range(3) if q != d and T[q, d, 1])
            assert spans[0][0] == 0
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(a % 2 == 0 for a, _ in spans)
            assert spans[-1][1] <= n
        assert n == max(blk(0, 1) + blk(2, 1), blk(1, 0), blk(1, 2))
