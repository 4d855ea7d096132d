"""Interface rows summed by a reduce over NCCL -- the north_star's partitioning variant.

Default multi-GPU assembly (partition.py) integrates the one ghost layer of elements on both neighbours and needs no
exchange.  Here every element is integrated on exactly ONE rank (the owner of its lowest-numbered node); a CSR row
of an interface node then receives partial sums from several ranks, and the non-owners ship theirs to the owner:

    setup (once per mesh)   rank r builds, per neighbour q, a thin "halo" handle over its own elements that touch
                            q-owned nodes, whose rows are exactly those nodes; it sends the halo pattern to q, and q
                            matches it against its own rows -> an index map into its CSR values;
    every assembly          r assembles its slab with the ghost elements masked out (pfg_mesh_set_element_mask) and its
                            halo handles; halo values travel with ncclSend/ncclRecv (torch.distributed P2P over
                            NVLink); q adds them at the mapped slots (pfg_add_indexed).  The cross-partition sum
                            completes the duplicate summation of the reference's coo->csr (pyfem.py:930-931).

Transport "p2p" fuses the halo assembly with its transfer: every rank keeps an inbox in symmetric memory
(torch.distributed._symmetric_memory, NVLink peer mappings), and a sender's halo handle assembles STRAIGHT INTO the
owner's inbox -- the tile kernel's TMA bulk stores (or the atomic kernel's red.global) target the peer's HBM through
NVLink / NVSwitch, chunk by chunk as they complete; no staging copy, no send/recv launch.  A device-side barrier on the
symmetric-memory signal pads orders "all halo kernels done" before the owners' indexed adds; every rank keeps two inboxes
and alternates between them, so that this one barrier per assembly also frees the inbox of the assembly before.

The planning below is numpy and runs anywhere (gloo test on CPU); `ReduceAssembler` is the device part.
"""
import numpy as np

from .partition import LocalMesh


def owner_of_nodes(node_gid, ranges):
    """Owner rank of every global node id, for contiguous ownership ranges [(b0,e0), (b1,e1), ...]."""
    starts = np.array([b for b, _ in ranges], dtype=np.int64)
    return np.searchsorted(starts, np.asarray(node_gid), side="right") - 1


class HaloSend:
    """One neighbour's share: my own elements that touch nodes owned by rank `dest`, as a self-contained mesh whose
    owned rows are exactly those nodes."""

    def __init__(self, dest, X, conn, own_range, node_gid, local_nodes, elem_local):
        self.dest = dest
        self.X, self.conn, self.own_range, self.node_gid = X, conn, own_range, node_gid
        self.local_nodes = local_nodes  # index of the halo mesh's nodes in the rank's local numbering (nodal fields)
        self.elem_local = elem_local    # index of the halo mesh's elements in the rank's local element list


class HaloPlan:
    """Which local elements this rank integrates, and what it owes to / expects from its neighbours."""

    def __init__(self, part: LocalMesh, ranges):
        self.part, self.ranges = part, list(ranges)
        rank = part.rank
        conn_gid = part.node_gid[part.conn]
        elem_owner = owner_of_nodes(conn_gid.min(axis=1), ranges)
        self.mine = elem_owner == rank
        self.skip_mask = (~self.mine).astype(np.uint8)           # ghost elements: in the pattern, not integrated
        self.recv_from = sorted(set(int(q) for q in np.unique(elem_owner[~self.mine])))
        node_owner = owner_of_nodes(conn_gid, ranges)             # (nelems, nne)
        self.sends = []
        for q in sorted(set(int(v) for v in np.unique(node_owner[self.mine])) - {rank}):
            sel = np.nonzero(self.mine & (node_owner == q).any(axis=1))[0]
            sub = part.conn[sel]
            local_nodes = np.unique(sub)
            conn_h = np.searchsorted(local_nodes, sub).astype(np.int64)
            gid_h = part.node_gid[local_nodes]
            b, e = ranges[q]
            lb, le = np.searchsorted(gid_h, [b, e])
            self.sends.append(HaloSend(q, np.ascontiguousarray(part.X[local_nodes]), conn_h, (int(lb), int(le)),
                                       gid_h.astype(np.int64), local_nodes, sel))


def match_rows(own_row_gid0, m, indptr, indices, row_gids, h_indptr, h_indices):
    """Slots of a neighbour's halo entries inside the owner's CSR values.

    own_row_gid0: global id of the owner's first node; indptr / indices: the owner's slab pattern (global columns);
    row_gids: global node ids of the halo rows; h_indptr / h_indices: the halo pattern (m dof rows per node, global
    columns).  Every halo entry must exist in the owner's row (the owner's pattern holds the ghost elements)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    h_indptr = np.asarray(h_indptr, dtype=np.int64)
    out = np.empty(int(h_indptr[-1]), dtype=np.int64)
    dof_rows = (np.repeat(np.asarray(row_gids, dtype=np.int64) - own_row_gid0, m) * m +
                np.tile(np.arange(m, dtype=np.int64), len(row_gids)))
    for i, r in enumerate(dof_rows):
        seg = np.asarray(indices[indptr[r]:indptr[r + 1]], dtype=np.int64)
        want = np.asarray(h_indices[h_indptr[i]:h_indptr[i + 1]], dtype=np.int64)
        pos = np.searchsorted(seg, want)
        if pos.size and (pos.max() >= seg.size or not np.array_equal(seg[pos], want)):
            raise ValueError("halo pattern is not contained in the owner's rows")
        out[h_indptr[i]:h_indptr[i + 1]] = indptr[r] + pos
    return out


def _even(n):
    return n + (n & 1)


def inbox_layout(send_table, m):
    """Layout of the symmetric-memory inboxes of the p2p transport.  send_table[q, d] = (node rows, nnz) that rank q
    sends to rank d.  Rank q's block in d's inbox holds its nnz CSR values, then one vector entry per dof row; blocks
    follow each other in sender order.  Returns (doubles per inbox -- the same on every rank, symmetric allocations
    must agree -- and offset[q][d] of q's block inside d's inbox)."""
    T = np.asarray(send_table)
    size = T.shape[0]
    # every block starts on a 16-byte boundary (an even number of doubles): the halo handles write their values with
    # cp.async.bulk stores, which need 16-byte aligned global addresses
    block = [[0 if q == d else _even(_even(int(T[q, d, 1])) + int(T[q, d, 0]) * m) for d in range(size)]
             for q in range(size)]
    offset = [[sum(block[qq][d] for qq in range(q)) for d in range(size)] for q in range(size)]
    n = max(1, max(sum(block[q][d] for q in range(size)) for d in range(size)))
    return n, offset


class ReduceAssembler:
    """Device side: the rank's slab handle with masked ghost elements + one halo handle per neighbour, and the
    exchange.  Methods mirror DeviceMesh.assemble_*; nodal fields are given in the rank's local numbering."""

    def __init__(self, part: LocalMesh, ndof_per_node, ranges, device=None, group=None, transport="nccl",
                 nelems_global=None):
        import torch
        import torch.distributed as dist
        from .engine import DeviceMesh
        if transport not in ("nccl", "p2p"):
            raise ValueError(f"unknown halo transport {transport!r}")
        self.torch, self.dist, self.group, self.transport = torch, dist, group, transport
        self.part, self.m = part, int(ndof_per_node)
        self.plan = HaloPlan(part, ranges)
        self.mesh = DeviceMesh(part.X, part.conn, self.m, device=device, own_range=part.own_range,
                               node_gid=part.node_gid, ncols_nodes=part.nnodes_global,
                               nelems_global=nelems_global if nelems_global is not None else part.nelems_global)
        self.device = self.mesh.device
        self.mesh.set_element_mask(self.plan.skip_mask)
        self.halo = [(s, DeviceMesh(s.X, s.conn, self.m, device=self.device, own_range=s.own_range, node_gid=s.node_gid,
                                    ncols_nodes=part.nnodes_global),
                      torch.as_tensor(s.local_nodes, device=self.device)) for s in self.plan.sends]
        self._exchange_patterns()
        if transport == "p2p":
            self._setup_p2p()

    # ---- p2p transport: inboxes in symmetric memory -----------------------------------------------------------------
    def _setup_p2p(self):
        import torch.distributed._symmetric_memory as symm_mem
        torch, rank, size, m = self.torch, self.part.rank, self.part.size, self.m
        n, off_table = inbox_layout(self._send_table, m)

        def offset(q, d):
            return off_table[q][d]

        # TWO inboxes per rank, used alternately: the sender of assembly k writes the half its owner consumed in
        # assembly k - 2, and the one barrier of assembly k - 1 (which the owner joins only after enqueueing those
        # adds) already orders the two -- so an assembly needs a single device-side barrier, not two
        n = _even(n)
        self.inbox = symm_mem.empty(2 * n, dtype=torch.float64, device=self.device)  # same size on every rank
        self.symm = symm_mem.rendezvous(self.inbox, self.group if self.group is not None else self.dist.group.WORLD)
        self.peer_out, self.recv_p2p, self._phase = [[], []], [[], []], 0
        for par in (0, 1):
            for s, hm, _ in self.halo:
                off = par * n + offset(rank, s.dest)
                self.peer_out[par].append((self.symm.get_buffer(s.dest, (hm.nnz,), torch.float64, off),
                                           self.symm.get_buffer(s.dest, (hm.nrows,), torch.float64, off + _even(hm.nnz))))
            for q, slots, rows, _, _ in self.recv:
                off = par * n + offset(q, rank)
                v0 = off + _even(len(slots))
                self.recv_p2p[par].append((q, slots, rows, self.inbox[off:off + len(slots)], self.inbox[v0:v0 + len(rows)]))
        self.symm.barrier()

    # ---- setup: halo patterns travel to the owners, owners build their index maps --------------------------------
    def _p2p(self, ops):
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()

    def _exchange_patterns(self):
        torch, dist = self.torch, self.dist
        rank, size = self.part.rank, self.part.size
        # sizes first (tiny, one all_gather), then the pattern arrays point to point
        mine = torch.zeros((size, 2), dtype=torch.int64, device=self.device)
        for s, hm, _ in self.halo:
            mine[s.dest, 0], mine[s.dest, 1] = hm.nrows // self.m, hm.nnz
        table = [torch.zeros_like(mine) for _ in range(size)]
        dist.all_gather(table, mine, group=self.group)
        self._send_table = torch.stack(table).cpu().numpy()
        ops, send_keep, recv_bufs = [], [], {}
        for s, hm, _ in self.halo:
            rows = torch.as_tensor(s.node_gid[s.own_range[0]:s.own_range[1]], device=self.device)
            indptr, indices = hm.pattern(idx_bytes=8)
            for t in (rows, indptr, indices):
                send_keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, s.dest, group=self.group))
        for q in range(size):
            nrows, nnz = int(table[q][rank, 0]), int(table[q][rank, 1])
            if q == rank or nrows == 0:
                continue
            bufs = (torch.empty(nrows, dtype=torch.int64, device=self.device),
                    torch.empty(nrows * self.m + 1, dtype=torch.int64, device=self.device),
                    torch.empty(nnz, dtype=torch.int64, device=self.device))
            recv_bufs[q] = bufs
            for t in bufs:
                ops.append(dist.P2POp(dist.irecv, t, q, group=self.group))
        self._p2p(ops)
        torch.cuda.synchronize(self.device)
        # owner side: slots of every received entry, and of every received vector row
        self.recv = []
        if recv_bufs:
            indptr, indices = self.mesh.pattern()
            ip = indptr.cpu().numpy().astype(np.int64)
            gid0 = int(self.part.node_gid[self.part.own_range[0]])
        for q, (rows, h_indptr, h_indices) in sorted(recv_bufs.items()):
            rows_np = rows.cpu().numpy()
            dof_rows = (np.repeat(rows_np - gid0, self.m) * self.m + np.tile(np.arange(self.m), len(rows_np)))
            # fetch only the owner's interface rows from the device pattern
            lo, hi = ip[dof_rows], ip[dof_rows + 1]
            lens = (hi - lo).astype(np.int64)
            flat = np.repeat(lo.astype(np.int64) - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(int(lens.sum()))
            seg = indices[torch.as_tensor(flat, device=self.device)].cpu().numpy()
            sub_indptr = np.concatenate(([0], np.cumsum(lens)))
            rel = match_rows(0, 1, sub_indptr, seg, np.arange(len(dof_rows)), h_indptr.cpu().numpy(), h_indices.cpu().numpy())
            slots = flat[rel]
            self.recv.append((q, torch.as_tensor(slots, device=self.device),
                              torch.as_tensor(dof_rows.astype(np.int64), device=self.device),
                              torch.empty(len(slots), dtype=torch.float64, device=self.device),
                              torch.empty(len(dof_rows), dtype=torch.float64, device=self.device)))

    # ---- every assembly ---------------------------------------------------------------------------------------------
    def _halo_out(self, i):
        """(values, vector) outputs of halo handle i: the owner's inbox for the p2p transport, else fresh tensors."""
        return self.peer_out[self._phase & 1][i] if self.transport == "p2p" else (None, None)

    def _before_halo(self):
        pass  # (the p2p inboxes alternate: no barrier is needed before the halo kernels, see _setup_p2p)

    def _reduce(self, main_vals, halo_vals, main_vec=None, halo_vecs=None):
        dist = self.dist
        if self.transport == "p2p":
            self.symm.barrier()  # every rank's halo kernels have written their peers' inboxes
            for q, slots, rows, vbuf, rbuf in self.recv_p2p[self._phase & 1]:
                if halo_vals is not None:
                    self.mesh.add_indexed(main_vals, slots, vbuf)
                if halo_vecs is not None:
                    self.mesh.add_indexed(main_vec, rows, rbuf)
            self._phase += 1
            return
        ops = []
        for i, (s, _, _) in enumerate(self.halo):
            if halo_vals is not None:
                ops.append(dist.P2POp(dist.isend, halo_vals[i], s.dest, group=self.group))
            if halo_vecs is not None:
                ops.append(dist.P2POp(dist.isend, halo_vecs[i], s.dest, group=self.group))
        for q, _, _, vbuf, rbuf in self.recv:
            if halo_vals is not None:
                ops.append(dist.P2POp(dist.irecv, vbuf, q, group=self.group))
            if halo_vecs is not None:
                ops.append(dist.P2POp(dist.irecv, rbuf, q, group=self.group))
        self._p2p(ops)
        for q, slots, rows, vbuf, rbuf in self.recv:
            if halo_vals is not None:
                self.mesh.add_indexed(main_vals, slots, vbuf)
            if halo_vecs is not None:
                self.mesh.add_indexed(main_vec, rows, rbuf)

    def _field(self, f, local_nodes):
        if f is None or np.ndim(f) == 0:  # scalars, numpy scalars and 0-d arrays / tensors are constant fields
            return f if f is None else float(f)
        return self.torch.as_tensor(f, device=self.device)[local_nodes]

    def assemble_elasticity(self, rho=1.0, p=0.0, E=10.0, nu=0.3, out=None, mode="auto"):
        self._before_halo()
        hv = [hm.assemble_elasticity(self._field(rho, ln), p, E, nu, out=self._halo_out(i)[0], mode=mode)
              for i, (_, hm, ln) in enumerate(self.halo)]
        vals = self.mesh.assemble_elasticity(rho, p, E, nu, out=out, mode=mode)  # overlaps the halo stores' flight
        self._reduce(vals, hv)
        return vals

    def assemble_poisson(self, rho=1.0, p=0.0, out=None, mode="auto"):
        self._before_halo()
        hv = [hm.assemble_poisson(self._field(rho, ln), p, out=self._halo_out(i)[0], mode=mode)
              for i, (_, hm, ln) in enumerate(self.halo)]
        vals = self.mesh.assemble_poisson(rho, p, out=out, mode=mode)
        self._reduce(vals, hv)
        return vals

    def assemble_nlpoisson(self, xdv, u, mode="auto"):
        self._before_halo()
        hk, hr = [], []
        for i, (_, hm, ln) in enumerate(self.halo):
            ok, orr = self._halo_out(i)
            k, r = hm.assemble_nlpoisson(xdv, self._field(u, ln), out_K=ok, out_res=orr, mode=mode)
            hk.append(k)
            hr.append(r)
        K, res = self.mesh.assemble_nlpoisson(xdv, u, mode=mode)
        self._reduce(K, hk, res, hr)
        return K, res

    def assemble_helmholtz(self, r0, out_K=None, out_R=None, mode="auto"):
        """K and R of the Helmholtz filter (pyfem.py:2084-2097): two value arrays per handle.  The halo handles keep
        their outputs local and ship them with NCCL send/recv for either transport (an inbox holds one matrix)."""
        hk, hr = [], []
        for _, hm, _ in self.halo:
            k, r = hm.assemble_helmholtz(r0, mode=mode)
            hk.append(k)
            hr.append(r)
        K, R = self.mesh.assemble_helmholtz(r0, out_K=out_K, out_R=out_R, mode=mode)
        transport, self.transport = self.transport, "nccl"
        try:
            self._reduce(K, hk)
            self._reduce(R, hr)
        finally:
            self.transport = transport
        return K, R
