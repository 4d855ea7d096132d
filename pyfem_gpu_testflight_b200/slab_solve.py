"""Krylov solves on a matrix distributed by row slabs (one rank = one GPU = one slab): the multi-GPU form of
Assembler._solve_linear_system(K, rhs, method="cg") (pyfem.py:2403-2423) and of the solves inside compliance
(pyfem.py:1050-1068, 1814-1828).

The device work is the library's (pfg_cg_dist: SpMV on the rank's slab, fused vector kernels, fixed-order dot
products); what crosses ranks is (a) the ghost entries of the search direction -- the columns of a rank's rows that
belong to other ranks' nodes, one mesh layer either side of a slab -- exchanged with batched send / recv before every
product, and (b) three all-reduced scalars per iteration.  `HaloExchange` is host logic over torch.distributed only
(it runs under gloo on CPU tensors as well, tests/test_partition_gloo.py); `SlabKrylov` binds it to a slab handle.
"""
import ctypes

import numpy as np

from . import _lib


class HaloExchange:
    """Who needs which entries of a row-distributed dof vector.

    ghost_gid  global ids of the nodes this rank reads but does not own (sorted)
    ranges     [(begin, end)] global node range of every rank
    m          dofs per node
    The lists are agreed on once (one all_gather_object of the requests); `refresh(x_full)` then copies, for every
    neighbour, the owned entries it asked for into a send buffer and scatters what arrives into the ghost entries.
    """

    def __init__(self, ghost_gid, ranges, m, rank, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.group, self.rank, self.m = group, int(rank), int(m)
        ghost_gid = np.asarray(ghost_gid, dtype=np.int64)
        begins = np.array([b for b, _ in ranges], dtype=np.int64)
        owner = np.searchsorted(begins, ghost_gid, side="right") - 1
        need = {int(q): ghost_gid[owner == q] for q in np.unique(owner)}  # what this rank asks of rank q
        if self.rank in need:
            raise ValueError("ghost nodes must not include owned nodes")
        everyone = [None] * len(ranges)
        dist.all_gather_object(everyone, need, group=group)
        dofs = lambda nodes: (np.asarray(nodes, dtype=np.int64)[:, None] * self.m + np.arange(self.m)).ravel()
        self.peers = sorted(set(need) | {q for q, asks in enumerate(everyone) if self.rank in asks})
        dev = torch.device("cpu") if device is None else torch.device(device)
        # gloo moves host memory only: device vectors are staged through the host then (two processes sharing one
        # GPU in the tests; under NCCL everything stays on the device and on the stream)
        self.staged = dev.type == "cuda" and dist.get_backend(group) == "gloo"
        self.send_idx, self.recv_idx, self.send_buf, self.recv_buf = {}, {}, {}, {}
        for q in self.peers:
            s = dofs(everyone[q].get(self.rank, np.empty(0, dtype=np.int64)))
            r = dofs(need.get(q, np.empty(0, dtype=np.int64)))
            self.send_idx[q] = torch.as_tensor(s, device=dev)
            self.recv_idx[q] = torch.as_tensor(r, device=dev)
            self.send_buf[q] = torch.empty(len(s), dtype=torch.float64, device=dev)
            self.recv_buf[q] = torch.empty(len(r), dtype=torch.float64, device=dev)
        self.bytes_per_refresh = 8 * sum(len(v) for v in self.send_idx.values())

    def _peer(self, q):
        import torch.distributed as dist
        return q if self.group is None else dist.get_global_rank(self.group, q)

    def refresh(self, x_full):
        """x_full: global-length dof vector whose owned entries are current; on return the ghost entries are too."""
        import torch
        import torch.distributed as dist
        ops, landing = [], {}
        for q in self.peers:
            if len(self.send_idx[q]):
                torch.index_select(x_full, 0, self.send_idx[q], out=self.send_buf[q])
                out = self.send_buf[q].cpu() if self.staged else self.send_buf[q]
                ops.append(dist.P2POp(dist.isend, out, self._peer(q), group=self.group))
            if len(self.recv_idx[q]):
                landing[q] = torch.empty(len(self.recv_idx[q]), dtype=torch.float64) if self.staged else self.recv_buf[q]
                ops.append(dist.P2POp(dist.irecv, landing[q], self._peer(q), group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for q, buf in landing.items():
            x_full.index_copy_(0, self.recv_idx[q], buf.to(x_full.device) if self.staged else buf)
        return x_full

    def all_reduce(self, t):
        import torch.distributed as dist
        if self.staged:
            host = t.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=self.group)
            t.copy_(host)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class SlabKrylov:
    """Jacobi-preconditioned conjugate gradients (symmetric systems) and BiCGStab (the Newton Jacobian) on the device
    CSR of a row-slab model (models.py, `group=`)."""

    def __init__(self, mesh, part, ranges, rank, group=None):
        import torch
        self.mesh, self.group = mesh, group
        m = mesh.ndof_per_node
        lb, le = part.own_range
        gid = np.asarray(part.node_gid, dtype=np.int64)
        ghost = np.concatenate([gid[:lb], gid[le:]])
        self.exchange = HaloExchange(ghost, ranges, m, rank, group=group, device=mesh.device)
        self.row0 = int(part.owned_global_range[0]) * m
        self.x_full2 = torch.zeros(2 * mesh.ncols, dtype=torch.float64, device=mesh.device)  # BiCGStab uses both halves
        self.x_full = self.x_full2[:mesh.ncols]
        self.scal = torch.zeros(8, dtype=torch.float64, device=mesh.device)
        self._error = None

        def reduce_cb(_user, offset, count):
            try:
                self.exchange.all_reduce(self.scal[offset:offset + count])
                return 0
            except BaseException as exc:  # must not propagate through the C frames
                self._error = exc
                return 1

        def halo_cb(_user):
            try:
                self.exchange.refresh(self.x_full)
                return 0
            except BaseException as exc:
                self._error = exc
                return 1

        def halo2_cb(_user, which):
            try:
                n = self.mesh.ncols
                self.exchange.refresh(self.x_full2[which * n:(which + 1) * n])
                return 0
            except BaseException as exc:
                self._error = exc
                return 1

        # (the ctypes thunks are kept alive with self)
        self._reduce_cb, self._halo_cb, self._halo2_cb = _lib.REDUCE_FN(reduce_cb), _lib.HALO_FN(halo_cb), _lib.HALO2_FN(halo2_cb)

    def cg(self, vals, b_owned, x0=None, rtol=1e-8, atol=0.0, max_iter=None, check_every=16, graph=None):
        """(x_owned, iterations, |r| over all ranks); RuntimeError like the reference when max_iter is reached.

        graph=True (the default under NCCL): after two eager iterations the next two -- kernels, halo send / recv and
        all-reduces alike -- are captured in a CUDA graph and replayed, so the host enqueues one graph launch per two
        iterations instead of ~25 operations; the residual norm is read every `check_every` iterations as before.
        graph=False (and always under gloo, whose staged exchange syncs with the host): one C call runs the loop."""
        import torch
        mesh = self.mesh
        b = mesh._dev_f64(b_owned, mesh.nrows, "b")
        x = torch.empty_like(b) if x0 is None else mesh._dev_f64(x0, mesh.nrows, "x0").clone()
        max_iter = 10 * mesh.ncols if max_iter is None else int(max_iter)  # scipy's default, on the global size
        check_every = 16 if check_every <= 0 else int(check_every)
        if graph is None:
            graph = not self.exchange.staged and self._graph_ok
        self._error = None
        if not graph:
            iters, resid = ctypes.c_int(0), ctypes.c_double(0.0)
            with torch.cuda.device(mesh.device):
                st = mesh._lib.pfg_cg_dist(mesh._handle, vals.data_ptr(), b.data_ptr(), x.data_ptr(),
                                           1 if x0 is None else 0, self.x_full.data_ptr(), self.scal.data_ptr(), self.row0,
                                           float(rtol), float(atol), max_iter, check_every, self._reduce_cb,
                                           self._halo_cb, None, ctypes.byref(iters), ctypes.byref(resid), mesh._stream())
            if self._error is not None:
                raise self._error
            if st == _lib.PFG_ERR_NOCONV:
                raise RuntimeError(f"cg failed with code {iters.value}")
            _lib.check(st)
            return x, int(iters.value), float(resid.value)

        def steps(first, count):
            _lib.check(mesh._lib.pfg_cg_dist_steps(mesh._handle, vals.data_ptr(), x.data_ptr(), self.x_full.data_ptr(),
                                                   self.scal.data_ptr(), self.row0, first, count, self._reduce_cb,
                                                   self._halo_cb, None, mesh._stream()))
            if self._error is not None:
                raise self._error

        with torch.cuda.device(mesh.device):
            _lib.check(mesh._lib.pfg_cg_dist_begin(mesh._handle, vals.data_ptr(), b.data_ptr(), x.data_ptr(),
                                                   1 if x0 is None else 0, self.x_full.data_ptr(), self.scal.data_ptr(),
                                                   self.row0, self._reduce_cb, self._halo_cb, None, mesh._stream()))
            if self._error is not None:
                raise self._error
            host = self.scal.tolist()
            target = max(rtol * host[4] ** 0.5, atol)  # scipy's cg: |r| <= max(rtol |b|, atol), global norms
            rr, it, pair = host[2], 0, None
            while rr ** 0.5 > target and it < max_iter:
                batch = min(check_every, max_iter - it)
                done = 0
                if it == 0 and batch >= 2:  # connections and buffers are set up outside the capture
                    steps(0, 2)
                    done = 2
                if (it + done) % 2 == 1 and batch > done:  # (an odd check_every: the graph starts on even iterations)
                    steps(it + done, 1)
                    done += 1
                if pair is None and self._graph_ok and batch - done >= 2:
                    pair = self._capture(lambda: steps(it + done, 2))
                    if pair is not None:
                        done += 2
                while pair is not None and batch - done >= 2:
                    pair.replay()
                    done += 2
                if batch > done:
                    steps(it + done, batch - done)
                it += batch
                rr = float(self.scal[2])
                if rr != rr:
                    raise ValueError(f"pfg_cg_dist: the residual became NaN after {it} iterations "
                                     "(matrix not positive definite?)")
        if rr ** 0.5 > target:
            raise RuntimeError(f"cg failed with code {it}")
        return x, it, rr ** 0.5

    _graph_ok = True

    def _capture(self, enqueue):
        """Run `enqueue` under CUDA-graph capture on a side stream (the work it enqueues also EXECUTES once, through the
        first replay).  None if this torch / NCCL combination cannot capture the collectives: the caller goes on
        eagerly and capture is not tried again."""
        import torch
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.mesh.device)
        side.wait_stream(torch.cuda.current_stream(self.mesh.device))
        try:
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    enqueue()
        except Exception:  # noqa: BLE001 -- any capture failure means "no graphs here"
            self._graph_ok = False
            self._error = None
            torch.cuda.synchronize(self.mesh.device)
            return None
        torch.cuda.current_stream(self.mesh.device).wait_stream(side)
        g.replay()
        return g

    solve = cg

    def bicgstab(self, vals, b_owned, rtol=1e-8, atol=0.0, max_iter=None, check_every=8):
        """Non-symmetric systems (x0 = 0): (x_owned, iterations, |r| over all ranks)."""
        import torch
        mesh = self.mesh
        b = mesh._dev_f64(b_owned, mesh.nrows, "b")
        x = torch.empty_like(b)
        iters, resid = ctypes.c_int(0), ctypes.c_double(0.0)
        max_iter = 10 * mesh.ncols if max_iter is None else int(max_iter)
        self._error = None
        with torch.cuda.device(mesh.device):
            st = mesh._lib.pfg_bicgstab_dist(mesh._handle, vals.data_ptr(), b.data_ptr(), x.data_ptr(),
                                             self.x_full2.data_ptr(), self.scal.data_ptr(), self.row0, float(rtol),
                                             float(atol), max_iter, int(check_every), self._reduce_cb, self._halo2_cb,
                                             None, ctypes.byref(iters), ctypes.byref(resid), mesh._stream())
        if self._error is not None:
            raise self._error
        if st == _lib.PFG_ERR_NOCONV:
            raise RuntimeError(f"bicgstab failed with code {iters.value}")
        _lib.check(st)
        return x, int(iters.value), float(resid.value)


SlabCG = SlabKrylov  # first name of the class
