"""Two real GPUs, one process each (NCCL): the reduce variant of the multi-GPU assembly with both transports --
ncclSend/ncclRecv, and "p2p", where a sender's halo handle assembles straight into the owner's symmetric-memory inbox
over NVLink (halo.py).  Every rank's slab must equal the corresponding rows of the oracle's global matrix (pattern
bit-exact, values 1e-12), for the matrix and for the nonlinear-Poisson residual.  Skipped on boxes with one GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import pyfem_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, transport, out_dir):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pyfem_gpu_testflight_b200.halo import ReduceAssembler
    from pyfem_gpu_testflight_b200.partition import partition_mesh, split_range
    ok = True
    for dims in ((41, 37, None), (9, 8, 11)):
        X, conn = orc.structured_mesh(*dims)
        X = X + np.random.default_rng(3).uniform(-0.004, 0.004, size=X.shape)
        if dims[2] is None:
            X = (X - X.min(axis=0)) / (X.max(axis=0) - X.min(axis=0))
        m_el = X.shape[1]
        plane = dims[0] if dims[2] is None else dims[0] * dims[1]
        nslow = dims[1] if dims[2] is None else dims[2]
        ranges = [(b * plane, e * plane) for b, e in split_range(nslow, world)]
        part = partition_mesh(X, conn, rank, world, ranges)
        rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
        gb, ge = part.owned_global_range

        def slab_ok(vals, Kg, m):
            ref = Kg[gb * m: ge * m].data
            return np.max(np.abs(vals.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(Kg.data))

        ra = ReduceAssembler(part, m_el, ranges, device=torch.device("cuda", rank), transport=transport)
        Kg = orc.assemble_elasticity(X, conn, rho, 3.0)
        for _ in range(4):  # the two inboxes of the p2p transport alternate: each is reused
            ok &= slab_ok(ra.assemble_elasticity(rho[part.node_gid], 3.0), Kg, m_el)
        if dims[2] is None:
            rs = ReduceAssembler(part, 1, ranges, device=torch.device("cuda", rank), transport=transport)
            u = np.random.default_rng(5).random(X.shape[0]) - 0.4
            xdv = np.ones(10) / 10.0
            Kg, rg = orc.assemble_nlpoisson(X, conn, xdv, u)
            K, res = rs.assemble_nlpoisson(xdv, u[part.node_gid])
            ok &= slab_ok(K, Kg, 1)
            ok &= np.max(np.abs(res.cpu().numpy() - rg[gb:ge])) <= 1e-12 * np.max(np.abs(rg))
            # Helmholtz K and R (two value arrays per handle; their halo shares travel by send / recv)
            Hk, Hr = orc.assemble_helmholtz(X, conn, 0.05)
            for _ in range(2):
                Kh, Rh = rs.assemble_helmholtz(0.05)
                ok &= slab_ok(Kh, Hk, 1) and slab_ok(Rh, Hr, 1)
            K2, res2 = rs.assemble_nlpoisson(xdv, u[part.node_gid])  # the inboxes still work afterwards
            ok &= slab_ok(K2, Kg, 1)
    # the same through the model API: LinearElasticity(..., group=WORLD, halo=transport) is one rank of the partition,
    # compute_jacobian returns its row slab and gather() rebuilds the reference's global matrix on rank 0
    import pyfem_gpu_testflight_b200 as pf
    X, conn = orc.structured_mesh(41, 37)
    X = X + np.random.default_rng(3).uniform(-0.004, 0.004, size=X.shape)
    rho = 0.05 + 0.95 * np.random.default_rng(0).random(X.shape[0])
    q = pf.QuadratureBilinear2D()
    for halo in ("ghost", transport):
        model = pf.LinearElasticity(X, conn, [0, 1], None, {5: [0.0, -1.0]}, q, pf.BasisBilinear2D(q), p=3.0,
                                    group=dist.group.WORLD, halo=halo)
        Kall = model.gather(model.compute_jacobian(rho))
        if rank == 0:
            Kg = orc.assemble_elasticity(X, conn, rho, 3.0)
            ok &= Kall.indices.dtype == Kg.indices.dtype and np.array_equal(Kall.indptr, Kg.indptr)
            ok &= np.array_equal(Kall.indices, Kg.indices)
            ok &= np.max(np.abs(Kall.data - Kg.data)) <= 1e-12 * np.max(np.abs(Kg.data))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write(str(int(flag.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_reduce_variant_two_gpus(tmp_path, transport):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, transport, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
