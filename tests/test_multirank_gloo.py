"""World-size-2 (and 3) gloo test of the multi-GPU host logic on CPU: slab ownership, per-rank row blocks and the gather of
the blocks into one CSC (finetools.jl_b200/parallel.py).  The per-rank block is produced by the oracle restricted to
the rank's rows -- exactly what the GPU path returns for a partitioned mesh (tests/test_gpu_parity.py checks that part)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, interleaved, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import finetools_jl_b200 as fe
    from oracle import oracle as orc
    from conftest import isotropic_C
    fens, fes = fe.H8block(1.0, 1.0, 2.0, 3, 3, 5)
    u = fe.NodalField(np.zeros((fens.count(), 3)))
    if interleaved:
        fe.setebc(u, [2, 11, 40], True, None, 0.0)  # fixed dofs are numbered last: owned dof ranges interleave
    fe.numberdofs(u)
    n = u.nalldofs()
    rule = fe.GaussRule(3, 2)
    I, J, V = orc.bilform_lin_elastic_coo("H8", fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, isotropic_C())
    owner = fe.slab_owner(fens.count(), world)
    owned_rows = np.zeros(n + 1, bool)
    owned_rows[u.dofnums[owner == rank].reshape(-1)] = True
    keep = owned_rows[I]
    # the halo rule: every triplet is kept by exactly one rank (the owner of its row node)
    cp, rv, nz = orc.sparse(I[keep], J[keep], V[keep], n, n)
    out = fe.gather_row_blocks(cp, rv, nz, n, n, dist=dist)
    if rank == 0:
        fcp, frv, fnz = orc.sparse(I, J, V, n, n)
        ok = np.array_equal(out[0], fcp) and np.array_equal(out[1], frv) and np.abs(out[2] - fnz).max() <= 1e-12 * np.abs(fnz).max()
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,interleaved", [(2, False), (3, False), (2, True)])
def test_gather_row_blocks_gloo(world, interleaved):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, interleaved, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    for p in procs:
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
