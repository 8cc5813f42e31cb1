"""World-size-2 NCCL test of the multi-GPU path on real devices (skipped below two GPUs): each rank assembles its node-owned row
block on its own GPU (no collective), then the blocks are gathered INTO HBM of rank 0 by finetools.jl_b200/parallel.py
(gather_row_blocks_device: NCCL all-gather of the column counts, NCCL send/recv of the rowval / nzval slabs, the library's
plan / interleave / column-sort kernels) and compared entry by entry with the oracle's full matrix.  Covers ownership ordered by
rank (concatenation stays sorted) and a free-first numbering whose owned dof ranges interleave (per-column sort)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


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
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import finetools_jl_b200 as fe
    from conftest import isotropic_C
    ok = False
    try:
        fens, fes = fe.H8block(1.0, 1.0, 2.0, 14, 13, 22)
        u = fe.NodalField(np.zeros((fens.count(), 3)))
        if interleaved:
            fe.setebc(u, list(range(3, fens.count(), 41)), True, None, 0.0)  # fixed dofs are numbered last: owned dof ranges interleave
        fe.numberdofs(u)
        n = u.nalldofs()
        rule = fe.GaussRule(3, 2)
        C6 = isotropic_C()
        owner = fe.slab_owner(fens.count(), world)
        ctx = fe.GPUContext(rank, stream=torch.cuda.current_stream().cuda_stream)
        a = fe.SysmatAssemblerSparseGPU(0.0, ctx=ctx)
        a.setnomatrixresult(True)  # the block stays on the device
        femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
        fe.bilform_lin_elastic(femm, a, fe.NodalField(fens.xyz), u, fe.DeforModelRed3D, fe.DataCache(C6), raw=True, node_owner=owner, my_rank=rank)
        a.setnomatrixresult(False)
        ordered = fe.owned_ranges_ordered(u.dofnums, owner, world)
        assert ordered == (not interleaved)
        out = fe.gather_row_blocks_device(a, dist, dst=0, ordered=ordered)
        if rank == 0:
            from oracle import oracle as orc
            orc.build()
            I, J, V = orc.bilform_lin_elastic_coo("H8", fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, C6)
            fcp, frv, fnz = orc.sparse(I, J, V, n, n)
            gcp, grv, gnz = (t.cpu().numpy() for t in out)
            ok = bool(np.array_equal(gcp, fcp) and np.array_equal(grv, frv) and np.abs(gnz - fnz).max() <= 1e-12 * np.abs(fnz).max())
        else:
            ok = out is None
    finally:
        q.put((rank, bool(ok)))
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("interleaved", [False, True])
def test_gather_row_blocks_device_nccl(interleaved):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, interleaved, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    for p in procs:
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {0: True, 1: True}
