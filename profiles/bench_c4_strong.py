#!/usr/bin/env python
"""Strong scaling of BASELINE config 4 (bilform_diffusion, 256^3 H8, kappa 3x3, 16.8 M elements) over N node-owned row blocks
(z-slabs, halo elements recomputed, no data-path collective).  Secondary benchmark, not the driver's contract (bench.py is).
  real:      python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
                 profiles/bench_c4_strong.py --steps K --warmup W
  emulated:  python profiles/bench_c4_strong.py --emulate N          (one GPU plays every rank in turn; value uses the slowest)
Timing: CUDA events around K fresh assemblies (pattern invalidated every step), barrier + synchronize on both sides, max over
ranks.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

KAPPA3 = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--edge", type=int, default=256)
    ap.add_argument("--emulate", type=int, default=0, help="play ranks 0..N-1 of an N-way partition one after the other on one GPU")
    args = ap.parse_args()
    import torch
    import finetools_jl_b200 as fe
    from finetools_jl_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    P = args.emulate if args.emulate > 0 else world
    n = args.edge
    fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
    u = fe.NodalField(np.zeros((fens.count(), 1)))
    fe.numberdofs(u)
    rule = fe.GaussRule(3, 2)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    owner = fe.slab_owner(fens.count(), P) if P > 1 else None
    own_stream = os.environ.get("C4S_OWN_STREAM") == "1"  # debugging knob: the library's own stream instead of torch's current one
    ctx = fe.GPUContext(local_rank) if own_stream else fe.GPUContext(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    use_async = os.environ.get("C4S_SYNC") != "1"         # debugging knob: blocking calls
    L = _lib.lib()
    Kf = np.asfortranarray(KAPPA3)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(r):
        a = fe.SysmatAssemblerSparseGPU(0.0, ctx=ctx)
        a.setnomatrixresult(True)
        fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(KAPPA3), raw=True, node_owner=owner, my_rank=r)
        dmesh = a._device_cache[id(fes)]
        dof = dmesh.dofmap(u)

        def step(fresh):
            if fresh:
                _lib.check(L.fegpu_pattern_invalidate(dof), ctx.handle)
            _lib.check(L.fegpu_bilform_diffusion(dmesh.handle, dof, 1, _lib.fptr(Kf), a.handle), ctx.handle)

        def timed(fresh):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ev0.record()
            for _ in range(args.steps):
                step(fresh)
            ev1.record()
            barrier()
            return ev0.elapsed_time(ev1) / args.steps

        ctx.set_async(use_async)
        for _ in range(args.warmup):
            step(True)
        fresh_ms = timed(True)
        step(False)
        cached_ms = timed(False)
        ctx.set_async(False)
        ctx.set_overlap(False)
        step(True)
        ctx.synchronize()
        ph = a.timings()
        ctx.set_overlap(True)
        _, _, nnz = a.sizes()
        for dm in a._device_cache.values():
            dm.destroy()
        return fresh_ms, cached_ms, ph, nnz

    if args.emulate > 0:
        res = [measure(r) for r in range(P)]
        fresh_ms, cached_ms = max(x[0] for x in res), max(x[1] for x in res)
        nnz = sum(x[3] for x in res)
        ph = res[int(np.argmax([x[0] for x in res]))][2]
        per_rank = [round(x[0], 4) for x in res]
    else:
        fresh_ms, cached_ms, ph, nnz = measure(rank)
        per_rank = None
        if dist is not None:
            t = torch.tensor([fresh_ms, cached_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fresh_ms, cached_ms = float(t[0].item()), float(t[1].item())
            t = torch.tensor([nnz], device="cuda", dtype=torch.int64)
            dist.all_reduce(t)
            nnz = int(t.item())
    if rank == 0:
        nel = fes.count()
        print(json.dumps({"metric": "elements/s assembled into CSC (H8 diffusion stiffness, fresh assembly incl. pattern build)",
                          "value": nel / (fresh_ms * 1e-3), "unit": "elements/s", "n_gpus": P, "emulated_on_one_gpu": args.emulate > 0,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": fresh_ms, "scaling": "strong",
                          "config": {"workload": "BASELINE configs[3]: bilform_diffusion, H8 block %d^3 (%d elements, %d nnz), GaussRule(3,2), "
                                                 "kappa 3x3; %d node-owned row blocks (z-slabs), halo recomputed" % (n, nel, nnz, P)},
                          "cached": {"value": nel / (cached_ms * 1e-3), "ms_per_step": cached_ms},
                          "phases_ms_slowest_rank": ph, "fresh_ms_per_rank": per_rank}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
