#!/usr/bin/env python
"""Where does the end-to-end step (public API, host buffers) go at N ranks?  BASELINE config 4 (256^3 H8 diffusion) split into N
row blocks, one process per GPU:
  link   : every rank copies its block's bytes (12 B / nnz + colptr) device -> pinned host with plain cudaMemcpyAsync at the same
           time -- the ceiling the PCIe links + host memory give N concurrent DMA streams, no decode
  e2e    : the public call (coordinates in, CSC out into pinned arrays), fresh pattern every step, for the two transports of
           rowval: int32 on the link + widening by host threads (FEGPU_XFER_NARROW=1) and plain int64 DMA (=0)
  values : re-assembly on the cached pattern, nzval only
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/e2e_scaling.py
Prints one JSON line on rank 0 (times = max over ranks)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import finetools_jl_b200 as fe
    import bench

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps = int(os.environ.get("E2E_STEPS", "5"))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def wall(fn, n):
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        barrier()
        return allmax(time.perf_counter() - t0) / n

    out = {"n_gpus": world, "steps": steps, "host_cores": os.cpu_count()}
    results = {}
    for label, narrow in (("narrow_int32_plus_host_widen", "1"), ("plain_int64_dma", "0")):
        os.environ["FEGPU_XFER_NARROW"] = narrow  # read when the context builds its transport
        ctx = fe.GPUContext(local_rank, stream=torch.cuda.current_stream().cuda_stream)
        W = bench.Workload("c4", fe, ctx, world, rank, torch)
        m_, n_, nnz = W.a.sizes()
        pin = lambda cnt, dt: torch.empty(max(cnt, 1), dtype=dt, pin_memory=True).numpy()[:cnt]
        outarr = (pin(n_ + 1, torch.int64), pin(nnz, torch.int64), pin(nnz, torch.float64))

        def fresh():
            W.a.invalidate_patterns()
            W.api_call(W.a, outarr)

        def values_only():
            W.api_call(W.a, None, fetch=False)
            W.a.fetch_values(outarr[2])

        fresh()
        t_fresh = wall(fresh, steps)
        values_only()
        t_vals = wall(values_only, steps)
        results[label] = {"e2e_fresh_ms": t_fresh * 1e3, "e2e_fresh_elements_per_s": W.nelem / t_fresh,
                          "e2e_values_only_ms": t_vals * 1e3, "transfer_stats": ctx.transfer_stats()}
        if label.startswith("narrow"):
            # link ceiling: the same bytes by plain concurrent DMA (device buffers of the sizes the transport ships)
            dev_rv = torch.empty(nnz, dtype=torch.int32, device="cuda")
            dev_nz = torch.empty(nnz, dtype=torch.float64, device="cuda")
            dev_cp = torch.empty(n_ + 1, dtype=torch.int64, device="cuda")
            h_rv = torch.empty(nnz, dtype=torch.int32, pin_memory=True)
            h_nz = torch.empty(nnz, dtype=torch.float64, pin_memory=True)
            h_cp = torch.empty(n_ + 1, dtype=torch.int64, pin_memory=True)

            def link():
                h_cp.copy_(dev_cp, non_blocking=True)
                h_rv.copy_(dev_rv, non_blocking=True)
                h_nz.copy_(dev_nz, non_blocking=True)
                torch.cuda.synchronize()

            link()
            t_link = wall(link, steps)
            nbytes = nnz * 12 + (n_ + 1) * 8
            tot = nbytes
            if dist is not None:
                t = torch.tensor([nbytes], device="cuda", dtype=torch.int64)
                dist.all_reduce(t)
                tot = int(t.item())
            out["link_ceiling"] = {"ms": t_link * 1e3, "bytes_all_ranks": tot, "aggregate_GBps": tot / t_link / 1e9,
                                   "per_rank_GBps": nbytes / t_link / 1e9}
            out["rank0"] = {"nnz": nnz, "h2d_bytes": W.dmesh.h2d_bytes_last, "d2h_bytes": nnz * 16 + (n_ + 1) * 8}
            del dev_rv, dev_nz, dev_cp, h_rv, h_nz, h_cp
        W.release()
        del W, outarr
    out["transport"] = results
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
