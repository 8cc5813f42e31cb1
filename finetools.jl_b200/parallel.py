"""Multi-GPU helpers: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).

Assembly itself needs NO collective: every rank integrates the elements touching its nodes (halo recomputed) and builds
the CSC of its own rows (fegpu_partition_set).  The only exchange is the OPTIONAL gather of the row-block CSCs into one
matrix (the reference's makematrix! returns one SparseMatrixCSC, AssemblyModule.jl:319-325):

  all-gather of the per-column counts -> global colptr = prefix sum of the summed counts; a rank's entries of column j land
  at colptr[j] + (entries of lower ranks in column j).  When the owned dof ranges are ordered by rank (contiguous node
  ranges, default numbering) concatenation keeps rowval sorted; otherwise the merged columns are sorted by row.

gather_row_blocks_device: the blocks stay in HBM.  NCCL moves 8 B x ncols counts per rank and every block's rowval / nzval
slabs once (16 B per non-zero, no index array on the wire); the library's kernels (csrc/fegpu_blocks.cu) do the counts, the
plan and the interleave.  gather_row_blocks: the same choreography on host arrays (any backend), kept for callers that
already fetched their blocks and as the CPU-testable statement of the host logic.
"""
import ctypes as C

import numpy as np


class _DevArray:
    """Zero-copy view of a device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _as_tensor(torch, ptr, n, dtype, device):
    if n == 0:
        return torch.empty(0, dtype=dtype, device=device)
    return torch.as_tensor(_DevArray(ptr, n, "<i8" if dtype == torch.int64 else "<f8"), device=device)


def owned_ranges_ordered(dofnums, owner, world):
    """True when every rank's owned dof numbers lie above those of the ranks below it (then concatenating the blocks' column
    segments in rank order keeps the rows ascending)."""
    top = -1
    for r in range(world):
        d = dofnums[np.asarray(owner) == r]
        if d.size == 0:
            continue
        if int(d.min()) <= top:
            return False
        top = int(d.max())
    return True


def gather_row_blocks_device(assembler, dist, dst=0, ordered=True):
    """Merge every rank's device-resident row-block CSC (the result `assembler` holds after a partitioned form call) into one
    matrix in the HBM of rank `dst`.  Returns (colptr, rowval, nzval) as CUDA tensors (1-based int64 / float64) on `dst`, None
    elsewhere.  `ordered` = owned_ranges_ordered(...): False adds the per-column sort of the merged result."""
    import torch
    from . import _lib
    L = _lib.lib()
    ctx = assembler.ctx
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", ctx.device)
    m, n, nnz = assembler.sizes()
    cp, rp, vp = assembler.device_pointers()
    ctx.synchronize()  # the block is complete; from here on torch's streams and the context's alternate
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(L.fegpu_block_counts(assembler.handle, C.c_void_p(counts.data_ptr())), ctx.handle)
    ctx.synchronize()
    allc = torch.empty(world * n, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, counts)                     # 8 B x ncols per rank
    torch.cuda.current_stream(dev).synchronize()
    gcolptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    nnz_rank = torch.empty(world, dtype=torch.int64, device=dev)
    _lib.check(L.fegpu_gather_plan(ctx.handle, C.c_void_p(allc.data_ptr()), world, n, C.c_void_p(gcolptr.data_ptr()), C.c_void_p(nnz_rank.data_ptr())),
               ctx.handle)
    ctx.synchronize()
    nnz_all = [int(x) for x in nnz_rank.cpu().tolist()]
    assert nnz_all[rank] == nnz, "block size disagrees with its own column counts"
    my_rv, my_nz = _as_tensor(torch, rp, nnz, torch.int64, dev), _as_tensor(torch, vp, nnz, torch.float64, dev)
    if rank != dst:
        if nnz:
            dist.send(my_rv, dst=dst)
            dist.send(my_nz, dst=dst)
        torch.cuda.current_stream(dev).synchronize()
        return None
    total = sum(nnz_all)
    out_rv = torch.empty(total, dtype=torch.int64, device=dev)
    out_nz = torch.empty(total, dtype=torch.float64, device=dev)

    def place(src, rv, nz):
        _lib.check(L.fegpu_gather_place(ctx.handle, C.c_void_p(allc.data_ptr()), world, src, n, C.c_void_p(gcolptr.data_ptr()),
                                        C.c_void_p(rv.data_ptr()), C.c_void_p(nz.data_ptr()), C.c_void_p(out_rv.data_ptr()),
                                        C.c_void_p(out_nz.data_ptr())), ctx.handle)

    if nnz:
        place(rank, my_rv, my_nz)
    cap = max([nnz_all[s] for s in range(world) if s != dst] + [1])
    stage_rv = [torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(2)]   # double buffered: receive the next block
    stage_nz = [torch.empty(cap, dtype=torch.float64, device=dev) for _ in range(2)]  # while the previous one is interleaved
    k = 0
    for src in range(world):
        if src == dst or nnz_all[src] == 0:
            continue
        rv, nz = stage_rv[k % 2][:nnz_all[src]], stage_nz[k % 2][:nnz_all[src]]
        if k >= 2:
            ctx.synchronize()  # the kernel that read this staging buffer two blocks ago is done
        dist.recv(rv, src=src)
        dist.recv(nz, src=src)
        torch.cuda.current_stream(dev).synchronize()
        place(src, rv, nz)
        k += 1
    if not ordered:
        fixed = C.c_int64(0)
        _lib.check(L.fegpu_gather_sort_columns(ctx.handle, n, C.c_void_p(gcolptr.data_ptr()), C.c_void_p(out_rv.data_ptr()),
                                               C.c_void_p(out_nz.data_ptr()), C.byref(fixed)), ctx.handle)
    ctx.synchronize()
    return gcolptr, out_rv, out_nz


def gather_row_blocks(colptr, rowval, nzval, nrows, ncols, dist=None, device=None, dst=0):
    """Merge the calling rank's 1-based row-block CSC (HOST arrays) with everybody else's.  Returns (colptr, rowval, nzval) of
    the full matrix on rank `dst`, None elsewhere.  `dist` is torch.distributed (initialised) or None for a single process."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return colptr, rowval, nzval
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    counts = torch.as_tensor(np.diff(colptr), dtype=torch.int64, device=dev)
    allc = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)                      # 8 B x ncols per rank
    allc = torch.stack(allc)                           # [world][ncols]
    total = allc.sum(dim=0)
    gcolptr = torch.ones(ncols + 1, dtype=torch.int64, device=dev)
    gcolptr[1:] += torch.cumsum(total, 0)
    before = torch.cumsum(allc, 0) - allc              # [world][ncols]: entries of lower ranks, per column
    nnz_all = [int(x) for x in allc.sum(dim=1).tolist()]
    rv = torch.as_tensor(rowval, dtype=torch.int64, device=dev)
    nz = torch.as_tensor(nzval, dtype=torch.float64, device=dev)

    def positions(src):
        """0-based destination of every entry of the block of rank src, from the counts alone (nothing extra on the wire)."""
        c = allc[src]
        start = torch.cumsum(c, 0) - c
        col_of = torch.repeat_interleave(torch.arange(ncols, device=dev), c)
        k = torch.arange(int(c.sum().item()), device=dev)
        return gcolptr[col_of] - 1 + before[src][col_of] + (k - start[col_of])

    if rank == dst:
        nnz = int(sum(nnz_all))
        out_rv = torch.empty(nnz, dtype=torch.int64, device=dev)
        out_nz = torch.empty(nnz, dtype=torch.float64, device=dev)
        d = positions(rank)
        out_rv[d] = rv
        out_nz[d] = nz
        for src in range(world):
            if src == dst or nnz_all[src] == 0:
                continue
            r = torch.empty(nnz_all[src], dtype=torch.int64, device=dev)
            v = torch.empty(nnz_all[src], dtype=torch.float64, device=dev)
            dist.recv(r, src=src); dist.recv(v, src=src)
            d = positions(src)
            out_rv[d] = r
            out_nz[d] = v
        gc = gcolptr.cpu().numpy()
        orv, onz = out_rv.cpu().numpy(), out_nz.cpu().numpy()
        # owned dof ranges interleaved between ranks? then sort every column by row (stable)
        seg_unsorted = np.nonzero(np.diff(orv) <= 0)[0] + 1
        col_starts = set((gc[:-1] - 1).tolist())
        if any(int(p) not in col_starts for p in seg_unsorted):
            col_id = np.repeat(np.arange(ncols), np.diff(gc))
            order = np.lexsort((orv, col_id))
            orv, onz = orv[order], onz[order]
        return gc, orv, onz
    if rv.numel():
        dist.send(rv, dst=dst); dist.send(nz, dst=dst)
    return None
