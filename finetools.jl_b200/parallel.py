"""Multi-GPU helpers: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).

Assembly itself needs NO collective: every rank integrates the elements touching its nodes (halo recomputed) and builds
the CSC of its own rows (fegpu_partition_set).  The only exchange is the OPTIONAL gather of the row-block CSCs into one
matrix, implemented here:
  all_gather of the per-column counts -> global colptr = prefix sum of the summed counts; a rank's entries of column j land
  at colptr[j] + (entries of lower ranks in column j).  When the owned dof ranges are ordered by rank (contiguous node
  ranges, default numbering) concatenation keeps rowval sorted; otherwise the merged columns are sorted by row (stable).
"""
import numpy as np


def gather_row_blocks(colptr, rowval, nzval, nrows, ncols, dist=None, device=None, dst=0):
    """Merge the calling rank's 1-based row-block CSC with everybody else's.  Returns (colptr, rowval, nzval) of the full
    matrix on rank `dst`, None elsewhere.  `dist` is torch.distributed (initialised) or None for a single process."""
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
    before = (torch.cumsum(allc, 0) - allc)[rank]      # entries of lower ranks, per column
    # destination (0-based) of each local entry
    local_start = torch.as_tensor(colptr[:-1] - 1, dtype=torch.int64, device=dev)
    col_of = torch.repeat_interleave(torch.arange(ncols, device=dev), counts)
    k = torch.arange(int(counts.sum().item()), device=dev)
    dest = gcolptr[col_of] - 1 + before[col_of] + (k - local_start[col_of])
    rv = torch.as_tensor(rowval, dtype=torch.int64, device=dev)
    nz = torch.as_tensor(nzval, dtype=torch.float64, device=dev)
    nnz_all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(nnz_all, torch.tensor([rv.numel()], dtype=torch.int64, device=dev))
    nnz_all = [int(t.item()) for t in nnz_all]
    if rank == dst:
        nnz = int(sum(nnz_all))
        out_rv = torch.empty(nnz, dtype=torch.int64, device=dev)
        out_nz = torch.empty(nnz, dtype=torch.float64, device=dev)
        out_rv[dest] = rv
        out_nz[dest] = nz
        for src in range(world):
            if src == dst:
                continue
            n = nnz_all[src]
            d = torch.empty(n, dtype=torch.int64, device=dev)
            r = torch.empty(n, dtype=torch.int64, device=dev)
            v = torch.empty(n, dtype=torch.float64, device=dev)
            dist.recv(d, src=src); dist.recv(r, src=src); dist.recv(v, src=src)
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
    dist.send(dest, dst=dst); dist.send(rv, dst=dst); dist.send(nz, dst=dst)
    return None
