"""Node partitioning for the multi-GPU row-block split.

pointpartitioning: recursive inertial bisection, restating src/MeshModificationModule.jl:814-884 (_nodepartitioning3) and
:1029-1039: split along the principal axis of the point cloud's covariance at the median of the projected
coordinate; labels 1..2^ceil(log2(npartitions)).  slab_owner: contiguous node ranges, which for the block generators'
x-fastest node numbering are z-slabs.
"""
import numpy as np


def slab_owner(nnodes, nparts):
    """owner = floor(node * P / nnodes), 0-based ranks, contiguous ascending node ranges."""
    return ((np.arange(nnodes, dtype=np.int64) * int(nparts)) // int(nnodes)).astype(np.int32)


def pointpartitioning(xyz, npartitions=2):
    """Partition labels 1..2^k (k = ceil(log2(npartitions))) by recursive inertial bisection."""
    xyz = np.asarray(xyz, dtype=np.float64)
    n = xyz.shape[0]
    nlevels = int(np.ceil(np.log2(max(int(npartitions), 1)))) if npartitions > 1 else 0
    part = np.ones(n, dtype=np.int64)
    for level in range(nlevels):
        newpart = part.copy()
        for p in range(1, 2 ** level + 1):
            idx = np.nonzero(part == p)[0]
            if idx.size == 0:
                continue
            X = xyz[idx]
            Xc = X - X.mean(axis=0)
            # principal direction = eigenvector of the largest eigenvalue of the covariance
            w, v = np.linalg.eigh(Xc.T @ Xc)
            d = Xc @ v[:, -1]
            med = np.median(d)
            right = d > med
            ties = np.nonzero(d == med)[0]       # zero-distance points alternate sides (:869-872)
            right[ties[1::2]] = True
            newpart[idx[right]] = p + 2 ** level
        part = newpart
    return part
