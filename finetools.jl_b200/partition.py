"""Node partitioning for the multi-GPU row-block split.

pointpartitioning: recursive inertial bisection, restating src/MeshModificationModule.jl:814-884 (_nodepartitioning3) and
:1029-1039: split along the principal axis of the point cloud's covariance at the median of the projected
coordinate; labels 1..2^ceil(log2(npartitions)).  slab_owner: contiguous node ranges, which for the block generators'
x-fastest node numbering are z-slabs.
"""
import numpy as np


def slab_owner(nnodes, nparts):
    """owner = floor(node * P / nnodes), 0-based ranks, contiguous ascending node ranges."""
    own = ((np.arange(nnodes, dtype=np.int64) * int(nparts)) // int(nnodes)).astype(np.int32)
    own.flags.writeable = False  # device partitions are cached against the array: frozen = trusted by identity
    return own


def pointpartitioning(xyz, npartitions=2):
    """Partition labels 1..2^k (k = ceil(log2(npartitions))) by recursive inertial bisection: every current part is cut
    through its centroid, normal to the eigenvector of the SMALLEST eigenvalue of its inertia matrix (the long direction);
    points with d < 0 get label 2p-1, d > 0 label 2p, exact zeros alternate (MeshModificationModule.jl:856-876).  The
    reference mirrors only the (1,2) entry of the inertia matrix (:846-850); restated as is."""
    xyz = np.asarray(xyz, dtype=np.float64)
    if npartitions < 2:
        raise ValueError("Number of partitions must be >= 2")
    sdim = xyz.shape[1]
    if sdim not in (2, 3):
        raise ValueError("Not implemented for 1D")  # the reference only warns (MeshModificationModule.jl:1037)
    nlevels = int(round(np.ceil(np.log(npartitions) / np.log(2))))
    part = np.ones(xyz.shape[0], dtype=np.int64)
    for level in range(nlevels):
        newpart = part.copy()
        for p in range(1, 2 ** level + 1):
            idx = np.nonzero(part == p)[0]
            if idx.size == 0:
                continue
            X = xyz[idx]
            r = X - X.sum(axis=0) / idx.size
            if sdim == 3:
                x, y, z = r[:, 0], r[:, 1], r[:, 2]
                M = np.zeros((3, 3))
                M[0, 0] = (y * y + z * z).sum(); M[1, 1] = (x * x + z * z).sum(); M[2, 2] = (y * y + x * x).sum()
                M[0, 1] = -(x * y).sum(); M[0, 2] = -(x * z).sum(); M[1, 2] = -(y * z).sum()
                M[1, 0] = M[0, 1]
            else:  # planar point sets (_nodepartitioning2, MeshModificationModule.jl:886-951): the full symmetric 2 x 2 matrix
                x, y = r[:, 0], r[:, 1]
                M = np.array([[(y * y).sum(), -(x * y).sum()], [-(x * y).sum(), (x * x).sum()]])
            vals, vecs = np.linalg.eig(M)
            v = np.real(vecs[:, np.argsort(np.real(vals))[0]])
            d = r @ v
            c = np.where(d < 0.0, 1, 0)
            zeros = np.nonzero(d == 0.0)[0]
            c[zeros[0::2]] = 1      # toggle starts at +1 -> c = 1, then alternates
            c[zeros[1::2]] = 0
            newpart[idx] = 2 * p - c
        part = newpart
    return part
