"""Block mesh generators and boundary extraction (benchmark-input side of the hot path).

Vectorised NumPy restatements of the numbering conventions of
  H8blockx   src/MeshHexahedronModule.jl:59-112   (nodes x-fastest, ELEMENTS z-fastest)
  H8toH20    src/MeshHexahedronModule.jl:580-628  H8toH27 :207-307 (edges, then faces, then bodies)
  T4blockx   src/MeshTetrahedronModule.jl:81-151  T4toT10 :158-212
  Q4blockx   src/MeshQuadrilateralModule.jl:171-215, T3blockx src/MeshTriangleModule.jl:26-75
  meshboundary src/MeshModificationModule.jl:69-178 (faces kept when they occur once; output in lexicographic
               order of the sorted node ids, orientation as in the owning element)
  linearspace src/MeshUtilModule.jl:116-118 (Julia `range(start, stop=, length=)`)
New mid-side nodes are numbered in first-encounter order over (element, local edge), as the reference's
hyperface container does (src/MeshUtilModule.jl:41-80).
"""
from fractions import Fraction

import numpy as np

from .fesets import FESetH8, FESetH20, FESetH27, FESetQ4, FESetT3, FESetT4, FESetT10
from .fields import FENodeSet


def linearspace(start, stop, N):
    """Julia's range(start, stop=stop, length=N): values are (nearly always) the correctly rounded members of the
    exact arithmetic progression.  We evaluate the progression in rationals when both ends are short decimals,
    else fall back to numpy.linspace."""
    start, stop = float(start), float(stop)
    if N == 1:
        return np.array([start])
    try:
        a, b = Fraction(repr(start)), Fraction(repr(stop))
        if float(a) == start and float(b) == stop and max(a.denominator, b.denominator) <= 10 ** 9:
            step = (b - a) / (N - 1)
            return np.array([float(a + i * step) for i in range(N)])
    except (ValueError, OverflowError):
        pass
    return np.linspace(start, stop, N)


def _grid_nodes(xs, ys, zs=None):
    if zs is None:
        X, Y = np.meshgrid(xs, ys, indexing="xy")  # x fastest
        return np.column_stack([X.reshape(-1), Y.reshape(-1)])
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # x fastest, then y, then z
    return np.column_stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)])


def _cell_nodes(nL, nW, nH):
    """8 corner node numbers (1-based, H8 order) of every cell, cells ordered i outer, j, k inner."""
    i, j, k = np.meshgrid(np.arange(1, nL + 1), np.arange(1, nW + 1), np.arange(1, nH + 1), indexing="ij")
    i, j, k = (a.reshape(-1).astype(np.int64) for a in (i, j, k))
    f = (k - 1) * ((nL + 1) * (nW + 1)) + (j - 1) * (nL + 1) + i
    lo = np.column_stack([f, f + 1, f + (nL + 1) + 1, f + (nL + 1)])
    return np.hstack([lo, lo + (nL + 1) * (nW + 1)]), (i, j, k)


def H8blockx(xs, ys, zs):
    xs, ys, zs = (np.asarray(v, dtype=np.float64).reshape(-1) for v in (xs, ys, zs))
    nL, nW, nH = len(xs) - 1, len(ys) - 1, len(zs) - 1
    conn, _ = _cell_nodes(nL, nW, nH)
    return FENodeSet(_grid_nodes(xs, ys, zs)), FESetH8(conn)


def H8block(Length, Width, Height, nL, nW, nH):
    return H8blockx(linearspace(0.0, Length, nL + 1), linearspace(0.0, Width, nW + 1), linearspace(0.0, Height, nH + 1))


def _first_encounter_ids(keys):
    """keys: (n, k) int64 rows (already sorted within a row).  Returns for each row the rank of its key in
    first-encounter order, and the representative rows (unique keys in that order)."""
    if keys.shape[1] == 2:   # pack a node pair into one int64 (node numbers < 2^31)
        flat = keys[:, 0] * (int(keys.max()) + 1) + keys[:, 1]
        _, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    else:
        _, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique keys sorted by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv.reshape(-1)], keys[first[order]]


def _hyperface_nodes(xyz, conn, table, newn0):
    """Number one new node per distinct hyperface listed by `table` (local 0-based node tuples) over all elements,
    first-encounter order; coordinates = mean of the hyperface's nodes summed as (others ascending, anchor last)."""
    nel = conn.shape[0]
    hv = conn[:, np.asarray(table)].reshape(nel * len(table), -1)   # element-major, then local hyperface
    keys = np.sort(hv, axis=1)
    ids, reps = _first_encounter_ids(keys)
    k = reps.shape[1]
    s = xyz[reps[:, 1] - 1].copy()
    for c in range(2, k):
        s = s + xyz[reps[:, c] - 1]
    s = s + xyz[reps[:, 0] - 1]
    newxyz = s / float(k)
    return (ids + newn0).reshape(nel, len(table)), newxyz


_H8_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
_H8_FACES = [(0, 3, 2, 1), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (5, 6, 7, 4)]


def H8toH20(fens, fes):
    xyz, conn = fens.xyz, fes.conn
    econn, exyz = _hyperface_nodes(xyz, conn, _H8_EDGES, xyz.shape[0] + 1)
    out = FESetH20(np.hstack([conn, econn]))
    out.label = fes.label.copy()
    return FENodeSet(np.vstack([xyz, exyz])), out


def H8toH27(fens, fes):
    xyz, conn = fens.xyz, fes.conn
    econn, exyz = _hyperface_nodes(xyz, conn, _H8_EDGES, xyz.shape[0] + 1)
    fconn, fxyz = _hyperface_nodes(xyz, conn, _H8_FACES, xyz.shape[0] + exyz.shape[0] + 1)
    vconn, vxyz = _hyperface_nodes(xyz, conn, [tuple(range(8))], xyz.shape[0] + exyz.shape[0] + fxyz.shape[0] + 1)
    out = FESetH27(np.hstack([conn, econn, fconn, vconn]))
    out.label = fes.label.copy()
    return FENodeSet(np.vstack([xyz, exyz, fxyz, vxyz])), out


def H20block(Length, Width, Height, nL, nW, nH):
    return H8toH20(*H8block(Length, Width, Height, nL, nW, nH))


def H27block(Length, Width, Height, nL, nW, nH):
    return H8toH27(*H8block(Length, Width, Height, nL, nW, nH))


_T4_TABLES = {
    "a": ([[1, 8, 5, 6], [3, 4, 2, 7], [7, 2, 6, 8], [4, 7, 8, 2], [2, 1, 6, 8], [4, 8, 1, 2]],) * 2,
    "b": ([[2, 7, 5, 6], [1, 8, 5, 7], [1, 3, 4, 8], [2, 1, 5, 7], [1, 2, 3, 7], [3, 7, 8, 1]],) * 2,
    "ca": ([[8, 4, 7, 5], [6, 7, 2, 5], [3, 4, 2, 7], [1, 2, 4, 5], [7, 4, 2, 5]],
           [[7, 3, 6, 8], [5, 8, 6, 1], [2, 3, 1, 6], [4, 1, 3, 8], [6, 3, 1, 8]]),
    "cb": ([[7, 3, 6, 8], [5, 8, 6, 1], [2, 3, 1, 6], [4, 1, 3, 8], [6, 3, 1, 8]],
           [[8, 4, 7, 5], [6, 7, 2, 5], [3, 4, 2, 7], [1, 2, 4, 5], [7, 4, 2, 5]]),
}


def T4blockx(xs, ys, zs, orientation="a"):
    orientation = str(orientation).lstrip(":")
    if orientation not in _T4_TABLES:
        raise ValueError("Unknown orientation")
    xs, ys, zs = (np.asarray(v, dtype=np.float64).reshape(-1) for v in (xs, ys, zs))
    nL, nW, nH = len(xs) - 1, len(ys) - 1, len(zs) - 1
    nn, (i, j, k) = _cell_nodes(nL, nW, nH)
    ta, tb = (np.asarray(t, dtype=np.int64) - 1 for t in _T4_TABLES[orientation])
    even = ((i + j + k) % 2 == 0)
    tets = np.where(even[:, None, None], nn[:, tb], nn[:, ta])     # (ncell, ntet, 4)
    return FENodeSet(_grid_nodes(xs, ys, zs)), FESetT4(tets.reshape(-1, 4))


def T4block(Length, Width, Height, nL, nW, nH, orientation="a"):
    return T4blockx(linearspace(0.0, Length, nL + 1), linearspace(0.0, Width, nW + 1), linearspace(0.0, Height, nH + 1),
                    orientation)


_T4_EDGES = [(0, 1), (1, 2), (2, 0), (3, 0), (3, 1), (3, 2)]


def T4toT10(fens, fes):
    xyz, conn = fens.xyz, fes.conn
    econn, exyz = _hyperface_nodes(xyz, conn, _T4_EDGES, xyz.shape[0] + 1)
    out = FESetT10(np.hstack([conn, econn]))
    out.label = fes.label.copy()
    return FENodeSet(np.vstack([xyz, exyz])), out


def T10block(Length, Width, Height, nL, nW, nH, orientation="a"):
    return T4toT10(*T4block(Length, Width, Height, nL, nW, nH, orientation))


def Q4blockx(xs, ys):
    xs, ys = (np.asarray(v, dtype=np.float64).reshape(-1) for v in (xs, ys))
    nL, nW = len(xs) - 1, len(ys) - 1
    i, j = np.meshgrid(np.arange(1, nL + 1), np.arange(1, nW + 1), indexing="ij")
    i, j = i.reshape(-1).astype(np.int64), j.reshape(-1).astype(np.int64)
    f = (j - 1) * (nL + 1) + i
    return FENodeSet(_grid_nodes(xs, ys)), FESetQ4(np.column_stack([f, f + 1, f + (nL + 1) + 1, f + (nL + 1)]))


def Q4block(Length, Width, nL, nW):
    return Q4blockx(linearspace(0.0, Length, nL + 1), linearspace(0.0, Width, nW + 1))


def T3blockx(xs, ys, orientation="a"):
    xs, ys = (np.asarray(v, dtype=np.float64).reshape(-1) for v in (xs, ys))
    nL, nW = len(xs) - 1, len(ys) - 1
    i, j = np.meshgrid(np.arange(1, nL + 1), np.arange(1, nW + 1), indexing="ij")
    i, j = i.reshape(-1).astype(np.int64), j.reshape(-1).astype(np.int64)
    f = (j - 1) * (nL + 1) + i
    if str(orientation).lstrip(":") == "a":
        t1 = np.column_stack([f, f + 1, f + (nL + 1)])
        t2 = np.column_stack([f + 1, f + (nL + 1) + 1, f + (nL + 1)])
    else:
        t1 = np.column_stack([f, f + 1, f + (nL + 1) + 1])
        t2 = np.column_stack([f, f + (nL + 1) + 1, f + (nL + 1)])
    conn = np.stack([t1, t2], axis=1).reshape(-1, 3)
    return FENodeSet(_grid_nodes(xs, ys)), FESetT3(conn)


def T3block(Length, Width, nL, nW, orientation="a"):
    return T3blockx(linearspace(0.0, Length, nL + 1), linearspace(0.0, Width, nW + 1), orientation)


def meshboundary(fes):
    """Boundary facets of a volume mesh (H8 -> Q4, T4 -> T3)."""
    hypf = fes.boundaryconn()
    skey = np.sort(hypf, axis=1)
    # lexicographic order of the sorted node ids, ties broken by the row index (stable LSD column sort)
    order = np.lexsort(tuple(skey[:, c] for c in range(skey.shape[1] - 1, -1, -1)))
    s = skey[order]
    diff = np.any(s[1:] != s[:-1], axis=1)
    first = np.concatenate([[True], diff])   # differs from predecessor
    last = np.concatenate([diff, [True]])    # differs from successor
    keep = order[first & last]
    return fes.boundaryfe(hypf[keep])
