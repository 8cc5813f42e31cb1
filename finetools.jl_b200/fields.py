"""Nodal fields and degree-of-freedom numbering (reference: src/FieldModule.jl:47-53, 102, 263-275, 304-314,
328-377, 400-420; src/NodalFieldModule.jl:20-40; src/FENodeSetModule.jl:26-30).  Host-side only: the numbering is
an INPUT of the assembly path."""
import numpy as np

DOF_KIND_FREE = 1
DOF_KIND_DATA = 2


class FENodeSet:
    def __init__(self, xyz):
        self.xyz = np.array(xyz, dtype=np.float64)  # copy, like the reference

    def count(self):
        return self.xyz.shape[0]


class NodalField:
    """values (nents, ndn) float64, dofnums (nents, ndn) int64 (0 until numbered), kind (nents, ndn) int8.
    values and dofnums are stored column-major (order="F"), i.e. with the bytes of the Julia matrices they mirror, so they
    cross the C ABI without a transposing copy."""

    def __init__(self, data):
        data = np.asarray(data, dtype=np.float64)
        if data.ndim == 1:
            data = data.reshape(-1, 1)
        self.values = np.array(data, dtype=np.float64, order="F")  # copy, like the reference (NodalFieldModule.jl:34-40)
        self.dofnums = np.zeros(self.values.shape, dtype=np.int64, order="F")
        self.kind = np.full(self.values.shape, DOF_KIND_FREE, dtype=np.int8)
        self.ranges = []
        self._dofver = 0  # bumped whenever the numbering changes (device dof maps are cached against it)

    def ndofs(self):
        return self.values.shape[1]

    def nents(self):
        return self.values.shape[0]

    nnodes = nents

    def nalldofs(self):
        return self.values.size

    def nfreedofs(self):
        return int(np.count_nonzero(self.kind == DOF_KIND_FREE))

    def nfixeddofs(self):
        return self.nalldofs() - self.nfreedofs()


def ndofs(f):
    return f.ndofs()


def nents(f):
    return f.nents()


def nalldofs(f):
    return f.nalldofs()


def nfreedofs(f):
    return f.nfreedofs()


def numberdofs(self, entperm=None, kinds=(DOF_KIND_FREE, DOF_KIND_DATA)):
    """Free dofs first (entity order, component inner), then the fixed ones (FieldModule.jl:360-377)."""
    n, dim = self.values.shape
    perm = np.arange(n) if entperm is None else np.asarray(entperm, dtype=np.int64) - 1
    kind = self.kind[perm, :].reshape(-1)  # entity-major, component inner
    flat = np.zeros(n * dim, dtype=np.int64)
    nxt = 1
    self.ranges = []
    for k in kinds:
        sel = np.nonzero(kind == k)[0]
        flat[sel] = np.arange(nxt, nxt + sel.size)
        self.ranges.append((nxt, nxt + sel.size - 1))
        nxt += sel.size
    dn = np.zeros(self.values.shape, dtype=np.int64, order="F")
    dn[perm, :] = flat.reshape(n, dim)
    # a NEW, read-only array per numbering: device dof maps are cached against (field, version, array) and a frozen array cannot
    # be edited behind the cache's back (an in-place write raises; whoever wants other numbers renumbers or replaces the array)
    dn.flags.writeable = False
    self.dofnums = dn
    self._dofver = getattr(self, "_dofver", 0) + 1
    return self


def setebc(self, fenids=None, is_fixed=True, comp=None, val=0.0):
    """setebc!(field, fenids, is_fixed, comp, val) (FieldModule.jl:400-560 family); comp 1-based, None = all."""
    n, dim = self.values.shape
    ids = np.arange(1, n + 1) if fenids is None else np.atleast_1d(np.asarray(fenids, dtype=np.int64))
    comps = range(1, dim + 1) if comp is None else np.atleast_1d(comp)
    vals = np.broadcast_to(np.asarray(val, dtype=np.float64), ids.shape)
    for c in comps:
        if not 1 <= c <= dim:
            raise ValueError("Requested  nonexistent  degree of freedom")
        if ids.size and (ids.min() < 1 or ids.max() > n):
            raise ValueError("Requested nonexistent node")
        if is_fixed:
            self.kind[ids - 1, c - 1] = DOF_KIND_DATA
            self.values[ids - 1, c - 1] = vals
        else:
            self.kind[ids - 1, c - 1] = DOF_KIND_FREE
            self.values[ids - 1, c - 1] = 0.0
    self.ranges = []
    self._dofver = getattr(self, "_dofver", 0) + 1
    return self


def applyebc(self):
    return self


def gathervalues_asmat(self, dest, conn):
    dest[:, :] = self.values[np.asarray(conn) - 1, :]
    return dest


def gatherdofnums(self, dest, conn):
    dest[:] = self.dofnums[np.asarray(conn) - 1, :].reshape(-1)
    return dest


def gathersysvec(self, kind="all"):
    """Values ordered by dof number (FieldModule.jl gathersysvec)."""
    out = np.zeros(self.nalldofs())
    out[self.dofnums.reshape(-1) - 1] = self.values.reshape(-1)
    if kind == "free":
        return out[: self.nfreedofs()]
    return out
