"""ctypes face of the CPU oracle (oracle/fe_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (finetools.jl_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_tag():
    """The oracle is compiled -march=native (it doubles as the timed CPU baseline), so the binary is keyed by the host's
    CPU feature flags: a .so built in the build container is never loaded on a different CPU."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


_SO = os.path.join(_HERE, "libfe_oracle.%s.so" % _cpu_tag())

ET = {"T3": 1, "Q4": 2, "T4": 3, "T10": 4, "H8": 5, "H20": 6, "H27": 7}

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "fe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "OUT=" + os.path.basename(_SO)], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_nne.argtypes = [C.c_int]
        L.orc_mdim.argtypes = [C.c_int]
        L.orc_bfun.argtypes = [C.c_int, _f64p, _f64p]
        L.orc_bfundpar.argtypes = [C.c_int, _f64p, _f64p]
        for f in (L.orc_gauss_rule,):
            f.argtypes = [C.c_int, C.c_int, _f64p, _f64p]
        L.orc_tet_rule.argtypes = [C.c_int, _f64p, _f64p]
        L.orc_tri_rule.argtypes = [C.c_int, _f64p, _f64p]
        L.orc_add_mggt_ut_only.argtypes = [_f64p, _f64p, C.c_double, C.c_int, C.c_int]
        L.orc_add_mggt_ut_only.restype = None
        L.orc_add_gkgt_ut_only.argtypes = [_f64p, _f64p, C.c_double, _f64p, _f64p, C.c_int, C.c_int]
        L.orc_add_gkgt_ut_only.restype = None
        L.orc_add_btdb_ut_only.argtypes = [_f64p, _f64p, C.c_double, _f64p, _f64p, C.c_int, C.c_int]
        L.orc_add_btdb_ut_only.restype = None
        L.orc_complete_lt.argtypes = [_f64p, C.c_int]
        L.orc_complete_lt.restype = None
        L.orc_assemble.argtypes = [_i64p, _i64p, _f64p, _i64p, _f64p, _i64p, C.c_int, _i64p, C.c_int, C.c_int64, C.c_int64]
        common = [C.c_int, C.c_int64, _i64p, C.c_int64, C.c_int, _f64p]
        L.orc_bilform_diffusion.argtypes = common + [_i64p, C.c_int64, C.c_int, _f64p, _f64p, C.c_int, _f64p, C.c_double, C.c_void_p, _i64p, _i64p, _f64p]
        L.orc_bilform_lin_elastic.argtypes = common + [_i64p, C.c_int64, C.c_int, _f64p, _f64p, _f64p, C.c_void_p, _i64p, _i64p, _f64p]
        L.orc_bilform_dot.argtypes = common + [C.c_int, _i64p, C.c_int64, C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double,
                                               _i64p, _i64p, _f64p]
        L.orc_bilform_convection.argtypes = common + [_f64p, C.c_int, _i64p, C.c_int64, C.c_int, _f64p, _f64p, C.c_double, C.c_double, _i64p, _i64p, _f64p]
        L.orc_bilform_div_grad.argtypes = common + [C.c_int, _i64p, C.c_int64, C.c_int, _f64p, _f64p, C.c_double, C.c_double, _i64p, _i64p, _f64p]
        L.orc_bilform_masslike.argtypes = common + [C.c_int, _i64p, C.c_int64, C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double,
                                                    _i64p, _i64p, _f64p]
        L.orc_linform_dot.argtypes = common + [C.c_int, _i64p, C.c_int64, C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double, _f64p]
        L.orc_sparse.argtypes = [C.c_int64, _i64p, _i64p, _f64p, C.c_int64, C.c_int64, _i64p, C.c_void_p, C.c_void_p]
        L.orc_sparse.restype = C.c_int64
        _lib = L
    return _lib


def _F(a):  # column-major (Julia) matrix -> flat C-contiguous buffer holding the column-major bytes
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1)


def _I(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).T).reshape(-1)


def bfun(et, pc):
    n = lib().orc_nne(ET[et])
    out = np.zeros(n)
    lib().orc_bfun(ET[et], np.ascontiguousarray(pc, dtype=np.float64), out)
    return out


def bfundpar(et, pc):
    n, m = lib().orc_nne(ET[et]), lib().orc_mdim(ET[et])
    out = np.zeros(n * m)
    lib().orc_bfundpar(ET[et], np.ascontiguousarray(pc, dtype=np.float64), out)
    return out.reshape(m, n).T.copy()


def gauss_rule(dim, order):
    npts = order ** dim
    pc, w = np.zeros(npts * dim), np.zeros(npts)
    r = lib().orc_gauss_rule(dim, order, pc, w)
    assert r == npts
    return pc.reshape(dim, npts).T.copy(), w


def tet_rule(npts):
    pc, w = np.zeros(npts * 3), np.zeros(npts)
    assert lib().orc_tet_rule(npts, pc, w) == npts
    return pc.reshape(3, npts).T.copy(), w


def tri_rule(npts):
    pc, w = np.zeros(npts * 2), np.zeros(npts)
    assert lib().orc_tri_rule(npts, pc, w) == npts
    return pc.reshape(2, npts).T.copy(), w


def _prep(et, conn, xyz, dofnums, pc, w):
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    nelem, nne = conn.shape
    assert nne == lib().orc_nne(ET[et])
    xyz = np.asarray(xyz, dtype=np.float64)
    nnodes, sdim = xyz.shape
    dofnums = np.asarray(dofnums, dtype=np.int64).reshape(nnodes, -1)
    ndn = dofnums.shape[1]
    pc = np.asarray(pc, dtype=np.float64)
    npts = pc.shape[0]
    return conn, nelem, nne, _F(xyz), nnodes, sdim, _I(dofnums), ndn, _F(pc), np.ascontiguousarray(w, dtype=np.float64).reshape(-1), npts


def _rm(Rm):
    """constant material coordinate system matrix (CSys(csmat), CSysModule.jl:133-144) as a column-major buffer, or NULL = identity"""
    if Rm is None:
        return None, None
    buf = _F(np.asarray(Rm, dtype=np.float64))
    return buf, buf.ctypes.data_as(C.c_void_p)


def bilform_diffusion_coo(et, conn, xyz, dofnums, nalldofs, pc, w, kappa, otherdim=1.0, Rm=None, out=None):
    """Reference-order COO triplets (I, J, V) of bilform_diffusion; kappa scalar -> iso path, matrix -> general.
    out = preallocated (I, J, V) contiguous slices for these elements (several threads may fill one COO buffer)."""
    conn, nelem, nne, X, nnodes, sdim, dn, ndn, P, W, npts = _prep(et, conn, xyz, dofnums, pc, w)
    assert ndn == 1
    n = nelem * nne * nne
    I, J, V = out if out is not None else (np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n))
    assert I.size == n
    kind = 0 if np.ndim(kappa) == 0 else 1
    kap = _F(np.atleast_2d(np.asarray(kappa, dtype=np.float64)))
    rmbuf, rmp = _rm(Rm)
    rc = lib().orc_bilform_diffusion(ET[et], nelem, conn.reshape(-1), nnodes, sdim, X, dn, nalldofs, npts, P, W, kind, kap, float(otherdim), rmp, I, J, V)
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return I, J, V


def bilform_lin_elastic_coo(et, conn, xyz, dofnums, nalldofs, pc, w, Cmat, out=None, Rm=None):
    """out = preallocated (I, J, V) contiguous slices for these elements (lets several threads fill one COO buffer;
    ctypes drops the GIL during the call)."""
    conn, nelem, nne, X, nnodes, sdim, dn, ndn, P, W, npts = _prep(et, conn, xyz, dofnums, pc, w)
    assert ndn == 3
    n = nelem * (3 * nne) ** 2
    I, J, V = out if out is not None else (np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n))
    assert I.size == n
    rmbuf, rmp = _rm(Rm)
    rc = lib().orc_bilform_lin_elastic(ET[et], nelem, conn.reshape(-1), nnodes, sdim, X, dn, nalldofs, npts, P, W, _F(Cmat), rmp, I, J, V)
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return I, J, V


def bilform_dot_coo(et, conn, xyz, dofnums, nalldofs, pc, w, c, m=3, otherdim=1.0):
    conn, nelem, nne, X, nnodes, sdim, dn, ndn, P, W, npts = _prep(et, conn, xyz, dofnums, pc, w)
    n = nelem * (ndn * nne) ** 2
    I, J, V = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n)
    cm = _F(np.asarray(c, dtype=np.float64).reshape(ndn, ndn))
    rc = lib().orc_bilform_dot(ET[et], nelem, conn.reshape(-1), nnodes, sdim, X, ndn, dn, nalldofs, npts, P, W, cm, m, otherdim, I, J, V)
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return I, J, V


def bilform_convection_coo(et, conn, xyz, uvals, dofnums, nalldofs, pc, w, rho=1.0, otherdim=1.0):
    """Reference-order COO triplets of bilform_convection (FEMMBaseModule.jl:1583-1625); uvals = nodal velocities nnodes x sdim."""
    conn, nelem, nne, X, nnodes, sdim, dn, ndn, P, W, npts = _prep(et, conn, xyz, dofnums, pc, w)
    assert ndn == 1
    uv = np.asarray(uvals, dtype=np.float64).reshape(nnodes, -1)
    n = nelem * nne * nne
    I, J, V = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n)
    rc = lib().orc_bilform_convection(ET[et], nelem, conn.reshape(-1), nnodes, sdim, X, _F(uv), uv.shape[1], dn, nalldofs, npts, P, W,
                                      float(rho), float(otherdim), I, J, V)
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return I, J, V


def bilform_div_grad_coo(et, conn, xyz, dofnums, nalldofs, pc, w, mu, otherdim=1.0):
    """Reference-order COO triplets of bilform_div_grad (FEMMBaseModule.jl:1672-1713)."""
    conn, nelem, nne, X, nnodes, sdim, dn, ndn, P, W, npts = _prep(et, conn, xyz, dofnums, pc, w)
    n = nelem * (ndn * nne) ** 2
    I, J, V = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n)
    rc = lib().orc_bilform_div_grad(ET[et], nelem, conn.reshape(-1), nnodes, sdim, X, ndn, dn, nalldofs, npts, P, W, float(mu), float(otherdim), I, J, V)
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return I, J, V


def bilform_masslike_coo(et, conn, xyz, dofnums, nalldofs, pc, w, c, m=3, otherdim=1.0):
    """Reference-order COO of bilform_masslike (FEMMBaseModule.jl:1865-1912): (nelem*ndn) x nalldofs, rows numbered by element."""
    conn, nelem, nne, X, nnodes, sdim, dn, ndn, P, W, npts = _prep(et, conn, xyz, dofnums, pc, w)
    n = nelem * ndn * nne * ndn
    I, J, V = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n)
    cm = _F(np.asarray(c, dtype=np.float64).reshape(ndn, ndn))
    rc = lib().orc_bilform_masslike(ET[et], nelem, conn.reshape(-1), nnodes, sdim, X, ndn, dn, nalldofs, npts, P, W, cm, m, otherdim, I, J, V)
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return I, J, V


def linform_dot(et, conn, xyz, dofnums, nalldofs, pc, w, force, m=3, otherdim=1.0):
    """The assembled vector of linform_dot / distribloads with a constant force (FEMMBaseModule.jl:1207-1244)."""
    conn, nelem, nne, X, nnodes, sdim, dn, ndn, P, W, npts = _prep(et, conn, xyz, dofnums, pc, w)
    fv = np.ascontiguousarray(np.asarray(force, dtype=np.float64).reshape(-1))
    assert fv.size == ndn
    F = np.zeros(nalldofs)
    rc = lib().orc_linform_dot(ET[et], nelem, conn.reshape(-1), nnodes, sdim, X, ndn, dn, nalldofs, npts, P, W, fv, m, otherdim, F)
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return F


_ERR = {1: "Column degree of freedom < 1", 2: "Column degree of freedom > size", 3: "Row degree of freedom < 1",
        4: "Row degree of freedom > size", -2: "manifold/space dimension mismatch", -3: "That is the only acceptable option here."}


def sparse(I, J, V, m, n):
    """Julia's sparse(I,J,V,m,n): returns 1-based (colptr, rowval, nzval)."""
    I = np.ascontiguousarray(I, dtype=np.int64)
    J = np.ascontiguousarray(J, dtype=np.int64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    nt = I.size
    colptr = np.zeros(n + 1, np.int64)
    rowval = np.zeros(max(nt, 1), np.int64)
    nzval = np.zeros(max(nt, 1))
    nnz = lib().orc_sparse(nt, I, J, V, m, n, colptr, rowval.ctypes.data, nzval.ctypes.data)
    if nnz < 0:
        raise ValueError("row/column index out of range")
    return colptr, rowval[:nnz].copy(), nzval[:nnz].copy()


def lumped_coo(I, J, V, em, mode):
    """COO of SysmatAssemblerSparseDiag (mode 1, AssemblyModule.jl:718-743: V = mat[j, j]) or
    SysmatAssemblerSparseHRZLumpingSymm (mode 2, :1070-1101: V = mat[j, j] * sum(mat) / trace(mat)) from the full reference-order
    triplets of `nelem` square em x em element matrices; feed the result to sparse(I, I, V, n, n) (:770-778, :1123-1131)."""
    nelem = V.size // (em * em)
    mats = V.reshape(nelem, em, em).transpose(0, 2, 1)          # [e][row][col] from the column-major emission order
    cols = J.reshape(nelem, em, em)[:, :, 0]                     # column dof of every column
    diag = np.einsum("eii->ei", mats).copy()
    if mode == 2:
        em2 = mats.sum(axis=1).sum(axis=1)                       # sum(sum(mat, dims = 1))
        dem2 = np.zeros(nelem)
        for i in range(em):                                      # dem2 += mat[i, i], in order
            dem2 = dem2 + mats[:, i, i]
        diag = diag * (em2 / dem2)[:, None]
    return cols.reshape(-1).copy(), diag.reshape(-1)


def matrix_block(csc, r0, r1, c0, c1):
    """Julia's A[r0:r1, c0:c1] on a SparseMatrixCSC given as 1-based (colptr, rowval, nzval): stored zeros are kept, rows are
    rebased to 1 (what matrix_blocked_ff/fd/df/dd do, src/MatrixUtilityModule.jl:675-793).  Pure NumPy."""
    colptr, rowval, nzval = csc
    ncols = max(c1 - c0 + 1, 0)
    cp = np.ones(ncols + 1, np.int64)
    rv, nz = [], []
    for k, c in enumerate(range(c0, c1 + 1)):
        b, e = colptr[c - 1] - 1, colptr[c] - 1
        r = rowval[b:e]
        sel = (r >= r0) & (r <= r1)
        rv.append(r[sel] - r0 + 1)
        nz.append(nzval[b:e][sel])
        cp[k + 1] = cp[k] + int(sel.sum())
    rv = np.concatenate(rv) if rv else np.zeros(0, np.int64)
    nz = np.concatenate(nz) if nz else np.zeros(0)
    return cp, rv.astype(np.int64), nz


def lower_triangle_mask(em, nelem):
    """Positions, inside the full emission-order triplet stream (AssemblyModule.jl:261-280), of the entries
    SysmatAssemblerSparseSymm's assemble! collects (local i >= j, src/AssemblyModule.jl:517-530)."""
    k = np.arange(em * em)
    return np.tile((k % em) >= (k // em), nelem)


def sparse_symm(I, J, V, n):
    """makematrix!(::SysmatAssemblerSparseSymm) (src/AssemblyModule.jl:551-583) on the lower-triangle triplets:
    S = sparse(I, J, V, n, n); S = S + transpose(S); S[j, j] *= 0.5.  SparseArrays' sparse `+` is a zero-preserving map
    (higherorderfns.jl `_map_zeropres!`), which stores a result only when it is non-zero: entries that cancel to exactly
    0.0 (and stored zeros) are absent.  Restated from the published stdlib algorithm; Julia is not available here."""
    cp, rv, nz = sparse(I, J, V, n, n)
    cols = np.repeat(np.arange(1, n + 1), np.diff(cp))
    # S[i,j] + S'[i,j]: at most one stored entry from each operand, so the two-term sum is order-independent
    cp2, rv2, nz2 = sparse(np.concatenate([rv, cols]), np.concatenate([cols, rv]), np.concatenate([nz, nz]), n, n)
    cols2 = np.repeat(np.arange(1, n + 1), np.diff(cp2))
    keep = nz2 != 0.0
    nz2 = np.where(rv2 == cols2, nz2 * 0.5, nz2)
    rv3, nz3, cols3 = rv2[keep], nz2[keep], cols2[keep]
    cp3 = np.ones(n + 1, np.int64)
    np.add.at(cp3, cols3, 1)
    return np.cumsum(cp3) - np.arange(n + 1), rv3, nz3


def to_scipy(colptr, rowval, nzval, m, n):
    import scipy.sparse as sp
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m, n))
