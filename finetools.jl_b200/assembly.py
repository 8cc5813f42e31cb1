"""SysmatAssemblerSparseGPU: drop-in for FinEtools' SysmatAssemblerSparse (src/AssemblyModule.jl:88-329) whose
element loop, triplet storage and COO->CSC conversion run on a B200 through libfinegpu.so.

Protocol mirrored: startassembly! (:209-238), assemble! (:250-282), makematrix! (:304-329), setnomatrixresult (:29),
expectedntriples (:54-59), eltype (:27).  Julia's `!` is dropped from the names.  Error strings are the reference's.
"""
import collections
import contextlib
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import VP, check, fptr


class GPUContext:
    """One CUDA device + stream; owner of every device handle created through it."""
    _default = {}
    MAX_MESHES = 4  # device twins of FESets kept per context (least recently used ones are destroyed beyond this)

    def __init__(self, device=0, stream=None):
        self.handle = VP()
        check(_lib.lib().fegpu_create(C.byref(self.handle), int(device)))
        self.device = int(device)
        # device meshes / dof maps / cached patterns, keyed by the FESet's connectivity array: they belong to the context,
        # not to an assembler, so a fresh assembler per call (the reference's idiom) re-uses them
        self._meshes = collections.OrderedDict()
        if stream is not None:
            self.set_stream(stream)

    def device_mesh(self, fes):
        """The device twin of `fes` on this context (None before its first assembly)."""
        dm = self._meshes.get(id(fes.conn))
        return dm if dm is not None and dm.conn_ref is fes.conn else None

    def release_meshes(self):
        """Destroy every device mesh / dof map / pattern of this context (assemblers keep their own results)."""
        for dm in self._meshes.values():
            dm.destroy()
        self._meshes.clear()

    def marks_begin(self):
        check(_lib.lib().fegpu_marks_begin(self.handle), self.handle)

    def marks_read(self):
        """[(kernel or phase name, ms since the previous mark)] of the step run since marks_begin()."""
        buf = C.create_string_buffer(8192)
        check(_lib.lib().fegpu_marks_read(self.handle, buf, 8192), self.handle)
        out = []
        for item in buf.value.decode().split(";"):
            if "=" in item:
                k, v = item.rsplit("=", 1)
                out.append((k, float(v)))
        return out

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def set_stream(self, cuda_stream_ptr):
        check(_lib.lib().fegpu_set_stream(self.handle, VP(int(cuda_stream_ptr))), self.handle)

    def set_async(self, on):
        """Form calls return once their work is queued (the result calls synchronise); off = every call blocks (default)."""
        self._async = bool(on)
        check(_lib.lib().fegpu_set_async(self.handle, 1 if on else 0), self.handle)

    @contextlib.contextmanager
    def queued_forms(self, enable=True):
        """Scope in which a form call only queues its device work, for callers that fetch the result right after: the
        transport then ships the pattern's arrays (and the host threads rebuild rowval) while the integration and the
        numeric phase are still running.  Restores the caller's own set_async choice on exit."""
        prev = getattr(self, "_async", False)
        enable = enable and os.environ.get("FEGPU_QUEUED_FORMS", "1") != "0"  # A/B knob
        if enable and not prev:
            self.set_async(True)
        try:
            yield self
        finally:
            if enable and not prev:
                self.set_async(False)

    def set_overlap(self, on):
        """Fresh assemblies overlap element integration (second stream) with the symbolic phase; off = strictly serial phases."""
        check(_lib.lib().fegpu_set_overlap(self.handle, 1 if on else 0), self.handle)

    def synchronize(self):
        check(_lib.lib().fegpu_synchronize(self.handle), self.handle)

    def release_cache(self):
        """Hand the device blocks cached by the symbolic phase back to the driver (results and handles stay valid)."""
        check(_lib.lib().fegpu_cache_release(self.handle), self.handle)

    def launch_count(self):
        return int(_lib.lib().fegpu_launch_count(self.handle))

    def transfer_stats(self):
        a, b = C.c_int64(0), C.c_int64(0)
        check(_lib.lib().fegpu_transfer_stats(self.handle, C.byref(a), C.byref(b)), self.handle)
        c = C.c_int64(0)
        check(_lib.lib().fegpu_transfer_compressed(self.handle, C.byref(c)), self.handle)
        d = C.c_int64(0)
        check(_lib.lib().fegpu_transfer_stenciled(self.handle, C.byref(d)), self.handle)
        return {"staged_chunks": a.value, "bypassed_chunks": b.value, "compressed_results": c.value, "stenciled_results": d.value}

    def measure_peaks(self):
        a, b = C.c_double(0), C.c_double(0)
        check(_lib.lib().fegpu_measure_peaks(self.handle, C.byref(a), C.byref(b)), self.handle)
        return {"dfma_tflops": a.value, "copy_gbs": b.value}


class AbstractSysmatAssembler:
    pass


class SysmatAssemblerSparseGPU(AbstractSysmatAssembler):
    def __init__(self, z=0.0, nomatrixresult=False, ctx=None, device=0):
        if not isinstance(z, float):
            raise TypeError("SysmatAssemblerSparseGPU assembles Float64 matrices only")
        self.ctx = ctx if ctx is not None else GPUContext.default(device)
        self.handle = VP()
        check(_lib.lib().fegpu_asm_create(self.ctx.handle, C.byref(self.handle)), self.ctx.handle)
        self._nomatrixresult = bool(nomatrixresult)
        self._force_init = False
        self._pending_form = None     # (mesh, dofmap) of the last bilform assembly, for raw-COO export
        self._mode = None             # "form" after a bilform call, "generic" inside startassembly/assemble
        self._row_nalldofs = 0
        self._col_nalldofs = 0
        self._last_mesh = None        # device twin of the last form call (owned by the context, see femm.py)

    @property
    def _device_cache(self):
        """The context's device meshes (kept under this name for callers that walk them)."""
        return self.ctx._meshes

    # ---- reference protocol -----------------------------------------------------------------------------
    def eltype(self):
        return np.float64

    def expectedntriples(self, elem_mat_nrows, elem_mat_ncols, n_elem_mats):
        return elem_mat_nrows * elem_mat_ncols * n_elem_mats

    def startassembly(self, elem_mat_nrows, elem_mat_ncols, n_elem_mats, row_nalldofs, col_nalldofs, force_init=False):
        check(_lib.lib().fegpu_startassembly(self.handle, elem_mat_nrows, elem_mat_ncols, n_elem_mats, row_nalldofs, col_nalldofs),
              self.ctx.handle)
        if self._mode != "generic":
            self._row_nalldofs, self._col_nalldofs = int(row_nalldofs), int(col_nalldofs)
        self._mode = "generic"
        self._force_init = force_init
        return self

    def assemble(self, mat, dofnums_row, dofnums_col):
        dr = np.ascontiguousarray(np.asarray(dofnums_row, dtype=np.int64).reshape(-1))
        dc = np.ascontiguousarray(np.asarray(dofnums_col, dtype=np.int64).reshape(-1))
        mat = np.asarray(mat, dtype=np.float64)
        if mat.shape != (dr.size, dc.size):
            raise _lib.FEGPUError(-15, "Wrong size of matrix")
        m = np.asfortranarray(mat)
        check(_lib.lib().fegpu_assemble(self.handle, fptr(m), fptr(dr), dr.size, fptr(dc), dc.size), self.ctx.handle)
        return self

    def _build(self):
        """Device side of makematrix!: after this the CSC (or its current view) is resident on the GPU.  Returns False when
        `nomatrixresult` asks for the dummy zero matrix instead (AssemblyModule.jl:309-317)."""
        L = _lib.lib()
        if self._mode == "generic":
            if self._nomatrixresult:
                return False
            check(L.fegpu_makematrix(self.handle), self.ctx.handle)
            self._mode = "done"
        elif self._mode is None:
            raise _lib.FEGPUError(-17, "makematrix! before any assembly")
        elif self._nomatrixresult:
            return False
        return True

    def makematrix(self, raw=False, out=None):
        """Returns scipy.sparse.csc_matrix (raw=False) or the 1-based (colptr, rowval, nzval, m, n) arrays exactly as they
        would be handed to Julia's SparseMatrixCSC(m, n, colptr, rowval, nzval) (raw=True).  `out` = preallocated
        (colptr, rowval, nzval) host arrays to fill (e.g. pinned), like a shim that reuses its result buffers."""
        if not self._build():
            return self._zeros(raw)
        return self._fetch(raw, out)

    def view(self, row_first, row_last, col_first, col_last, drop_exact_zeros=False):
        """Restrict the device-resident result to A[row_first:row_last, col_first:col_last] (1-based, inclusive) and/or drop
        exact zeros; sizes()/_fetch() then refer to the view.  view_reset() restores the full matrix."""
        check(_lib.lib().fegpu_makematrix_view(self.handle, int(row_first), int(row_last), int(col_first), int(col_last),
                                               1 if drop_exact_zeros else 0), self.ctx.handle)
        return self

    def view_reset(self):
        check(_lib.lib().fegpu_makematrix_view(self.handle, 1, self._row_nalldofs, 1, self._col_nalldofs, 0), self.ctx.handle)
        return self

    # ---- helpers ------------------------------------------------------------------------------------------
    def setnomatrixresult(self, flag):
        self._nomatrixresult = bool(flag)
        return self

    def setforceinit(self, flag):
        self._force_init = bool(flag)
        return self

    def sizes(self):
        m, n, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check(_lib.lib().fegpu_makematrix_sizes(self.handle, C.byref(m), C.byref(n), C.byref(nnz)), self.ctx.handle)
        return m.value, n.value, nnz.value

    def _zeros(self, raw):
        m, n = self._row_nalldofs, self._col_nalldofs
        if raw:
            return np.ones(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0), m, n
        import scipy.sparse as sp
        return sp.csc_matrix((m, n))

    def _fetch(self, raw, out=None):
        m, n, nnz = self.sizes()
        if out is None:
            colptr, rowval, nzval = np.empty(n + 1, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float64)
        else:
            colptr, rowval, nzval = out
            for arr, cnt, dt, what in ((colptr, n + 1, np.int64, "colptr"), (rowval, nnz, np.int64, "rowval"), (nzval, nnz, np.float64, "nzval")):
                # the C side writes cnt entries straight into these buffers: refuse anything that could overflow or reinterpret
                if not isinstance(arr, np.ndarray) or arr.dtype != dt or arr.ndim != 1 or arr.size != cnt or not arr.flags.c_contiguous \
                        or not arr.flags.writeable:
                    raise _lib.FEGPUError(-2, "out[%s] must be a writeable contiguous %s array of length %d" % (what, np.dtype(dt).name, cnt))
        check(_lib.lib().fegpu_makematrix_copy(self.handle, fptr(colptr), fptr(rowval), fptr(nzval)), self.ctx.handle)
        if raw:
            return colptr, rowval, nzval, m, n
        import scipy.sparse as sp
        return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m, n))

    def fetch_values(self, nzval):
        """Only nzval (re-assembly on a cached pattern)."""
        _, _, nnz = self.sizes()
        if not isinstance(nzval, np.ndarray) or nzval.dtype != np.float64 or nzval.size != nnz or not nzval.flags.c_contiguous:
            raise _lib.FEGPUError(-2, "nzval must be a contiguous float64 array of length %d" % nnz)
        check(_lib.lib().fegpu_makematrix_copy_values(self.handle, fptr(nzval)), self.ctx.handle)
        return nzval

    def device_pointers(self):
        c, r, v = VP(), VP(), VP()
        check(_lib.lib().fegpu_makematrix_device(self.handle, C.byref(c), C.byref(r), C.byref(v)), self.ctx.handle)
        return c.value, r.value, v.value

    def coo(self):
        """Raw triplets (I, J, V) of the last bilform assembly, reference emission order (the `nomatrixresult` flow)."""
        if self._pending_form is None:
            raise _lib.FEGPUError(-17, "no bilform assembly to export")
        mesh, dm, ntrip = self._pending_form[:3]
        I, J, V = np.empty(ntrip, np.int64), np.empty(ntrip, np.int64), np.empty(ntrip, np.float64)
        check(_lib.lib().fegpu_coo_copy(self.handle, mesh, dm, fptr(I), fptr(J), fptr(V)), self.ctx.handle)
        return I, J, V

    def timings(self):
        ms = (C.c_double * 4)()
        check(_lib.lib().fegpu_last_timings(self.handle, ms), self.ctx.handle)
        return {"integrate_ms": ms[0], "symbolic_ms": ms[1], "numeric_ms": ms[2], "total_ms": ms[3]}

    def invalidate_patterns(self):
        """Drop every cached sparsity pattern of the context's meshes (the next assembly rebuilds it)."""
        for dm in self.ctx._meshes.values():
            dm.invalidate_patterns()

    def pattern_was_cached(self):
        return bool(_lib.lib().fegpu_pattern_was_cached(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().fegpu_asm_destroy(self.handle)
                self.handle = VP()
        except Exception:
            pass


class AbstractSysvecAssembler:
    pass


class SysvecAssemblerGPU(AbstractSysvecAssembler):
    """SysvecAssembler (AssemblyModule.jl:853-917) on the device: startassembly!(a, row_nalldofs), assemble!(a, vec, dofnums),
    makevector!(a).  linform_dot / distribloads (femm.py) fill it without the element loop on the host.  It can share the
    device mesh / dof-map cache of a matrix assembler (`like=`), so the node -> element adjacency is built once."""

    def __init__(self, z=0.0, ctx=None, device=0, like=None):
        if not isinstance(z, float):
            raise TypeError("SysvecAssemblerGPU assembles Float64 vectors only")
        self.ctx = like.ctx if like is not None else (ctx if ctx is not None else GPUContext.default(device))
        self.handle = VP()
        check(_lib.lib().fegpu_asm_create(self.ctx.handle, C.byref(self.handle)), self.ctx.handle)
        self._last_mesh = None
        self._row_nalldofs = 1  # the reference's blank assembler holds a one-entry buffer (AssemblyModule.jl:861)

    def startassembly(self, row_nalldofs):
        check(_lib.lib().fegpu_vec_startassembly(self.handle, int(row_nalldofs)), self.ctx.handle)
        self._row_nalldofs = int(row_nalldofs)
        return self

    def assemble(self, vec, dofnums):
        d = np.ascontiguousarray(np.asarray(dofnums, dtype=np.int64).reshape(-1))
        v = np.ascontiguousarray(np.asarray(vec, dtype=np.float64).reshape(-1))
        if v.size < d.size:
            raise _lib.FEGPUError(-15, "Wrong size of vector")
        check(_lib.lib().fegpu_vec_assemble(self.handle, fptr(v), fptr(d), d.size), self.ctx.handle)
        return self

    def makevector(self, out=None):
        check(_lib.lib().fegpu_makevector(self.handle), self.ctx.handle)
        return self._fetch(out)

    def _fetch(self, out=None):
        n = C.c_int64(0)
        check(_lib.lib().fegpu_makevector_size(self.handle, C.byref(n)), self.ctx.handle)
        F = out if out is not None else np.empty(n.value, dtype=np.float64)
        if F.size != n.value or F.dtype != np.float64 or not F.flags.c_contiguous:
            raise _lib.FEGPUError(-2, "output vector must be a contiguous Float64 array of length nalldofs")
        check(_lib.lib().fegpu_makevector_copy(self.handle, fptr(F)), self.ctx.handle)
        return F

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().fegpu_asm_destroy(self.handle)
                self.handle = VP()
        except Exception:
            pass


def makevector(a, out=None):
    return a.makevector(out=out)


class SysmatAssemblerSparseSymmGPU(SysmatAssemblerSparseGPU):
    """Drop-in for SysmatAssemblerSparseSymm (src/AssemblyModule.jl:342-583), the reference's DEFAULT assembler when none is
    passed (FEMMBaseModule.jl:1374, 1408, 1543, 1822).  assemble! keeps the lower triangle of each element matrix (:517-530);
    makematrix! returns S + transpose(S) with the doubled diagonal halved (:576-579).  Because SparseArrays' sparse `+` stores
    only non-zero sums, entries that cancel to exactly 0.0 are absent from the result (unlike SysmatAssemblerSparse)."""

    def __init__(self, z=0.0, nomatrixresult=False, ctx=None, device=0):
        super().__init__(z, nomatrixresult, ctx, device)
        check(_lib.lib().fegpu_asm_set_symmetric(self.handle, 1), self.ctx.handle)

    def expectedntriples(self, elem_mat_nrows, elem_mat_ncols, n_elem_mats):
        return int((elem_mat_nrows * elem_mat_ncols + elem_mat_nrows) / 2 * n_elem_mats)

    def assemble(self, mat, dofnums_row, dofnums_col):
        mat = np.asarray(mat, dtype=np.float64)
        nr, nc = np.asarray(dofnums_row).size, np.asarray(dofnums_col).size
        if nr != nc or mat.shape != (nr, nc):
            raise _lib.FEGPUError(-15, "Size mismatch")
        return super().assemble(mat, dofnums_row, dofnums_col)

    def coo(self):
        """Lower-triangle triplets (local i >= j) in the reference's emission order (AssemblyModule.jl:517-530)."""
        I, J, V = super().coo()
        em = int(round(np.sqrt(I.size / max(self._pending_form[3], 1)))) if I.size else 0
        if em == 0:
            return I, J, V
        k = np.arange(em * em)
        keep = np.tile((k % em) >= (k // em), I.size // (em * em))
        return I[keep], J[keep], V[keep]


class SysmatAssemblerSparseDiagGPU(SysmatAssemblerSparseGPU):
    """SysmatAssemblerSparseDiag (AssemblyModule.jl:599-794): only the diagonals of the square element matrices are assembled;
    the result is sparse(I = J = dof, V)."""
    _LUMP = 1

    def __init__(self, z=0.0, nomatrixresult=False, ctx=None, device=0):
        super().__init__(z, nomatrixresult, ctx, device)
        check(_lib.lib().fegpu_asm_set_lumping(self.handle, self._LUMP), self.ctx.handle)

    def expectedntriples(self, elem_mat_nrows, elem_mat_ncols, n_elem_mats):
        return max(elem_mat_nrows, elem_mat_ncols) * n_elem_mats  # :616-621, :963-968

    def assemble(self, mat, dofnums_row, dofnums_col):
        if np.asarray(dofnums_row).size != np.asarray(dofnums_col).size or np.asarray(mat).shape != (np.asarray(dofnums_row).size,) * 2:
            raise _lib.FEGPUError(-15, "Size mismatch")  # :724-729
        return super().assemble(mat, dofnums_row, dofnums_col)


class SysmatAssemblerSparseHRZLumpingSymmGPU(SysmatAssemblerSparseDiagGPU):
    """SysmatAssemblerSparseHRZLumpingSymm (AssemblyModule.jl:943-1141): the diagonal of every element matrix scaled by
    sum(mat) / trace(mat) (Hinton-Rock-Zienkiewicz lumping), assembled into a diagonal matrix."""
    _LUMP = 2


class SysmatAssemblerFFBlock(AbstractSysmatAssembler):
    """Drop-in for SysmatAssemblerFFBlock (src/AssemblyModule.jl:1149-1231): delegates to a wrapped GPU assembler and returns
    the free-free block A[1:row_nfreedofs, 1:col_nfreedofs] of its matrix, cut out on the device (only the block crosses the
    PCIe link)."""

    def __init__(self, row_nfreedofs, col_nfreedofs=None, inner=None, ctx=None, device=0):
        self._a = inner if inner is not None else SysmatAssemblerSparseGPU(0.0, ctx=ctx, device=device)
        if not isinstance(self._a, SysmatAssemblerSparseGPU):
            raise TypeError("SysmatAssemblerFFBlock wraps a GPU assembler (there is no CPU path)")
        self._row_nfreedofs = int(row_nfreedofs)
        self._col_nfreedofs = int(row_nfreedofs if col_nfreedofs is None else col_nfreedofs)

    ctx = property(lambda self: self._a.ctx)

    def eltype(self):
        return self._a.eltype()

    def expectedntriples(self, *args):
        return self._a.expectedntriples(*args)

    def startassembly(self, *args, **kw):
        self._a.startassembly(*args, **kw)
        return self

    def assemble(self, mat, dofnums_row, dofnums_col):
        self._a.assemble(mat, dofnums_row, dofnums_col)
        return self

    def makematrix(self, raw=False, out=None):
        a = self._a
        if not a._build():
            return matrix_blocked_zeros(self._row_nfreedofs, self._col_nfreedofs, raw)
        return matrix_blocked_ff(a, self._row_nfreedofs, self._col_nfreedofs, raw=raw, out=out)


def matrix_blocked_zeros(m, n, raw):
    if raw:
        return np.ones(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0), m, n
    import scipy.sparse as sp
    return sp.csc_matrix((m, n))


def _blocked(a, rows, cols, raw, out):
    """One block of the matrix held on the device by assembler `a` (rows / cols = 1-based inclusive ranges)."""
    if not isinstance(a, SysmatAssemblerSparseGPU):
        raise TypeError("matrix_blocked_* cut blocks from the device-resident matrix of a GPU assembler; there is no CPU path")
    m, n = a._row_nalldofs, a._col_nalldofs
    if rows[1] > m:
        raise _lib.FEGPUError(-2, "The ff block has too many rows")
    if cols[1] > n:
        raise _lib.FEGPUError(-2, "The ff block has too many columns")
    if rows[1] - rows[0] + 1 <= 0 or cols[1] - cols[0] + 1 <= 0:
        return matrix_blocked_zeros(max(rows[1] - rows[0] + 1, 0), max(cols[1] - cols[0] + 1, 0), raw)  # spzeros, :682-686
    a.view(rows[0], rows[1], cols[0], cols[1])
    try:
        return a._fetch(raw, out)
    finally:
        a.view_reset()


def matrix_blocked_ff(a, row_nfreedofs, col_nfreedofs=None, raw=False, out=None):
    """A[1:row_nfreedofs, 1:col_nfreedofs] (src/MatrixUtilityModule.jl:675-688)."""
    cn = row_nfreedofs if col_nfreedofs is None else col_nfreedofs
    if row_nfreedofs > a._row_nalldofs:
        raise _lib.FEGPUError(-2, "The ff block has too many rows")
    if cn > a._col_nalldofs:
        raise _lib.FEGPUError(-2, "The ff block has too many columns")
    return _blocked(a, (1, row_nfreedofs), (1, cn), raw, out)


def matrix_blocked_fd(a, row_nfreedofs, col_nfreedofs=None, raw=False, out=None):
    """A[1:row_nfreedofs, col_nfreedofs+1:end] (src/MatrixUtilityModule.jl:707-724)."""
    cn = row_nfreedofs if col_nfreedofs is None else col_nfreedofs
    return _blocked(a, (1, row_nfreedofs), (cn + 1, a._col_nalldofs), raw, out)


def matrix_blocked_df(a, row_nfreedofs, col_nfreedofs=None, raw=False, out=None):
    """A[row_nfreedofs+1:end, 1:col_nfreedofs] (src/MatrixUtilityModule.jl:743-760)."""
    cn = row_nfreedofs if col_nfreedofs is None else col_nfreedofs
    return _blocked(a, (row_nfreedofs + 1, a._row_nalldofs), (1, cn), raw, out)


def matrix_blocked_dd(a, row_nfreedofs, col_nfreedofs=None, raw=False, out=None):
    """A[row_nfreedofs+1:end, col_nfreedofs+1:end] (src/MatrixUtilityModule.jl:779-793)."""
    cn = row_nfreedofs if col_nfreedofs is None else col_nfreedofs
    return _blocked(a, (row_nfreedofs + 1, a._row_nalldofs), (cn + 1, a._col_nalldofs), raw, out)


def startassembly(a, *args, **kw):
    return a.startassembly(*args, **kw)


def assemble(a, mat, dofnums_row, dofnums_col):
    return a.assemble(mat, dofnums_row, dofnums_col)


def makematrix(a, raw=False, out=None):
    return a.makematrix(raw=raw, out=out)


def setnomatrixresult(a, flag):
    return a.setnomatrixresult(flag)


def expectedntriples(a, nr, nc, n):
    return a.expectedntriples(nr, nc, n)
