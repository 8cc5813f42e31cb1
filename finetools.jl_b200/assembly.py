"""SysmatAssemblerSparseGPU: drop-in for FinEtools' SysmatAssemblerSparse (src/AssemblyModule.jl:88-329) whose
element loop, triplet storage and COO->CSC conversion run on a B200 through libfinegpu.so.

Protocol mirrored: startassembly! (:209-238), assemble! (:250-282), makematrix! (:304-329), setnomatrixresult (:29),
expectedntriples (:54-59), eltype (:27).  Julia's `!` is dropped from the names.  Error strings are the reference's.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VP, check, fptr


class GPUContext:
    """One CUDA device + stream; owner of every device handle created through it."""
    _default = {}

    def __init__(self, device=0, stream=None):
        self.handle = VP()
        check(_lib.lib().fegpu_create(C.byref(self.handle), int(device)))
        self.device = int(device)
        if stream is not None:
            self.set_stream(stream)

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def set_stream(self, cuda_stream_ptr):
        check(_lib.lib().fegpu_set_stream(self.handle, VP(int(cuda_stream_ptr))), self.handle)

    def set_async(self, on):
        check(_lib.lib().fegpu_set_async(self.handle, 1 if on else 0), self.handle)

    def synchronize(self):
        check(_lib.lib().fegpu_synchronize(self.handle), self.handle)

    def launch_count(self):
        return int(_lib.lib().fegpu_launch_count(self.handle))

    def transfer_stats(self):
        a, b = C.c_int64(0), C.c_int64(0)
        check(_lib.lib().fegpu_transfer_stats(self.handle, C.byref(a), C.byref(b)), self.handle)
        return {"staged_chunks": a.value, "bypassed_chunks": b.value}

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
        self._device_cache = {}       # (mesh/dofmap handles keyed by the host objects), see femm.py

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

    def makematrix(self, raw=False, out=None):
        """Returns scipy.sparse.csc_matrix (raw=False) or the 1-based (colptr, rowval, nzval, m, n) arrays exactly as they
        would be handed to Julia's SparseMatrixCSC(m, n, colptr, rowval, nzval) (raw=True).  `out` = preallocated
        (colptr, rowval, nzval) host arrays to fill (e.g. pinned), like a shim that reuses its result buffers."""
        L = _lib.lib()
        if self._mode == "generic":
            if self._nomatrixresult:
                # the reference returns spzeros and keeps the triplets (AssemblyModule.jl:309-317); same here
                return self._zeros(raw)
            check(L.fegpu_makematrix(self.handle), self.ctx.handle)
            self._mode = "done"
        elif self._mode is None:
            raise _lib.FEGPUError(-17, "makematrix! before any assembly")
        elif self._nomatrixresult:
            return self._zeros(raw)
        return self._fetch(raw, out)

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
        check(_lib.lib().fegpu_makematrix_copy(self.handle, fptr(colptr), fptr(rowval), fptr(nzval)), self.ctx.handle)
        if raw:
            return colptr, rowval, nzval, m, n
        import scipy.sparse as sp
        return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m, n))

    def fetch_values(self, nzval):
        """Only nzval (re-assembly on a cached pattern)."""
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
        mesh, dm, ntrip = self._pending_form
        I, J, V = np.empty(ntrip, np.int64), np.empty(ntrip, np.int64), np.empty(ntrip, np.float64)
        check(_lib.lib().fegpu_coo_copy(self.handle, mesh, dm, fptr(I), fptr(J), fptr(V)), self.ctx.handle)
        return I, J, V

    def timings(self):
        ms = (C.c_double * 4)()
        check(_lib.lib().fegpu_last_timings(self.handle, ms), self.ctx.handle)
        return {"integrate_ms": ms[0], "symbolic_ms": ms[1], "numeric_ms": ms[2], "total_ms": ms[3]}

    def invalidate_patterns(self):
        """Drop every cached sparsity pattern held for this assembler's meshes (next assembly rebuilds it)."""
        for dm in self._device_cache.values():
            for _, _, h in dm.dofmaps:
                check(_lib.lib().fegpu_pattern_invalidate(h), self.ctx.handle)

    def pattern_was_cached(self):
        return bool(_lib.lib().fegpu_pattern_was_cached(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().fegpu_asm_destroy(self.handle)
                self.handle = VP()
        except Exception:
            pass


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
