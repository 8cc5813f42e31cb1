"""FEMMBase and its bilinear forms, dispatched to the GPU when the assembler is a SysmatAssemblerSparseGPU.

Mirrors src/FEMMBaseModule.jl: FEMMBase :72-84, bilform_dot :1335-1366, innerproduct :1388-1401, bilform_diffusion
:1462-1535, bilform_convection :1583-1625, bilform_div_grad :1672-1713, bilform_lin_elastic :1774-1813.  User code keeps the reference's call shape
    K = bilform_diffusion(femm, assembler, geom, u, DataCache(kappa))
The eligibility checks of SURVEY.md section 8(b) raise instead of silently falling back to a CPU loop.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VP, FEGPUError, check, fptr
from .assembly import SysmatAssemblerFFBlock, SysmatAssemblerSparseGPU, SysvecAssemblerGPU
from .datacache import DataCache
from .integdomain import IntegDomain, integrationdata


class DeforModelRed3D:
    """Marker type (src/DeforModelRedModule.jl:30); nstressstrain = 6 (:72-76)."""
    nstressstrain = 6


class CSys:
    """Material coordinate system (src/CSysModule.jl): CSys(dim) = identity (the FEMMBase default, FEMMBaseModule.jl:82-84),
    CSys(csmat) = a given constant matrix (:133-144).  Position-dependent systems (a compute callback) are not GPU-eligible."""

    def __init__(self, sdim=3, mdim=None, isidentity=True, csmat=None):
        if isinstance(sdim, np.ndarray):  # CSys(csmat)
            csmat, sdim = sdim, sdim.shape[0]
        self.isconstant = True
        if csmat is not None:
            self.csmat = np.array(csmat, dtype=np.float64)
            self.sdim, self.mdim = self.csmat.shape
            self.isidentity = False
        else:
            self.isidentity = bool(isidentity)
            self.sdim, self.mdim = sdim, sdim if mdim is None else mdim
            self.csmat = np.eye(self.sdim, self.mdim) if self.isidentity else None


class FEMMBase:
    def __init__(self, integdomain, mcsys=None):
        if not isinstance(integdomain, IntegDomain):
            raise TypeError("FEMMBase needs an IntegDomain")
        self.integdomain = integdomain
        self.mcsys = mcsys if mcsys is not None else CSys(integdomain.fes.mdim)

    def finite_elements(self):
        return self.integdomain.fes


def _frozen(arr):
    """True when nobody can have edited `arr` in place since we last saw it (read-only array, or a read-only view chain)."""
    return isinstance(arr, np.ndarray) and not arr.flags.writeable


class _DeviceMesh:
    """Device twin of (fes, geom, rule), owned by the GPUContext (one per FESet connectivity, shared by every assembler on that
    device, so the reference's idiom of a fresh assembler per call -- FEMMBaseModule.jl:1374, 1408 -- still finds the uploaded
    mesh and the cached sparsity pattern).  Coordinates are refreshed on every form call.

    Cache keys are content-safe: a hit is accepted on object identity only when the host array is read-only (what
    numberdofs / slab_owner / pointpartitioning hand out), otherwise the arrays are compared in full."""

    def __init__(self, ctx, fes, geom):
        self.ctx = ctx
        self.handle = VP()
        xyz = _lib.colmajor_f64(geom.values)
        self.nnodes, self.sdim = xyz.shape
        conn = np.ascontiguousarray(fes.conn, dtype=np.int64)
        check(_lib.lib().fegpu_mesh_upload(ctx.handle, fes.etype, conn.shape[0], fptr(conn), self.nnodes, self.sdim, fptr(xyz),
                                           C.byref(self.handle)), ctx.handle)
        self.conn_ref = fes.conn    # keeps the array (and therefore its id) alive while the twin exists
        self.rule_key = None
        self.dofmaps = []           # [(host copy or frozen original of dofnums, nalldofs, handle)]
        self._dofkeys = {}          # handle -> (id(u), numbering version, id(dofnums)) it was last matched with
        self.partition_key = None
        self._owner_copy = None
        self.h2d_bytes_last = xyz.nbytes

    def update_geometry(self, geom):
        """Coordinates host -> device; a partitioned mesh ships only the node window of its active elements."""
        xyz = _lib.colmajor_f64(geom.values)
        if xyz.shape != (self.nnodes, self.sdim):
            raise FEGPUError(-2, "geometry field changed shape; build a new FESet / assembler")
        if self.partition_key is None:
            check(_lib.lib().fegpu_geom_update(self.handle, fptr(xyz)), self.ctx.handle)
            self.h2d_bytes_last = xyz.nbytes
        else:
            check(_lib.lib().fegpu_geom_update_window(self.handle, fptr(xyz)), self.ctx.handle)
            lo, hi, _ = self.window()
            self.h2d_bytes_last = (hi - lo) * self.sdim * 8

    def window(self):
        lo, hi, na = C.c_int64(), C.c_int64(), C.c_int64()
        check(_lib.lib().fegpu_mesh_window(self.handle, C.byref(lo), C.byref(hi), C.byref(na)), self.ctx.handle)
        return lo.value, hi.value, na.value

    def set_rule(self, integdomain):
        check(_lib.lib().fegpu_otherdimension_set(self.handle, float(integdomain.otherdimension)), self.ctx.handle)
        rule = integdomain.integration_rule
        # keyed on the rule's own numbers, not on its identity (ids are recycled; two rules may share a point count)
        key = (rule.npts, np.asarray(rule.param_coords, dtype=np.float64).tobytes(), np.asarray(rule.weights, dtype=np.float64).tobytes())
        if key == self.rule_key:
            return
        npts, Ns, gradNparams, w, _ = integrationdata(integdomain)
        N = np.ascontiguousarray(np.stack([n.reshape(-1) for n in Ns]))                       # [npts][nne]
        dN = np.ascontiguousarray(np.stack([np.asfortranarray(g).T for g in gradNparams]))    # [npts][mdim][nne]
        ww = np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(-1))
        check(_lib.lib().fegpu_rule_set(self.handle, npts, fptr(N), fptr(dN), fptr(ww)), self.ctx.handle)
        self.rule_key = key

    def set_partition(self, node_owner, my_rank):
        if node_owner is None:
            if self.partition_key is not None:
                check(_lib.lib().fegpu_partition_set(self.handle, None, 0), self.ctx.handle)
                self.partition_key, self._owner_copy = None, None
            return
        own = np.ascontiguousarray(node_owner, dtype=np.int32)
        if own.size != self.nnodes:
            raise FEGPUError(-2, "node_owner must have one entry per node")
        if self.partition_key is not None and self.partition_key[1] == int(my_rank):
            same_obj = self.partition_key[0] == id(node_owner) and _frozen(node_owner)
            if same_obj or np.array_equal(self._owner_copy, own):
                return
        check(_lib.lib().fegpu_partition_set(self.handle, fptr(own), int(my_rank)), self.ctx.handle)
        # a frozen array is kept by reference (its id stays valid); anything else is copied and compared by content next time
        self._owner_copy = node_owner if _frozen(node_owner) else own.copy()
        self.partition_key = (id(node_owner), int(my_rank))

    def dofmap(self, u):
        dn = u.dofnums
        nall = u.nalldofs()
        key = (id(u), getattr(u, "_dofver", None), id(dn))
        for host, n, h in self.dofmaps:
            if n != nall or host.shape != dn.shape:
                continue
            # identity + numbering version is trusted only for a read-only dofnums array (numberdofs freezes it: in-place edits
            # raise, a replaced array has another id); anything else is compared in full
            if (key[1] is not None and self._dofkeys.get(h.value) == key and _frozen(dn)) or np.array_equal(host, dn):
                self._dofkeys[h.value] = key
                return h
        if dn.shape[0] != self.nnodes:
            raise FEGPUError(-2, "field u and geometry have different node counts")
        d = _lib.colmajor_i64(dn)
        h = VP()
        check(_lib.lib().fegpu_dofmap_upload(self.ctx.handle, self.handle, dn.shape[1], fptr(d), nall, nall, C.byref(h)), self.ctx.handle)
        self.dofmaps.append((dn if _frozen(dn) else dn.copy(order="K"), nall, h))
        self._dofkeys[h.value] = key
        if len(self.dofmaps) > 4:
            _, _, old = self.dofmaps.pop(0)
            self._dofkeys.pop(old.value, None)
            _lib.lib().fegpu_dofmap_destroy(old)
        return h

    def invalidate_patterns(self):
        for _, _, h in self.dofmaps:
            check(_lib.lib().fegpu_pattern_invalidate(h), self.ctx.handle)

    def destroy(self):
        for _, _, h in self.dofmaps:
            _lib.lib().fegpu_dofmap_destroy(h)
        self.dofmaps = []
        if self.handle:
            _lib.lib().fegpu_mesh_destroy(self.handle)
            self.handle = VP()


def _device_mesh(assembler, fes, geom):
    """The context's twin of this FESet (created on first use, least recently used twins are destroyed beyond
    GPUContext.MAX_MESHES).  The geometry is NOT refreshed here: _prepare does it after the partition is known."""
    ctx = assembler.ctx
    cache = ctx._meshes
    key = id(fes.conn)
    dm = cache.get(key)
    if dm is not None and dm.conn_ref is fes.conn:
        cache.move_to_end(key)
        return dm, False
    if dm is not None:
        dm.destroy()
        del cache[key]
    dm = _DeviceMesh(ctx, fes, geom)
    cache[key] = dm
    while len(cache) > ctx.MAX_MESHES:
        _, old = cache.popitem(last=False)
        old.destroy()
    return dm, True


def _inner(assembler):
    """The GPU assembler doing the device work (SysmatAssemblerFFBlock delegates to the one it wraps, AssemblyModule.jl:1149)."""
    return assembler._a if isinstance(assembler, SysmatAssemblerFFBlock) else assembler


def _eligible(self, assembler, geom, u, cf):
    if not isinstance(_inner(assembler), SysmatAssemblerSparseGPU):
        raise TypeError("this package provides the GPU assembler path only (SysmatAssemblerSparseGPU); there is no CPU loop")
    if not isinstance(cf, DataCache):
        raise TypeError("coefficient must be a constant DataCache")
    if not self.mcsys.isconstant or self.mcsys.csmat is None:
        raise FEGPUError(-2, "only constant material coordinate systems (CSys(dim) or CSys(csmat)) are GPU-eligible")
    if self.integdomain.axisymmetric:
        raise FEGPUError(-2, "axisymmetric integration domains are not GPU-eligible")
    if geom.values.dtype != np.float64 or u.dofnums.dtype != np.int64:
        raise FEGPUError(-2, "geom must be Float64 and dofnums Int64")


def _prepare(self, assembler, geom, u, node_owner=None, my_rank=0):
    fes = self.integdomain.fes
    assembler = _inner(assembler)
    dmesh, fresh = _device_mesh(assembler, fes, geom)
    dmesh.set_rule(self.integdomain)
    if self.mcsys.isidentity:
        check(_lib.lib().fegpu_csys_set(dmesh.handle, None), assembler.ctx.handle)
    else:
        if self.mcsys.csmat.shape != (fes.mdim, fes.mdim) or geom.values.shape[1] != fes.mdim:
            raise FEGPUError(-2, "the material coordinate system matrix must be sdim x mdim with sdim == mdim")
        rm = np.asfortranarray(self.mcsys.csmat)
        check(_lib.lib().fegpu_csys_set(dmesh.handle, fptr(rm)), assembler.ctx.handle)
    dmesh.set_partition(node_owner, my_rank)
    if not fresh:
        dmesh.update_geometry(geom)  # the upload of a new twin already carried the coordinates
    dof = dmesh.dofmap(u)
    assembler._last_mesh = dmesh
    return fes, dmesh, dof


def _finish(assembler, fes, dmesh, dof, u, raw, out=None):
    elmdim = fes.nne * u.ndofs()
    a = _inner(assembler)
    a._mode = "form"
    a._row_nalldofs = a._col_nalldofs = u.nalldofs()
    a._pending_form = (dmesh.handle, dof, fes.count() * elmdim * elmdim if dmesh.partition_key is None else None, fes.count())
    return assembler.makematrix(raw=raw, out=out)


def bilform_diffusion(self, assembler, geom, u, cf, raw=False, node_owner=None, my_rank=0, out=None):
    """K_ij = int grad(N_i) . kappa . grad(N_j): scalar DataCache -> _iso path, matrix -> _general path."""
    _eligible(self, assembler, geom, u, cf)
    if u.ndofs() != 1:
        raise FEGPUError(-15, "Wrong size of matrix")  # add_gkgt_ut_only! asserts nne == Kedim
    fes, dmesh, dof = _prepare(self, assembler, geom, u, node_owner, my_rank)
    if fes.mdim != geom.values.shape[1]:
        raise FEGPUError(-2, "bilform_diffusion needs space dimension == manifold dimension")
    kap = cf.data
    if kap.ndim == 0:
        kind, k = 0, np.array([float(kap)])
    else:
        if kap.shape != (fes.mdim, fes.mdim):
            raise FEGPUError(-2, "conductivity matrix must be mdim x mdim")
        kind, k = 1, np.asfortranarray(kap)
    with assembler.ctx.queued_forms(not _inner(assembler)._nomatrixresult):
        check(_lib.lib().fegpu_bilform_diffusion(dmesh.handle, dof, kind, fptr(k), _inner(assembler).handle), assembler.ctx.handle)
        return _finish(assembler, fes, dmesh, dof, u, raw, out)


def bilform_lin_elastic(self, assembler, geom, u, mr, cf, raw=False, node_owner=None, my_rank=0, out=None):
    """K = int B' C B with the 3-D strain-displacement matrix (mr must be DeforModelRed3D)."""
    _eligible(self, assembler, geom, u, cf)
    if mr is not DeforModelRed3D and not isinstance(mr, DeforModelRed3D):
        raise FEGPUError(-2, "only DeforModelRed3D is GPU-eligible")
    if u.ndofs() != 3 or geom.values.shape[1] != 3:
        raise FEGPUError(-2, "Wrong dimensions")
    Cm = cf.data
    if Cm.shape != (6, 6):
        raise FEGPUError(-2, "material stiffness must be 6 x 6")
    fes, dmesh, dof = _prepare(self, assembler, geom, u, node_owner, my_rank)
    Cf = np.asfortranarray(Cm)
    with assembler.ctx.queued_forms(not _inner(assembler)._nomatrixresult):
        check(_lib.lib().fegpu_bilform_lin_elastic(dmesh.handle, dof, fptr(Cf), _inner(assembler).handle), assembler.ctx.handle)
        return _finish(assembler, fes, dmesh, dof, u, raw, out)


def bilform_dot(self, assembler, geom, u, cf, m=3, raw=False, node_owner=None, my_rank=0, out=None):
    """M_ij = int N_i c N_j over the m-dimensional manifold Jacobian."""
    _eligible(self, assembler, geom, u, cf)
    ndn = u.ndofs()
    c = cf.data
    if c.ndim == 0:
        c = c.reshape(1, 1)
    if c.shape != (ndn, ndn):
        raise FEGPUError(-2, "coefficient must be ndn x ndn")
    fes, dmesh, dof = _prepare(self, assembler, geom, u, node_owner, my_rank)
    cfm = np.asfortranarray(c)
    with assembler.ctx.queued_forms(not _inner(assembler)._nomatrixresult):
        check(_lib.lib().fegpu_bilform_dot(dmesh.handle, dof, fptr(cfm), int(m), float(self.integdomain.otherdimension), _inner(assembler).handle),
              assembler.ctx.handle)
        return _finish(assembler, fes, dmesh, dof, u, raw, out)


def bilform_convection(self, assembler, geom, u, Q, rhof, raw=False, node_owner=None, my_rank=0, out=None):
    """K_pr = int N_p (u . grad N_r): u = nodal convective velocity field, Q = scalar field that numbers the dofs
    (FEMMBaseModule.jl:1583-1625).  Non-symmetric: the symmetric assembler refuses it."""
    _eligible(self, assembler, geom, Q, rhof)
    if Q.ndofs() != 1:
        raise FEGPUError(-15, "Wrong size of matrix")
    fes, dmesh, dof = _prepare(self, assembler, geom, Q, node_owner, my_rank)
    sdim = geom.values.shape[1]
    if fes.mdim != sdim or u.values.shape != (geom.values.shape[0], sdim):
        raise FEGPUError(-2, "bilform_convection needs a velocity component per space dimension and sdim == manifold dimension")
    uv = _lib.colmajor_f64(u.values)
    with assembler.ctx.queued_forms(not _inner(assembler)._nomatrixresult):
        check(_lib.lib().fegpu_bilform_convection(dmesh.handle, dof, fptr(uv), float(rhof.data), _inner(assembler).handle), assembler.ctx.handle)
        return _finish(assembler, fes, dmesh, dof, Q, raw, out)


def bilform_div_grad(self, assembler, geom, u, viscf, raw=False, node_owner=None, my_rank=0, out=None):
    """G = int mu (grad w : grad u + grad w : grad u^T), vector field with one dof per space dimension
    (FEMMBaseModule.jl:1672-1713)."""
    _eligible(self, assembler, geom, u, viscf)
    fes, dmesh, dof = _prepare(self, assembler, geom, u, node_owner, my_rank)
    sdim = geom.values.shape[1]
    if fes.mdim != sdim or u.ndofs() != sdim:
        raise FEGPUError(-2, "bilform_div_grad needs one dof per space dimension and sdim == manifold dimension")
    with assembler.ctx.queued_forms(not _inner(assembler)._nomatrixresult):
        check(_lib.lib().fegpu_bilform_div_grad(dmesh.handle, dof, float(viscf.data), _inner(assembler).handle), assembler.ctx.handle)
        return _finish(assembler, fes, dmesh, dof, u, raw, out)


def bilform_masslike(self, assembler, geom, phi, cf, m=3, raw=False, out=None):
    """int chi c phi with chi the indicator function of each element (FEMMBaseModule.jl:1865-1912): a rectangular
    (count(fes) * ndn) x nalldofs(phi) matrix, element i owning rows (i-1)*ndn+1 .. i*ndn."""
    _eligible(self, assembler, geom, phi, cf)
    ndn = phi.ndofs()
    c = cf.data
    if c.ndim == 0:
        c = c.reshape(1, 1)
    if c.shape != (ndn, ndn):
        raise FEGPUError(-2, "coefficient must be ndn x ndn")
    fes, dmesh, dof = _prepare(self, assembler, geom, phi)
    cfm = np.asfortranarray(c)
    a = _inner(assembler)
    check(_lib.lib().fegpu_bilform_masslike(dmesh.handle, dof, fptr(cfm), int(m), float(self.integdomain.otherdimension), a.handle),
          assembler.ctx.handle)
    a._mode = "form"
    a._row_nalldofs, a._col_nalldofs = fes.count() * ndn, phi.nalldofs()
    a._pending_form = None
    return assembler.makematrix(raw=raw, out=out)


class ForceIntensity:
    """Constant distributed-load intensity (src/ForceIntensityModule.jl: the constant constructors wrap the vector in a
    DataCache); only constant intensities are GPU-eligible."""

    def __init__(self, force):
        self._cache = DataCache(np.atleast_1d(np.asarray(force, dtype=np.float64)))


def linform_dot(self, assembler, geom, P, f, m, node_owner=None, my_rank=0, out=None):
    """F_i = int N_i f over the m-dimensional manifold (FEMMBaseModule.jl:1207-1244); f: constant DataCache of ndn values."""
    if not isinstance(assembler, SysvecAssemblerGPU):
        raise TypeError("this package provides the GPU vector assembler only (SysvecAssemblerGPU); there is no CPU loop")
    if not isinstance(f, DataCache):
        raise TypeError("the load must be a constant DataCache")
    if self.integdomain.axisymmetric:
        raise FEGPUError(-2, "axisymmetric integration domains are not GPU-eligible")
    if geom.values.dtype != np.float64 or P.dofnums.dtype != np.int64:
        raise FEGPUError(-2, "geom must be Float64 and dofnums Int64")
    force = np.ascontiguousarray(np.atleast_1d(f.data).astype(np.float64).reshape(-1))
    if force.size != P.ndofs():
        raise FEGPUError(-2, "the load needs one component per degree of freedom of a node")
    fes, dmesh, dof = _prepare(self, assembler, geom, P, node_owner, my_rank)
    check(_lib.lib().fegpu_linform_dot(dmesh.handle, dof, fptr(force), int(m), float(self.integdomain.otherdimension), assembler.handle),
          assembler.ctx.handle)
    assembler._row_nalldofs = P.nalldofs()
    return assembler._fetch(out)


def distribloads(self, assembler, geom, P, fi, m, **kw):
    """FEMMBaseModule.jl:1277-1286: linform_dot with the ForceIntensity's cache."""
    if not isinstance(fi, ForceIntensity):
        raise TypeError("distribloads needs a ForceIntensity")
    return linform_dot(self, assembler, geom, P, fi._cache, m, **kw)


def innerproduct(self, assembler, geom, afield, raw=False):
    """bilform_dot with the identity coefficient (FEMMBaseModule.jl:1388-1401); Diagonal{Bool} densified to Float64."""
    return bilform_dot(self, assembler, geom, afield, DataCache(np.eye(afield.ndofs())), m=3, raw=raw)
