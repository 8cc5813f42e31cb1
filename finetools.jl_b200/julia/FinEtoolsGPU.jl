# FinEtoolsGPU.jl -- drop-in GPU assembler for FinEtools.jl (v8.2.x) backed by libfinegpu.so (include/fegpu.h).
#
# NOT EXECUTED in the build environment (no Julia on the image).  It is the reference-side binding a maintainer adds:
# every `ccall` below has a line-for-line twin in finetools.jl_b200/_lib.py + assembly.py + femm.py, which IS what the
# test-suite and bench.py run; tests/test_host_logic.py checks every ccall statically against the ABI and checks that the
# eligibility rules below are present.  parity.jl (next to this file) compares makematrix! outputs of the two assemblers.
#
#   using FinEtools, FinEtoolsGPU
#   K = bilform_diffusion(femm, SysmatAssemblerSparseGPU(0.0), geom, u, DataCache(kappa))   # unchanged call shape
#
# Device state (CUDA context, uploaded meshes, dof maps, cached sparsity patterns) belongs to the DEVICE, not to an assembler:
# the reference's idiom builds an assembler per call (FEMMBaseModule.jl:1374, 1408, 1543, 1822), and such a call must still
# find the mesh on the GPU and hit the cached-pattern re-assembly.
module FinEtoolsGPU

using FinEtools
using SparseArrays
import FinEtools.AssemblyModule: AbstractSysmatAssembler, AbstractSysvecAssembler, SysmatAssemblerFFBlock, startassembly!, assemble!,
    makematrix!, makevector!, eltype, expectedntriples
import FinEtools.FEMMBaseModule: bilform_diffusion, bilform_lin_elastic, bilform_dot, bilform_convection, bilform_div_grad,
    bilform_masslike, linform_dot, FEMMBase, finite_elements
using FinEtools.IntegDomainModule: integrationdata, otherdimensionunity
using FinEtools.DeforModelRedModule: DeforModelRed3D
using FinEtools.CSysModule: csmat

export SysmatAssemblerSparseGPU, SysmatAssemblerSparseSymmGPU, SysmatAssemblerSparseDiagGPU, SysmatAssemblerSparseHRZLumpingSymmGPU,
    SysvecAssemblerGPU, gpu_matrix_blocked, gpu_release_cache, gpu_coo, gpu_device_pointers, gpu_release_meshes

const LIB = get(ENV, "FEGPU_LIB", joinpath(@__DIR__, "..", "libfinegpu.so"))
const MAX_MESHES = 4   # device twins kept per device; the least recently used one is destroyed beyond this

_etype(::FESetT3) = 1; _etype(::FESetQ4) = 2; _etype(::FESetT4) = 3; _etype(::FESetT10) = 4
_etype(::FESetH8) = 5; _etype(::FESetH20) = 6; _etype(::FESetH27) = 7
_etype(fes) = error("Element type $(typeof(fes)) is not GPU-eligible")

function _check(status::Int32, ctx::Ptr{Cvoid} = C_NULL)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:fegpu_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    error(msg)   # same strings as AssemblyModule.jl:265-273 ("Row degree of freedom > size", ...)
end

# ---- per-device state -------------------------------------------------------------------------------------------------------
mutable struct DofTwin
    dofnums::Matrix{Int64}          # content key (copy): an in-place renumbering is a different matrix
    nall::Int
    handle::Ptr{Cvoid}
    colptr::Vector{Int64}           # pattern arrays of the last full fetch: shared by the results of cached re-assemblies
    rowval::Vector{Int64}
    have_pattern::Bool
end

mutable struct MeshTwin
    conn::Any                       # fes.conn: identity key, kept alive while the twin exists
    handle::Ptr{Cvoid}
    rule::Any                       # (npts, N, dN, w) last uploaded with fegpu_rule_set
    dofs::Vector{DofTwin}
    owner::Any                      # (copy of node_owner, rank) of the current partition, or nothing
end

mutable struct DeviceState
    ctx::Ptr{Cvoid}
    meshes::Vector{MeshTwin}        # most recently used last
end

const _DEVICES = Dict{Int,DeviceState}()

function _device_state(device::Integer)
    get!(_DEVICES, Int(device)) do
        ctx = Ref{Ptr{Cvoid}}(C_NULL)
        _check(ccall((:fegpu_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), ctx, device))
        DeviceState(ctx[], MeshTwin[])
    end
end

function _destroy!(t::MeshTwin)
    foreach(d -> ccall((:fegpu_dofmap_destroy, LIB), Int32, (Ptr{Cvoid},), d.handle), t.dofs)
    ccall((:fegpu_mesh_destroy, LIB), Int32, (Ptr{Cvoid},), t.handle)
    empty!(t.dofs)
    return nothing
end

"""
    gpu_release_meshes(device = 0)

Destroy every device mesh / dof map / cached pattern of the device (assemblers keep the results they already hold).
"""
function gpu_release_meshes(device::Integer = 0)
    st = _device_state(device)
    foreach(_destroy!, st.meshes)
    empty!(st.meshes)
    return nothing
end

"""
    SysmatAssemblerSparseGPU{T} <: AbstractSysmatAssembler

Same protocol as `SysmatAssemblerSparse` (AssemblyModule.jl:88-329); the element loop of the bilinear forms and the COO -> CSC
conversion run on the GPU.  Keywords: `device`; `node_owner` (one 0-based rank per node) + `rank` make the assembler build the
row block of the nodes this rank owns (multi-GPU split, one process or task per GPU); `pinned_results = true` keeps page-locked
result buffers per dof map and returns matrices that alias them (valid until the next assembly on that mesh): DMA at link speed.
"""
mutable struct SysmatAssemblerSparseGPU{T} <: AbstractSysmatAssembler
    ctx::Ptr{Cvoid}
    handle::Ptr{Cvoid}
    device::Int
    node_owner::Union{Nothing,Vector{Int32}}
    rank::Int32
    pinned_results::Bool
    _row_nalldofs::Int
    _col_nalldofs::Int
    _nomatrixresult::Bool
    _force_init::Bool
    _generic::Bool
    _last::Any                      # (MeshTwin, DofTwin) of the last form call
end

function SysmatAssemblerSparseGPU(z::Float64 = 0.0, nomatrixresult = false; device = 0, node_owner = nothing, rank = 0,
    pinned_results = false)
    st = _device_state(device)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fegpu_asm_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), st.ctx, h), st.ctx)
    own = node_owner === nothing ? nothing : Vector{Int32}(node_owner)
    a = SysmatAssemblerSparseGPU{Float64}(st.ctx, h[], Int(device), own, Int32(rank), pinned_results, 0, 0, nomatrixresult, false, false, nothing)
    finalizer(x -> ccall((:fegpu_asm_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), a)   # the context outlives every assembler
    return a
end

eltype(::SysmatAssemblerSparseGPU{T}) where {T} = T

# ---- generic protocol (any caller of startassembly!/assemble!/makematrix! keeps working) ------------------------------
function startassembly!(self::SysmatAssemblerSparseGPU, elem_mat_nrows::IT, elem_mat_ncols::IT, n_elem_mats::IT,
    row_nalldofs::IT, col_nalldofs::IT; force_init = false) where {IT<:Integer}
    _check(ccall((:fegpu_startassembly, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64),
            self.handle, elem_mat_nrows, elem_mat_ncols, n_elem_mats, row_nalldofs, col_nalldofs), self.ctx)
    self._row_nalldofs, self._col_nalldofs, self._generic = row_nalldofs, col_nalldofs, true
    self._last = nothing
    return self
end

function assemble!(self::SysmatAssemblerSparseGPU, mat::MBT, dofnums_row::CIT, dofnums_col::CIT) where {MBT,CIT}
    nrows, ncolumns = length(dofnums_row), length(dofnums_col)
    size(mat) == (nrows, ncolumns) || error("Wrong size of matrix")
    m = Matrix{Float64}(mat); dr = Vector{Int64}(vec(dofnums_row)); dc = Vector{Int64}(vec(dofnums_col))
    GC.@preserve m dr dc _check(ccall((:fegpu_assemble, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Int64}, Int64), self.handle, m, dr, nrows, dc, ncolumns), self.ctx)
    return self
end

function _sizes(self::SysmatAssemblerSparseGPU)
    m, n, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fegpu_makematrix_sizes, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), self.handle, m, n, nnz), self.ctx)
    return m[], n[], nnz[]
end

# page-locked vectors (fegpu_host_alloc) wrapped as Julia arrays; freed by a finalizer on the wrapper
function _pinned(::Type{T}, n::Integer) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fegpu_host_alloc, LIB), Int32, (Ref{Ptr{Cvoid}}, Int64), p, max(n, 1) * sizeof(T)))
    v = unsafe_wrap(Array, Ptr{T}(p[]), n; own = false)
    finalizer(_ -> ccall((:fegpu_host_free, LIB), Int32, (Ptr{Cvoid},), p[]), v)
    return v
end

function makematrix!(self::SysmatAssemblerSparseGPU)
    if self._nomatrixresult
        return spzeros(self._row_nalldofs, self._col_nalldofs)    # AssemblyModule.jl:309-317; the triplets: gpu_coo(assembler)
    end
    if self._generic
        _check(ccall((:fegpu_makematrix, LIB), Int32, (Ptr{Cvoid},), self.handle), self.ctx)
        self._generic = false
    end
    m, n, nnz = _sizes(self)
    alloc = self.pinned_results ? _pinned : (T, k) -> Vector{T}(undef, k)
    dt = self._last === nothing ? nothing : self._last[2]
    cached = ccall((:fegpu_pattern_was_cached, LIB), Int32, (Ptr{Cvoid},), self.handle) == 1
    if dt !== nothing && cached && dt.have_pattern && length(dt.colptr) == n + 1 && length(dt.rowval) == nnz
        # re-assembly on the cached pattern: colptr / rowval of the previous result are still the pattern -- only nzval crosses
        nzval = alloc(Float64, nnz)
        GC.@preserve nzval _check(ccall((:fegpu_makematrix_copy_values, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), self.handle, nzval), self.ctx)
        return SparseMatrixCSC(m, n, dt.colptr, dt.rowval, nzval)   # shares the index arrays with the earlier matrix (they are equal)
    end
    colptr = alloc(Int64, n + 1); rowval = alloc(Int64, nnz); nzval = alloc(Float64, nnz)
    GC.@preserve colptr rowval nzval _check(ccall((:fegpu_makematrix_copy, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), self.handle, colptr, rowval, nzval), self.ctx)
    if dt !== nothing
        dt.colptr, dt.rowval, dt.have_pattern = colptr, rowval, true
    end
    return SparseMatrixCSC(m, n, colptr, rowval, nzval)   # 1-based Int64 arrays, used as they are
end

"""
    gpu_coo(assembler) -> (I, J, V)

The raw triplets of the last bilinear-form assembly in the reference's emission order (AssemblyModule.jl:261-280): what
`_rowbuffer, _colbuffer, _matbuffer` hold after an assembly with `nomatrixresult = true` (:29-33, 309-317).
"""
function gpu_coo(self::SysmatAssemblerSparseGPU)
    self._last === nothing && error("no bilinear-form assembly to export")
    mt, dt = self._last
    lo, hi, na = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fegpu_mesh_window, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), mt.handle, lo, hi, na), self.ctx)
    elmdim = length(first(mt.conn)) * size(dt.dofnums, 2)
    n = na[] * elmdim * elmdim
    I = Vector{Int64}(undef, n); J = Vector{Int64}(undef, n); V = Vector{Float64}(undef, n)
    GC.@preserve I J V _check(ccall((:fegpu_coo_copy, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
            self.handle, mt.handle, dt.handle, I, J, V), self.ctx)
    return I, J, V
end

"""
    gpu_device_pointers(assembler) -> (colptr, rowval, nzval) as `Ptr`s into device memory

For a GPU solver that consumes the CSC where it is (CUDA.jl: `unsafe_wrap(CuArray, CuPtr{Float64}(UInt(p)), nnz)`).
"""
function gpu_device_pointers(self::SysmatAssemblerSparseGPU)
    c, r, v = Ref{Ptr{Int64}}(C_NULL), Ref{Ptr{Int64}}(C_NULL), Ref{Ptr{Float64}}(C_NULL)
    _check(ccall((:fegpu_makematrix_device, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Int64}}, Ref{Ptr{Int64}}, Ref{Ptr{Float64}}), self.handle, c, r, v), self.ctx)
    return c[], r[], v[]
end

"""
    SysmatAssemblerSparseSymmGPU(z = 0.0)

Same protocol and result as `SysmatAssemblerSparseSymm` (AssemblyModule.jl:342-583): lower triangles in, `S + transpose(S)`
with the diagonal halved out (entries that sum to exactly 0.0 are not stored).  It is a `SysmatAssemblerSparseGPU` whose
library handle is switched to symmetric semantics, so every method above and below applies unchanged.
"""
function SysmatAssemblerSparseSymmGPU(z::Float64 = 0.0, nomatrixresult = false; kw...)
    a = SysmatAssemblerSparseGPU(z, nomatrixresult; kw...)
    _check(ccall((:fegpu_asm_set_symmetric, LIB), Int32, (Ptr{Cvoid}, Int32), a.handle, 1), a.ctx)
    return a
end

"""
    SysmatAssemblerSparseDiagGPU(z = 0.0), SysmatAssemblerSparseHRZLumpingSymmGPU(z = 0.0)

`SysmatAssemblerSparseDiag` (AssemblyModule.jl:599-794) and `SysmatAssemblerSparseHRZLumpingSymm` (:943-1141): the handle is
switched to diagonal (1) / HRZ (2) lumping, every method of `SysmatAssemblerSparseGPU` applies unchanged, e.g.
`M = bilform_dot(femm, SysmatAssemblerSparseHRZLumpingSymmGPU(0.0), geom, u, DataCache(rho * I(3)))`.
"""
function SysmatAssemblerSparseDiagGPU(z::Float64 = 0.0, nomatrixresult = false; mode = 1, kw...)
    a = SysmatAssemblerSparseGPU(z, nomatrixresult; kw...)
    _check(ccall((:fegpu_asm_set_lumping, LIB), Int32, (Ptr{Cvoid}, Int32), a.handle, mode), a.ctx)
    return a
end
SysmatAssemblerSparseHRZLumpingSymmGPU(z::Float64 = 0.0, nomatrixresult = false; kw...) =
    SysmatAssemblerSparseDiagGPU(z, nomatrixresult; mode = 2, kw...)

# one block A[r0:r1, c0:c1] of the device-resident matrix (Julia range indexing: stored zeros kept, rows rebased)
function _block(a::SysmatAssemblerSparseGPU, r0, r1, c0, c1)
    m, n = a._row_nalldofs, a._col_nalldofs
    (r1 < r0 || c1 < c0) && return spzeros(max(r1 - r0 + 1, 0), max(c1 - c0 + 1, 0))
    _check(ccall((:fegpu_makematrix_view, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int32), a.handle, r0, r1, c0, c1, 0), a.ctx)
    last, a._last = a._last, nothing      # a view is not the pattern: fetch all three arrays, do not record them as the pattern
    B = try
        makematrix!(a)
    finally
        a._last = last
        ccall((:fegpu_makematrix_view, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int32), a.handle, 1, m, 1, n, 0)
    end
    return B
end

"""
    gpu_matrix_blocked(a::SysmatAssemblerSparseGPU, row_nfreedofs, col_nfreedofs = row_nfreedofs)

`matrix_blocked_ff/fd/df/dd` (MatrixUtilityModule.jl:675-793) cut on the device from the assembler's resident matrix; only the
blocks cross the PCIe link.
"""
function gpu_matrix_blocked(a::SysmatAssemblerSparseGPU, rf::Int, cf::Int = rf)
    m, n = a._row_nalldofs, a._col_nalldofs
    rf <= m || error("The ff block has too many rows")
    cf <= n || error("The ff block has too many columns")
    return (ff = _block(a, 1, rf, 1, cf), fd = _block(a, 1, rf, cf + 1, n), df = _block(a, rf + 1, m, 1, cf), dd = _block(a, rf + 1, m, cf + 1, n))
end

# ---- eligibility: what may cross the C ABI (SURVEY.md 8b) -- everything else is an error, never a silent CPU loop ----------
# DataCache (DataCacheModule.jl:65-89): the struct carries a function `_fillcache!`; only the closure of the CONSTANT constructor
# (:77-89, `_fillcache_constant!`, which returns its buffer untouched) can be evaluated once on the host.  A cache built from a
# user function depends on (XYZ, tangents, feid, qpid): reading its initial buffer would assemble a wrong matrix.
function _constant_cache(cf::DataCache)
    occursin("_fillcache_constant!", string(nameof(typeof(cf._fillcache!)))) ||
        error("only a constant DataCache (DataCache(data), DataCacheModule.jl:77-89) is GPU-eligible; this one evaluates a function per quadrature point")
    return cf._cache
end

# other dimension (IntegDomainModule.jl:43-48): `otherdimensionunity` (:150-152) or the closure `otherdimensionfu` of the constant
# constructors (:73-82, :128-142), which captures the number; a user-supplied function (:96-104) is position dependent
function _otherdim(self::FEMMBase)
    f = self.integdomain.otherdimension
    f === otherdimensionunity && return 1.0
    if occursin("otherdimensionfu", string(nameof(typeof(f)))) && hasfield(typeof(f), :otherdimension)
        t = getfield(f, :otherdimension)
        t isa Number && return Float64(t)
    end
    error("only the unit or a constant other-dimension (IntegDomain(fes, rule, t::Number)) is GPU-eligible")
end

function _eligible(self::FEMMBase, geom, u, cf)
    self.mcsys.isconstant || error("only constant material coordinate systems (CSys(dim), CSys(csmat)) are GPU-eligible")
    self.integdomain.axisymmetric && error("axisymmetric integration domains are not GPU-eligible")
    _otherdim(self)
    _constant_cache(cf)
    eltype(geom.values) == Float64 || error("geom must be Float64")
    eltype(u.dofnums) == Int64 || error("dofnums must be Int64")
    return nothing
end

# ---- device twins of (fes, geom, u), shared per device ----------------------------------------------------------------------
function _mesh_twin(st::DeviceState, self::FEMMBase, geom)
    fes = finite_elements(self)
    xyz = geom.values                                      # nnodes x sdim, column-major already
    k = findfirst(t -> t.conn === fes.conn, st.meshes)
    if k !== nothing
        t = st.meshes[k]
        deleteat!(st.meshes, k); push!(st.meshes, t)       # most recently used last
        return t, false
    end
    conn = Matrix{Int64}(undef, nodesperelem(fes), count(fes))   # [nelem][nne] row-major == nne x nelem column-major
    for (i, c) in enumerate(fes.conn), j in eachindex(c)
        conn[j, i] = c[j]
    end
    mh = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve conn xyz _check(ccall((:fegpu_mesh_upload, LIB), Int32,
            (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Ref{Ptr{Cvoid}}),
            st.ctx, _etype(fes), count(fes), conn, size(xyz, 1), size(xyz, 2), xyz, mh), st.ctx)
    t = MeshTwin(fes.conn, mh[], nothing, DofTwin[], nothing)
    push!(st.meshes, t)
    while length(st.meshes) > MAX_MESHES
        _destroy!(popfirst!(st.meshes))
    end
    return t, true
end

function _device(self::FEMMBase, a, geom, u)   # a: SysmatAssemblerSparseGPU or SysvecAssemblerGPU
    st = _device_state(a.device)
    t, fresh = _mesh_twin(st, self, geom)
    mh = t.handle
    # quadrature tables: uploaded whenever they differ from what the device holds (two FEMMs on one FESet may use different rules,
    # e.g. stiffness with GaussRule(3,2) and mass with GaussRule(3,3))
    npts, Ns, gradNparams, w, pc = integrationdata(self.integdomain)
    N = reduce(hcat, [vec(Ns[j]) for j in 1:npts])                       # nne x npts  == [npts][nne]
    dN = reduce(hcat, [vec(gradNparams[j]) for j in 1:npts])            # (nne*mdim) x npts == [npts][mdim][nne]
    ww = Vector{Float64}(vec(w))
    if t.rule === nothing || t.rule != (npts, N, dN, ww)
        GC.@preserve N dN ww _check(ccall((:fegpu_rule_set, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), mh, npts, N, dN, ww), a.ctx)
        t.rule = (npts, N, dN, ww)
    end
    _check(ccall((:fegpu_otherdimension_set, LIB), Int32, (Ptr{Cvoid}, Float64), mh, _otherdim(self)), a.ctx)
    if self.mcsys.isidentity
        _check(ccall((:fegpu_csys_set, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), mh, C_NULL), a.ctx)
    else
        rm = Matrix{Float64}(csmat(self.mcsys))       # constant: the buffer already holds the matrix (CSysModule.jl:133-144)
        GC.@preserve rm _check(ccall((:fegpu_csys_set, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), mh, rm), a.ctx)
    end
    # row-block partition of this assembler (nothing = the whole matrix); compared by content
    own, rank = a isa SysmatAssemblerSparseGPU ? (a.node_owner, a.rank) : (nothing, Int32(0))
    if own === nothing
        if t.owner !== nothing
            _check(ccall((:fegpu_partition_set, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32), mh, C_NULL, 0), a.ctx)
            t.owner = nothing
        end
    elseif t.owner === nothing || t.owner[2] != rank || t.owner[1] != own
        length(own) == size(geom.values, 1) || error("node_owner must have one entry per node")
        GC.@preserve own _check(ccall((:fegpu_partition_set, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32), mh, own, rank), a.ctx)
        t.owner = (copy(own), rank)
    end
    if !fresh   # the upload of a new twin carried the coordinates; a partitioned mesh ships only its node window
        xyz = geom.values
        if t.owner === nothing
            GC.@preserve xyz _check(ccall((:fegpu_geom_update, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), mh, xyz), a.ctx)
        else
            GC.@preserve xyz _check(ccall((:fegpu_geom_update_window, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), mh, xyz), a.ctx)
        end
    end
    k = findfirst(d -> d.nall == nalldofs(u) && d.dofnums == u.dofnums, t.dofs)
    if k === nothing
        d = Ref{Ptr{Cvoid}}(C_NULL)
        dn = u.dofnums
        GC.@preserve dn _check(ccall((:fegpu_dofmap_upload, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int64}, Int64, Int64, Ref{Ptr{Cvoid}}),
                a.ctx, mh, ndofs(u), dn, nalldofs(u), nalldofs(u), d), a.ctx)
        push!(t.dofs, DofTwin(copy(dn), nalldofs(u), d[], Int64[], Int64[], false))
        if length(t.dofs) > 4
            old = popfirst!(t.dofs)
            ccall((:fegpu_dofmap_destroy, LIB), Int32, (Ptr{Cvoid},), old.handle)
        end
        k = length(t.dofs)
    end
    dt = t.dofs[k]
    if a isa SysmatAssemblerSparseGPU
        a._row_nalldofs = a._col_nalldofs = nalldofs(u)
        a._generic = false
        a._last = (t, dt)
    end
    return mh, dt.handle
end

"""
    gpu_release_cache(assembler)

Hand the device blocks that the symbolic phase keeps for re-use (pattern arrays, temporaries) back to the CUDA driver, e.g. before
another library needs the memory.  Results and cached patterns stay valid.
"""
gpu_release_cache(a) = (_check(ccall((:fegpu_cache_release, LIB), Int32, (Ptr{Cvoid},), a.ctx), a.ctx); a)

# A form call only queues its device work when the result is fetched right after (makematrix! synchronises): the transport
# then ships the pattern's arrays, and the host threads rebuild rowval, while the integration and the numeric phase still run.
function _queued(f, a)
    on = !(a isa SysmatAssemblerSparseGPU) || !a._nomatrixresult
    on && ccall((:fegpu_set_async, LIB), Int32, (Ptr{Cvoid}, Int32), a.ctx, 1)
    try
        return f()
    finally
        on && ccall((:fegpu_set_async, LIB), Int32, (Ptr{Cvoid}, Int32), a.ctx, 0)
    end
end

# ---- the three forms: more specific methods than the generic drivers (FEMMBaseModule.jl:1335, 1462, 1774) ---------------
function bilform_diffusion(self::FEMMBase, assembler::SysmatAssemblerSparseGPU, geom::NodalField{FT}, u::NodalField{T},
    cf::DC) where {FT,T,DC<:DataCache}
    _eligible(self, geom, u, cf)
    mh, dh = _device(self, assembler, geom, u)
    kappa = _constant_cache(cf)
    kind = isempty(size(kappa)) ? 0 : 1
    k = kind == 0 ? Float64[kappa] : Matrix{Float64}(kappa)
    return _queued(assembler) do
        GC.@preserve k _check(ccall((:fegpu_bilform_diffusion, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Cvoid}),
                mh, dh, kind, k, assembler.handle), assembler.ctx)
        makematrix!(assembler)
    end
end

function bilform_lin_elastic(self::FEMMBase, assembler::SysmatAssemblerSparseGPU, geom::NodalField{FT}, u::NodalField{T},
    mr::Type{DeforModelRed3D}, cf::DC) where {FT,T,DC<:DataCache}
    _eligible(self, geom, u, cf)
    mh, dh = _device(self, assembler, geom, u)
    C = Matrix{Float64}(_constant_cache(cf))
    size(C) == (6, 6) || error("Wrong dimensions")
    return _queued(assembler) do
        GC.@preserve C _check(ccall((:fegpu_bilform_lin_elastic, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}),
                mh, dh, C, assembler.handle), assembler.ctx)
        makematrix!(assembler)
    end
end

function bilform_dot(self::FEMMBase, assembler::SysmatAssemblerSparseGPU, geom::NodalField{FT}, u::NodalField{T}, cf::DC;
    m = 3) where {FT,T,DC<:DataCache}
    _eligible(self, geom, u, cf)
    mh, dh = _device(self, assembler, geom, u)
    c = Matrix{Float64}(_constant_cache(cf))          # densifies LinearAlgebra.I(ndofs) (a Diagonal{Bool}), see innerproduct :1388-1401
    size(c) == (ndofs(u), ndofs(u)) || error("Wrong size of matrix")
    return _queued(assembler) do
        GC.@preserve c _check(ccall((:fegpu_bilform_dot, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ptr{Cvoid}),
                mh, dh, c, m, _otherdim(self), assembler.handle), assembler.ctx)
        makematrix!(assembler)
    end
end

# ---- SURVEY.md 8(f) rank 3: sibling forms on the same per-element pipeline ----------------------------------------------
# bilform_convection (FEMMBaseModule.jl:1583-1625): `u` is the nodal convective velocity field, `Q` numbers the dofs
function bilform_convection(self::FEMMBase, assembler::SysmatAssemblerSparseGPU, geom::NodalField{FT}, u::NodalField{T},
    Q::NodalField{QT}, rhof::DC) where {FT,T,QT,DC<:DataCache}
    _eligible(self, geom, Q, rhof)
    mh, dh = _device(self, assembler, geom, Q)
    uv = Matrix{Float64}(u.values)           # nnodes x sdim, column-major
    return _queued(assembler) do
        GC.@preserve uv _check(ccall((:fegpu_bilform_convection, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Float64, Ptr{Cvoid}),
                mh, dh, uv, Float64(_constant_cache(rhof)), assembler.handle), assembler.ctx)
        makematrix!(assembler)
    end
end

# bilform_div_grad (FEMMBaseModule.jl:1672-1713)
function bilform_div_grad(self::FEMMBase, assembler::SysmatAssemblerSparseGPU, geom::NodalField{FT}, u::NodalField{T},
    viscf::DC) where {FT,T,DC<:DataCache}
    _eligible(self, geom, u, viscf)
    mh, dh = _device(self, assembler, geom, u)
    return _queued(assembler) do
        _check(ccall((:fegpu_bilform_div_grad, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}),
                mh, dh, Float64(_constant_cache(viscf)), assembler.handle), assembler.ctx)
        makematrix!(assembler)
    end
end

# bilform_masslike (FEMMBaseModule.jl:1865-1912): rectangular (count(fes)*ndn) x nalldofs(phi), rows numbered by element
function bilform_masslike(self::FEMMBase, assembler::SysmatAssemblerSparseGPU, geom::NodalField{FT}, phi::NodalField{T}, cf::DC;
    m = 3) where {FT,T,DC<:DataCache}
    _eligible(self, geom, phi, cf)
    mh, dh = _device(self, assembler, geom, phi)
    c = Matrix{Float64}(reshape(collect(_constant_cache(cf)), ndofs(phi), ndofs(phi)))
    GC.@preserve c _check(ccall((:fegpu_bilform_masslike, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ptr{Cvoid}),
            mh, dh, c, m, _otherdim(self), assembler.handle), assembler.ctx)
    assembler._row_nalldofs, assembler._col_nalldofs = count(finite_elements(self)) * ndofs(phi), nalldofs(phi)
    assembler._last = nothing   # not the square pattern of the dof map
    return makematrix!(assembler)
end

# ---- SysmatAssemblerFFBlock wrapping a GPU assembler (AssemblyModule.jl:1149-1231) ------------------------------------------
# Without these methods the generic Julia driver would run the element loop on the CPU and feed the GPU assembler one element
# matrix per ccall.  Here the inner assembler runs the form on the device and leaves the matrix there; the free-free block
# A[1:row_nfreedofs, 1:col_nfreedofs] (matrix_blocked_ff, MatrixUtilityModule.jl:675-688) is cut on the device and only it comes back.
const FFBlockGPU = SysmatAssemblerFFBlock{<:SysmatAssemblerSparseGPU}

function _ffblock(form, assembler::FFBlockGPU)
    inner = assembler._a
    keep = inner._nomatrixresult
    inner._nomatrixresult = true            # run the form, keep the result on the device
    try
        form(inner)
    finally
        inner._nomatrixresult = keep
    end
    assembler._row_nfreedofs <= inner._row_nalldofs || error("The ff block has too many rows")
    assembler._col_nfreedofs <= inner._col_nalldofs || error("The ff block has too many columns")
    return _block(inner, 1, assembler._row_nfreedofs, 1, assembler._col_nfreedofs)
end

bilform_diffusion(self::FEMMBase, assembler::FFBlockGPU, geom::NodalField{FT}, u::NodalField{T}, cf::DC) where {FT,T,DC<:DataCache} =
    _ffblock(a -> bilform_diffusion(self, a, geom, u, cf), assembler)
bilform_lin_elastic(self::FEMMBase, assembler::FFBlockGPU, geom::NodalField{FT}, u::NodalField{T}, mr::Type{DeforModelRed3D},
    cf::DC) where {FT,T,DC<:DataCache} = _ffblock(a -> bilform_lin_elastic(self, a, geom, u, mr, cf), assembler)
bilform_dot(self::FEMMBase, assembler::FFBlockGPU, geom::NodalField{FT}, u::NodalField{T}, cf::DC; m = 3) where {FT,T,DC<:DataCache} =
    _ffblock(a -> bilform_dot(self, a, geom, u, cf; m), assembler)
bilform_convection(self::FEMMBase, assembler::FFBlockGPU, geom::NodalField{FT}, u::NodalField{T}, Q::NodalField{QT},
    rhof::DC) where {FT,T,QT,DC<:DataCache} = _ffblock(a -> bilform_convection(self, a, geom, u, Q, rhof), assembler)
bilform_div_grad(self::FEMMBase, assembler::FFBlockGPU, geom::NodalField{FT}, u::NodalField{T}, viscf::DC) where {FT,T,DC<:DataCache} =
    _ffblock(a -> bilform_div_grad(self, a, geom, u, viscf), assembler)

"""
    SysvecAssemblerGPU(like::SysmatAssemblerSparseGPU)

Same protocol as `SysvecAssembler` (AssemblyModule.jl:853-917).  It lives on the device of the matrix assembler it is built from
and therefore shares that device's twins (mesh, dof maps, node -> element adjacency).
`distribloads(femm, SysvecAssemblerGPU(a), geom, P, fi, m)` works through the reference's own forwarding method
(FEMMBaseModule.jl:1277-1286) and the `linform_dot` method below.
"""
mutable struct SysvecAssemblerGPU{T} <: AbstractSysvecAssembler
    ctx::Ptr{Cvoid}
    handle::Ptr{Cvoid}
    device::Int
    _row_nalldofs::Int
end

function SysvecAssemblerGPU(like::SysmatAssemblerSparseGPU)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fegpu_asm_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), like.ctx, h), like.ctx)
    a = SysvecAssemblerGPU{Float64}(like.ctx, h[], like.device, 1)
    finalizer(x -> ccall((:fegpu_asm_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), a)
    return a
end

function startassembly!(self::SysvecAssemblerGPU, row_nalldofs::IT) where {IT<:Integer}
    _check(ccall((:fegpu_vec_startassembly, LIB), Int32, (Ptr{Cvoid}, Int64), self.handle, row_nalldofs), self.ctx)
    self._row_nalldofs = row_nalldofs
    return self
end

function assemble!(self::SysvecAssemblerGPU, vec::MV, dofnums::IV) where {MV,IV}
    v = Vector{Float64}(vec); d = Vector{Int64}(dofnums)
    GC.@preserve v d _check(ccall((:fegpu_vec_assemble, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64),
            self.handle, v, d, length(d)), self.ctx)
end

function _fetchvector(self::SysvecAssemblerGPU)
    n = Ref{Int64}(0)
    _check(ccall((:fegpu_makevector_size, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), self.handle, n), self.ctx)
    F = Vector{Float64}(undef, n[])
    GC.@preserve F _check(ccall((:fegpu_makevector_copy, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), self.handle, F), self.ctx)
    return F
end

function makevector!(self::SysvecAssemblerGPU)
    _check(ccall((:fegpu_makevector, LIB), Int32, (Ptr{Cvoid},), self.handle), self.ctx)
    return _fetchvector(self)
end

# linform_dot (FEMMBaseModule.jl:1207-1244): constant DataCache only
function linform_dot(self::FEMMBase, assembler::SysvecAssemblerGPU, geom::NodalField{FT}, P::NodalField{T}, f::DC,
    m) where {FT<:Number,T,DC<:DataCache}
    _eligible(self, geom, P, f)
    mh, dh = _device(self, assembler, geom, P)
    force = Vector{Float64}(vec(collect(_constant_cache(f))))
    length(force) == ndofs(P) || error("the load needs one component per degree of freedom of a node")
    GC.@preserve force _check(ccall((:fegpu_linform_dot, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ptr{Cvoid}),
            mh, dh, force, m, _otherdim(self), assembler.handle), assembler.ctx)
    assembler._row_nalldofs = nalldofs(P)
    return _fetchvector(assembler)
end

end # module
