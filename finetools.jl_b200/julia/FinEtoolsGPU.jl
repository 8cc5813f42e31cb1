# FinEtoolsGPU.jl -- drop-in GPU assembler for FinEtools.jl (v8.2.x) backed by libfinegpu.so (include/fegpu.h).
#
# NOT EXECUTED in the build environment (no Julia on the image).  It is the reference-side binding a maintainer adds:
# every `ccall` below has a line-for-line twin in finetools.jl_b200/_lib.py + assembly.py + femm.py, which IS what the
# test-suite and bench.py run.  parity.jl (next to this file) compares makematrix! outputs of the two assemblers.
#
#   using FinEtools, FinEtoolsGPU
#   K = bilform_diffusion(femm, SysmatAssemblerSparseGPU(0.0), geom, u, DataCache(kappa))   # unchanged call shape
#
module FinEtoolsGPU

using FinEtools
using SparseArrays
import FinEtools.AssemblyModule: AbstractSysmatAssembler, AbstractSysvecAssembler, startassembly!, assemble!, makematrix!,
    makevector!, eltype, expectedntriples
import FinEtools.FEMMBaseModule: bilform_diffusion, bilform_lin_elastic, bilform_dot, bilform_convection, bilform_div_grad,
    bilform_masslike, linform_dot, FEMMBase, finite_elements
using FinEtools.IntegDomainModule: integrationdata, otherdimensionunity
using FinEtools.DeforModelRedModule: DeforModelRed3D
using FinEtools.CSysModule: csmat

export SysmatAssemblerSparseGPU, SysmatAssemblerSparseSymmGPU, SysmatAssemblerSparseDiagGPU, SysmatAssemblerSparseHRZLumpingSymmGPU,
    SysvecAssemblerGPU, gpu_matrix_blocked, gpu_release_cache

const LIB = get(ENV, "FEGPU_LIB", joinpath(@__DIR__, "..", "libfinegpu.so"))

_etype(::FESetT3) = 1; _etype(::FESetQ4) = 2; _etype(::FESetT4) = 3; _etype(::FESetT10) = 4
_etype(::FESetH8) = 5; _etype(::FESetH20) = 6; _etype(::FESetH27) = 7
_etype(fes) = error("Element type $(typeof(fes)) is not GPU-eligible")

function _check(status::Int32, ctx::Ptr{Cvoid} = C_NULL)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:fegpu_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    error(msg)   # same strings as AssemblyModule.jl:265-273 ("Row degree of freedom > size", ...)
end

"""
    SysmatAssemblerSparseGPU{T} <: AbstractSysmatAssembler

Same protocol as `SysmatAssemblerSparse` (AssemblyModule.jl:88-329); the element loop of the three bilinear forms and the
COO -> CSC conversion run on the GPU.
"""
mutable struct SysmatAssemblerSparseGPU{T} <: AbstractSysmatAssembler
    ctx::Ptr{Cvoid}
    handle::Ptr{Cvoid}
    meshes::IdDict{Any,Any}          # fes => (mesh handle, Dict(dofnums copy => dofmap handle))
    _row_nalldofs::Int
    _col_nalldofs::Int
    _nomatrixresult::Bool
    _force_init::Bool
    _generic::Bool
end

function SysmatAssemblerSparseGPU(z::Float64 = 0.0, nomatrixresult = false; device = 0)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fegpu_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), ctx, device))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fegpu_asm_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx[], h), ctx[])
    a = SysmatAssemblerSparseGPU{Float64}(ctx[], h[], IdDict(), 0, 0, nomatrixresult, false, false)
    finalizer(a) do x
        for (_, (m, dms)) in x.meshes
            foreach(d -> ccall((:fegpu_dofmap_destroy, LIB), Int32, (Ptr{Cvoid},), d), values(dms))
            ccall((:fegpu_mesh_destroy, LIB), Int32, (Ptr{Cvoid},), m)
        end
        ccall((:fegpu_asm_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle)
        ccall((:fegpu_destroy, LIB), Int32, (Ptr{Cvoid},), x.ctx)
    end
    return a
end

eltype(::SysmatAssemblerSparseGPU{T}) where {T} = T

# ---- generic protocol (any caller of startassembly!/assemble!/makematrix! keeps working) ------------------------------
function startassembly!(self::SysmatAssemblerSparseGPU, elem_mat_nrows::IT, elem_mat_ncols::IT, n_elem_mats::IT,
    row_nalldofs::IT, col_nalldofs::IT; force_init = false) where {IT<:Integer}
    _check(ccall((:fegpu_startassembly, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64),
            self.handle, elem_mat_nrows, elem_mat_ncols, n_elem_mats, row_nalldofs, col_nalldofs), self.ctx)
    self._row_nalldofs, self._col_nalldofs, self._generic = row_nalldofs, col_nalldofs, true
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

function makematrix!(self::SysmatAssemblerSparseGPU)
    if self._nomatrixresult
        return spzeros(self._row_nalldofs, self._col_nalldofs)
    end
    if self._generic
        _check(ccall((:fegpu_makematrix, LIB), Int32, (Ptr{Cvoid},), self.handle), self.ctx)
        self._generic = false
    end
    m, n, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fegpu_makematrix_sizes, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), self.handle, m, n, nnz), self.ctx)
    colptr = Vector{Int64}(undef, n[] + 1); rowval = Vector{Int64}(undef, nnz[]); nzval = Vector{Float64}(undef, nnz[])
    GC.@preserve colptr rowval nzval _check(ccall((:fegpu_makematrix_copy, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), self.handle, colptr, rowval, nzval), self.ctx)
    return SparseMatrixCSC(m[], n[], colptr, rowval, nzval)   # 1-based Int64 arrays, used as they are
end

"""
    SysmatAssemblerSparseSymmGPU(z = 0.0)

Same protocol and result as `SysmatAssemblerSparseSymm` (AssemblyModule.jl:342-583): lower triangles in, `S + transpose(S)`
with the diagonal halved out (entries that sum to exactly 0.0 are not stored).  It is a `SysmatAssemblerSparseGPU` whose
library handle is switched to symmetric semantics, so every method above and below applies unchanged.
"""
function SysmatAssemblerSparseSymmGPU(z::Float64 = 0.0, nomatrixresult = false; device = 0)
    a = SysmatAssemblerSparseGPU(z, nomatrixresult; device)
    _check(ccall((:fegpu_asm_set_symmetric, LIB), Int32, (Ptr{Cvoid}, Int32), a.handle, 1), a.ctx)
    return a
end

"""
    SysmatAssemblerSparseDiagGPU(z = 0.0), SysmatAssemblerSparseHRZLumpingSymmGPU(z = 0.0)

`SysmatAssemblerSparseDiag` (AssemblyModule.jl:599-794) and `SysmatAssemblerSparseHRZLumpingSymm` (:943-1141): the handle is
switched to diagonal (1) / HRZ (2) lumping, every method of `SysmatAssemblerSparseGPU` applies unchanged, e.g.
`M = bilform_dot(femm, SysmatAssemblerSparseHRZLumpingSymmGPU(0.0), geom, u, DataCache(rho * I(3)))`.
"""
function SysmatAssemblerSparseDiagGPU(z::Float64 = 0.0, nomatrixresult = false; device = 0, mode = 1)
    a = SysmatAssemblerSparseGPU(z, nomatrixresult; device)
    _check(ccall((:fegpu_asm_set_lumping, LIB), Int32, (Ptr{Cvoid}, Int32), a.handle, mode), a.ctx)
    return a
end
SysmatAssemblerSparseHRZLumpingSymmGPU(z::Float64 = 0.0, nomatrixresult = false; device = 0) =
    SysmatAssemblerSparseDiagGPU(z, nomatrixresult; device, mode = 2)

"""
    gpu_matrix_blocked(a::SysmatAssemblerSparseGPU, row_nfreedofs, col_nfreedofs = row_nfreedofs)

`matrix_blocked_ff/fd/df/dd` (MatrixUtilityModule.jl:675-793) cut on the device from the assembler's resident matrix; only the
blocks cross the PCIe link.  `SysmatAssemblerFFBlock(SysmatAssemblerSparseGPU(0.0), nf, nf)` works as it is (the wrapper
delegates to the inner assembler, AssemblyModule.jl:1149-1231) but copies the full matrix first; this does not.
"""
function gpu_matrix_blocked(a::SysmatAssemblerSparseGPU, rf::Int, cf::Int = rf)
    m, n = a._row_nalldofs, a._col_nalldofs
    function block(r0, r1, c0, c1)
        (r1 < r0 || c1 < c0) && return spzeros(max(r1 - r0 + 1, 0), max(c1 - c0 + 1, 0))
        _check(ccall((:fegpu_makematrix_view, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int32), a.handle, r0, r1, c0, c1, 0), a.ctx)
        B = makematrix!(a)
        _check(ccall((:fegpu_makematrix_view, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int32), a.handle, 1, m, 1, n, 0), a.ctx)
        return B
    end
    return (ff = block(1, rf, 1, cf), fd = block(1, rf, cf + 1, n), df = block(rf + 1, m, 1, cf), dd = block(rf + 1, m, cf + 1, n))
end

# ---- device twins of (fes, geom, u) -------------------------------------------------------------------------------------
function _eligible(self::FEMMBase, geom, u, cf)
    self.mcsys.isconstant || error("only constant material coordinate systems (CSys(dim), CSys(csmat)) are GPU-eligible")
    self.integdomain.axisymmetric && error("axisymmetric integration domains are not GPU-eligible")
    # other dimension: unity, or the constant closure of IntegDomain(fes, rule, t) (IntegDomainModule.jl:73-82); _otherdim evaluates it
    # DataCache: only the constant constructor (DataCacheModule.jl:77-89) may cross the boundary
    eltype(geom.values) == Float64 || error("geom must be Float64")
    eltype(u.dofnums) == Int64 || error("dofnums must be Int64")
    return nothing
end

# constant other-dimension: the closure ignores its arguments (IntegDomainModule.jl:76-78), so one evaluation gives the constant
_otherdim(self::FEMMBase) = Float64(self.integdomain.otherdimension(zeros(1, 3), finite_elements(self).conn[1], zeros(1)))

function _device(self::FEMMBase, a, geom, u)   # a: SysmatAssemblerSparseGPU or SysvecAssemblerGPU (same device-twin cache)
    fes = finite_elements(self)
    xyz = geom.values                                      # nnodes x sdim, column-major already
    entry = get(a.meshes, fes, nothing)
    if entry === nothing
        conn = Matrix{Int64}(undef, nodesperelem(fes), count(fes))   # [nelem][nne] row-major == nne x nelem column-major
        for (i, c) in enumerate(fes.conn), k in eachindex(c)
            conn[k, i] = c[k]
        end
        mh = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve conn xyz _check(ccall((:fegpu_mesh_upload, LIB), Int32,
                (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                a.ctx, _etype(fes), count(fes), conn, size(xyz, 1), size(xyz, 2), xyz, mh), a.ctx)
        npts, Ns, gradNparams, w, pc = integrationdata(self.integdomain)
        N = reduce(hcat, [vec(Ns[j]) for j in 1:npts])                       # nne x npts  == [npts][nne]
        dN = reduce(hcat, [vec(gradNparams[j]) for j in 1:npts])            # (nne*mdim) x npts == [npts][mdim][nne]
        ww = Vector{Float64}(vec(w))
        GC.@preserve N dN ww _check(ccall((:fegpu_rule_set, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), mh[], npts, N, dN, ww), a.ctx)
        entry = (mh[], Dict{Matrix{Int64},Ptr{Cvoid}}())
        a.meshes[fes] = entry
    else
        GC.@preserve xyz _check(ccall((:fegpu_geom_update, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), entry[1], xyz), a.ctx)
    end
    mh, dms = entry
    _check(ccall((:fegpu_otherdimension_set, LIB), Int32, (Ptr{Cvoid}, Float64), mh, _otherdim(self)), a.ctx)
    if self.mcsys.isidentity
        _check(ccall((:fegpu_csys_set, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), mh, C_NULL), a.ctx)
    else
        rm = Matrix{Float64}(csmat(self.mcsys))       # constant: the buffer already holds the matrix (CSysModule.jl:133-144)
        GC.@preserve rm _check(ccall((:fegpu_csys_set, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), mh, rm), a.ctx)
    end
    dh = get(dms, u.dofnums, C_NULL)
    if dh == C_NULL
        d = Ref{Ptr{Cvoid}}(C_NULL)
        dn = u.dofnums
        GC.@preserve dn _check(ccall((:fegpu_dofmap_upload, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int64}, Int64, Int64, Ref{Ptr{Cvoid}}),
                a.ctx, mh, ndofs(u), dn, nalldofs(u), nalldofs(u), d), a.ctx)
        dms[copy(dn)] = d[]
        dh = d[]
    end
    if a isa SysmatAssemblerSparseGPU
        a._row_nalldofs = a._col_nalldofs = nalldofs(u)
        a._generic = false
    end
    return mh, dh
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
    kappa = cf._cache
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
    C = Matrix{Float64}(cf._cache)
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
    c = Matrix{Float64}(cf._cache)          # densifies LinearAlgebra.I(ndofs) (a Diagonal{Bool}), see innerproduct :1388-1401
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
                mh, dh, uv, Float64(rhof._cache), assembler.handle), assembler.ctx)
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
                mh, dh, Float64(viscf._cache), assembler.handle), assembler.ctx)
        makematrix!(assembler)
    end
end

# bilform_masslike (FEMMBaseModule.jl:1865-1912): rectangular (count(fes)*ndn) x nalldofs(phi), rows numbered by element
function bilform_masslike(self::FEMMBase, assembler::SysmatAssemblerSparseGPU, geom::NodalField{FT}, phi::NodalField{T}, cf::DC;
    m = 3) where {FT,T,DC<:DataCache}
    _eligible(self, geom, phi, cf)
    mh, dh = _device(self, assembler, geom, phi)
    c = Matrix{Float64}(reshape(collect(cf._cache), ndofs(phi), ndofs(phi)))
    GC.@preserve c _check(ccall((:fegpu_bilform_masslike, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ptr{Cvoid}),
            mh, dh, c, m, _otherdim(self), assembler.handle), assembler.ctx)
    assembler._row_nalldofs, assembler._col_nalldofs = count(finite_elements(self)) * ndofs(phi), nalldofs(phi)
    return makematrix!(assembler)
end

"""
    SysvecAssemblerGPU(like::SysmatAssemblerSparseGPU)

Same protocol as `SysvecAssembler` (AssemblyModule.jl:853-917).  It shares the context and the device twins (mesh, dof maps,
node -> element adjacency) of the matrix assembler it is built from.  `distribloads(femm, SysvecAssemblerGPU(a), geom, P, fi, m)`
works through the reference's own forwarding method (FEMMBaseModule.jl:1277-1286) and the `linform_dot` method below.
"""
mutable struct SysvecAssemblerGPU{T} <: AbstractSysvecAssembler
    ctx::Ptr{Cvoid}
    handle::Ptr{Cvoid}
    meshes::IdDict{Any,Any}
    _row_nalldofs::Int
end

function SysvecAssemblerGPU(like::SysmatAssemblerSparseGPU)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fegpu_asm_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), like.ctx, h), like.ctx)
    a = SysvecAssemblerGPU{Float64}(like.ctx, h[], like.meshes, 1)
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
    force = Vector{Float64}(vec(collect(f._cache)))
    length(force) == ndofs(P) || error("the load needs one component per degree of freedom of a node")
    GC.@preserve force _check(ccall((:fegpu_linform_dot, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ptr{Cvoid}),
            mh, dh, force, m, _otherdim(self), assembler.handle), assembler.ctx)
    assembler._row_nalldofs = nalldofs(P)
    return _fetchvector(assembler)
end

end # module
