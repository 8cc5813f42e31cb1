# parity.jl -- NOT EXECUTED in the build environment (no Julia there).  Run with Julia >= 1.12 on a machine with a B200:
#     julia --project=. parity.jl
# Compares the reference's own makematrix! with the GPU assembler on BASELINE.json's configurations and times the serial CPU
# path.  Pass criterion (north_star): colptr / rowval identical, max|nzval - ref| <= 1e-12 * max|ref|.
using FinEtools, SparseArrays, LinearAlgebra
include(joinpath(@__DIR__, "FinEtoolsGPU.jl"))
using .FinEtoolsGPU

function compare(name, Kref, Kgpu)
    ok_pattern = (Kref.colptr == Kgpu.colptr) && (Kref.rowval == Kgpu.rowval)
    err = maximum(abs.(Kref.nzval .- Kgpu.nzval)) / maximum(abs.(Kref.nzval))
    println(name, ": pattern identical = ", ok_pattern, ", nzval rel. error = ", err, err <= 1e-12 ? "  PASS" : "  FAIL")
end

function run(n)
    kappa = [1.5 0.2 0.1; 0.2 2.5 0.3; 0.1 0.3 3.5]
    fens, fes = H8block(12.0, 1.1, 0.32, n, n, n)
    geom = NodalField(fens.xyz)
    psi = NodalField(zeros(count(fens), 1)); numberdofs!(psi)
    femm = FEMMBase(IntegDomain(fes, GaussRule(3, 2)))
    t = @elapsed Kref = bilform_diffusion(femm, SysmatAssemblerSparse(0.0), geom, psi, DataCache(kappa))
    println("reference serial bilform_diffusion: ", count(fes) / t, " elements/s")
    Kgpu = bilform_diffusion(femm, SysmatAssemblerSparseGPU(0.0), geom, psi, DataCache(kappa))
    compare("H8 diffusion $(n)^3", Kref, Kgpu)

    fens, fes = H8block(1.0, 1.0, 1.0, n, n, n)
    geom = NodalField(fens.xyz)
    u = NodalField(zeros(count(fens), 3)); numberdofs!(u)
    femm = FEMMBase(IntegDomain(fes, GaussRule(3, 2)))
    E, nu = 1.0, 0.3
    lam, mu = E * nu / ((1 + nu) * (1 - 2nu)), E / (2 * (1 + nu))
    C = [lam+2mu lam lam 0 0 0; lam lam+2mu lam 0 0 0; lam lam lam+2mu 0 0 0; 0 0 0 mu 0 0; 0 0 0 0 mu 0; 0 0 0 0 0 mu]
    t = @elapsed Kref = bilform_lin_elastic(femm, SysmatAssemblerSparse(0.0), geom, u, DeforModelRed3D, DataCache(C))
    println("reference serial bilform_lin_elastic: ", count(fes) / t, " elements/s")
    Kgpu = bilform_lin_elastic(femm, SysmatAssemblerSparseGPU(0.0), geom, u, DeforModelRed3D, DataCache(C))
    compare("H8 lin_elastic $(n)^3", Kref, Kgpu)

    fens, fes = T10block(1.0, 1.0, 1.0, n, n, n)
    geom = NodalField(fens.xyz)
    psi = NodalField(zeros(count(fens), 1)); numberdofs!(psi)
    femm = FEMMBase(IntegDomain(fes, TetRule(4)))
    Kref = bilform_dot(femm, SysmatAssemblerSparse(0.0), geom, psi, DataCache(LinearAlgebra.I(1)))
    Kgpu = bilform_dot(femm, SysmatAssemblerSparseGPU(0.0), geom, psi, DataCache(LinearAlgebra.I(1)))
    compare("T10 mass $(n)^3", Kref, Kgpu)
end

run(length(ARGS) > 0 ? parse(Int, ARGS[1]) : 20)
