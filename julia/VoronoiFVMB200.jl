# VoronoiFVMB200.jl -- Julia shim that routes VoronoiFVM.jl's Newton hot path to libvfvmb200.so (include/vfvm_b200.h).
#
# How the unchanged package API reaches the library
#   * Physics callbacks are *registered physics objects* (callable structs below).  Each one is an ordinary VoronoiFVM callback
#     -- `(f,u,edge,data)` etc., so the same System also runs on the reference's CPU path -- and additionally knows its device id
#     and parameter block.  Anything else in `Physics(...)` raises `UnregisteredPhysicsError` when a device state is created
#     (no silent CPU fallback).
#   * `SystemState(B200(), system)` builds a `VoronoiFVM.SystemState` whose `matrix` field is a `B200Matrix` (the type parameter
#     `TMatrix` of src/vfvm_state.jl:16-19).  The reference's own `solve!(state; ...)` (src/vfvm_solver.jl:551-617) and
#     `solve_transient!` (:267-537) then run unchanged and reach, by multiple dispatch on that matrix type,
#         solve_step!(state, solution, oldsol, control, time, tstep, embedparam, params, istep_factorize)   src/vfvm_solver.jl:13-23
#         eval_and_assemble(system, U, UOld, F, matrix, generic_matrix, dudp, time, tstep, λ, data, params; edge_cutoff)
#                                                                                                       src/vfvm_assembly.jl:520-534
#         _solve_linear!(u, state, nlhistory, control, method_linear, A, b, reuse_precs)                   src/vfvm_linsolve.jl:6
#     as defined here.  `solve_step!` keeps solution / residual / update resident on the device for the whole Newton iteration
#     and copies the solution back once.
#   * `solve(system, B200(); kwargs...)` = `solve!(SystemState(B200(), system); kwargs...)`.
#
# Julia is not installed in the build image of this repository, so this file is checked structurally by
# tests/test_julia_shim.py: every `ccall` names an exported symbol of the library and passes the number and kind of arguments the
# header declares.
module VoronoiFVMB200

using VoronoiFVM
using ExtendableGrids
using LinearAlgebra
using SparseArrays
import CommonSolve
import LinearSolve

import VoronoiFVM: eval_and_assemble, _solve_linear!, solve_step!, SystemState, NewtonSolverHistory, SolverControl,
    ConvergenceError, AssemblyError, LinearSolverError, num_species, unknowns, dofs, fbernoulli_pm,
    boundary_dirichlet!, boundary_neumann!, boundary_robin!, ramp

export B200, B200Matrix, UnregisteredPhysicsError
export LinearDiffusion, PowerDiffusion, CrossDiffusion2, UnipolarSGFlux, SedanFlux, BipolarSGFlux, MixtureFlux
export PowerReaction, SinhReaction, AffineReaction, BilinearReaction2, RegionAffineReaction, BipolarReaction
export LinearStorage, PowerStorage, BipolarStorage
export ConstSource, GaussSource, XSinYExpZSource, Step1DSource, AffineXSource, NodalSource
export LinearBoundaryReaction, CatalysisBoundaryReaction, PowerBoundaryReaction, BCondition, dirichlet!, neumann!, robin!
export DiamondEdgeReaction, JouleHeatEdgeReaction, LinearBoundaryStorage
export AMGPrecon, BlockJacobiPrecon, JacobiPrecon, ILUZeroPrecon, DeviceKrylov

const LIB = get(ENV, "VFVM_B200_LIB", joinpath(@__DIR__, "..", "voronoifvm.jl_b200", "libvfvmb200.so"))

# ---- constants of include/vfvm_b200.h -------------------------------------------------------------------------------------------
const VFVM_OK = 0
const VFVM_ERR_NAN = -4
const VFVM_ERR_LINSOLVE = -5
const VFVM_ERR_UNREGISTERED = -6
const VFVM_HOST = 0
const VFVM_DEVICE = 1
const SLOT_FLUX, SLOT_REACTION, SLOT_STORAGE, SLOT_SOURCE, SLOT_BREACTION, SLOT_EDGEREACTION, SLOT_BSTORAGE = 0, 1, 2, 3, 4, 5, 6
const VEC_SOLUTION, VEC_OLDSOL, VEC_RESIDUAL, VEC_UPDATE = 0, 1, 2, 3
const KRYLOV_BICGSTAB, KRYLOV_CG, KRYLOV_GMRES = 0, 1, 2
const PRECON_NONE, PRECON_JACOBI, PRECON_BLOCKJACOBI, PRECON_ILU0, PRECON_ILU0_MC, PRECON_AMG = 0, 1, 2, 3, 4, 5
const BC_DIRICHLET, BC_NEUMANN, BC_ROBIN = 1, 2, 3
const TIME_ASSEMBLE, TIME_LINSOLVE_SETUP, TIME_LINSOLVE_SOLVE = 1, 2, 3  # 1-based positions in vfvm_timings

struct UnregisteredPhysicsError <: Exception
    msg::String
end
Base.showerror(io::IO, e::UnregisteredPhysicsError) = print(io, "UnregisteredPhysicsError: ", e.msg)

struct B200Error <: Exception
    code::Int
    msg::String
end
Base.showerror(io::IO, e::B200Error) = print(io, "libvfvmb200 error ", e.code, ": ", e.msg)

"backend tag: `SystemState(B200(), system)`, `solve(system, B200(); ...)`"
struct B200
    device::Int
end
B200() = B200(0)

last_error(h::Ptr{Cvoid}) = unsafe_string(ccall((:vfvm_last_error, LIB), Cstring, (Ptr{Cvoid},), h))

function check(h::Ptr{Cvoid}, rc::Integer)
    rc == VFVM_OK && return nothing
    rc == VFVM_ERR_UNREGISTERED && throw(UnregisteredPhysicsError(last_error(h)))
    throw(B200Error(rc, last_error(h)))
end

# ---- registered physics library (ids and parameter layouts of include/vfvm_b200.h) ------------------------------------------------
abstract type RegisteredPhysics end
abstract type RegisteredFlux <: RegisteredPhysics end
abstract type RegisteredReaction <: RegisteredPhysics end
abstract type RegisteredStorage <: RegisteredPhysics end
abstract type RegisteredSource <: RegisteredPhysics end
abstract type RegisteredBReaction <: RegisteredPhysics end
abstract type RegisteredEdgeReaction <: RegisteredPhysics end
abstract type RegisteredBStorage <: RegisteredPhysics end

physics_slot(::RegisteredFlux) = SLOT_FLUX
physics_slot(::RegisteredReaction) = SLOT_REACTION
physics_slot(::RegisteredStorage) = SLOT_STORAGE
physics_slot(::RegisteredSource) = SLOT_SOURCE
physics_slot(::RegisteredBReaction) = SLOT_BREACTION
physics_slot(::RegisteredEdgeReaction) = SLOT_EDGEREACTION
physics_slot(::RegisteredBStorage) = SLOT_BSTORAGE

expand(x::Number, n) = fill(Float64(x), n)
function expand(x::AbstractVector, n)
    length(x) == n || error("expected $n values, got $(length(x))")
    return Vector{Float64}(x)
end

# flux(f,u,edge,data): u[i,1] = unknown i at edge.node[1], u[i,2] at edge.node[2]
"f_i = D_i (u_iK - u_iL)   examples/Example201_Laplace2D.jl:17-20, Example301:17-20, Example410:20-25"
struct LinearDiffusion{T} <: RegisteredFlux
    D::T
end
LinearDiffusion() = LinearDiffusion(1.0)
physics_id(::LinearDiffusion) = 1
physics_params(p::LinearDiffusion, n) = expand(p.D, n)
function (p::LinearDiffusion)(f, u, edge, data)
    D = expand(p.D, length(f))
    for i in eachindex(f)
        f[i] = D[i] * (u[i, 1] - u[i, 2])
    end
    return nothing
end

"f_i = D_i (u_iK^m - u_iL^m)   Example207:35-38 (m = 2), Example106:49-52"
struct PowerDiffusion{T} <: RegisteredFlux
    D::T
    m::Float64
end
physics_id(::PowerDiffusion) = 2
physics_params(p::PowerDiffusion, n) = vcat(expand(p.D, n), p.m)
function (p::PowerDiffusion)(f, u, edge, data)
    D = expand(p.D, length(f))
    for i in eachindex(f)
        f[i] = D[i] * (u[i, 1]^p.m - u[i, 2]^p.m)
    end
    return nothing
end

"Example110_ReactionDiffusion1D_TwoSpecies.jl:43-50"
struct CrossDiffusion2 <: RegisteredFlux
    eps1::Float64
    eps2::Float64
    c::Float64
end
physics_id(::CrossDiffusion2) = 3
physics_params(p::CrossDiffusion2, n) = [p.eps1, p.eps2, p.c]
function (p::CrossDiffusion2)(f, u, edge, data)
    f[1] = p.eps1 * (u[1, 1] - u[1, 2]) * (p.c + u[2, 1] + u[2, 2])
    f[2] = p.eps2 * (u[2, 1] - u[2, 2]) * (p.c + u[1, 1] + u[1, 2])
    return nothing
end

"Example160 classflux! :43-50"
struct UnipolarSGFlux <: RegisteredFlux
    eps::Float64
    iphi::Int
    ic::Int
end
physics_id(::UnipolarSGFlux) = 4
physics_params(p::UnipolarSGFlux, n) = [p.eps, p.iphi - 1, p.ic - 1]
function (p::UnipolarSGFlux)(f, u, edge, data)
    f[p.iphi] = p.eps * (u[p.iphi, 1] - u[p.iphi, 2])
    bp, bm = fbernoulli_pm(u[p.iphi, 1] - u[p.iphi, 2])
    f[p.ic] = bm * u[p.ic, 1] - bp * u[p.ic, 2]
    return nothing
end

"Example160 sedanflux! :68-77"
struct SedanFlux <: RegisteredFlux
    eps::Float64
    z::Float64
    iphi::Int
    ic::Int
    eps_reg::Float64
end
physics_id(::SedanFlux) = 5
physics_params(p::SedanFlux, n) = [p.eps, p.z, p.iphi - 1, p.ic - 1, p.eps_reg]
function (p::SedanFlux)(f, u, edge, data)
    f[p.iphi] = p.eps * (u[p.iphi, 1] - u[p.iphi, 2])
    mu1 = -log1p(max(-1 + p.eps_reg, -u[p.ic, 1]))
    mu2 = -log1p(max(-1 + p.eps_reg, -u[p.ic, 2]))
    bp, bm = fbernoulli_pm(p.z * 2 * (u[p.iphi, 1] - u[p.iphi, 2]) + (mu1 - mu2))
    f[p.ic] = bm * u[p.ic, 1] - bp * u[p.ic, 2]
    return nothing
end

"Example161 flux! :134-150; species order (iphin, iphip, ipsi) = (1, 2, 3)"
Base.@kwdef struct BipolarSGFlux <: RegisteredFlux
    lam::Float64 = 0.1
    mun::Float64 = 1.0
    mup::Float64 = 10.0
    zn::Float64 = -1.0
    zp::Float64 = 1.0
    En::Float64 = 1.0
    Ep::Float64 = 0.0
end
physics_id(::BipolarSGFlux) = 6
physics_params(p::BipolarSGFlux, n) = [p.lam, p.mun, p.mup, p.zn, p.zp, p.En, p.Ep, 0.0, 1.0, 2.0]
bipolar_density(z, phi, psi, E) = exp(z * (phi - psi + E))
function (p::BipolarSGFlux)(f, u, edge, data)
    dpsi = u[3, 1] - u[3, 2]
    f[3] = p.lam^2 * dpsi
    bp, bm = fbernoulli_pm(dpsi)
    nK, nL = bipolar_density(p.zn, u[1, 1], u[3, 1], p.En), bipolar_density(p.zn, u[1, 2], u[3, 2], p.En)
    f[1] = -p.zn * p.mun * (bm * nL - bp * nK)
    pK, pL = bipolar_density(p.zp, u[2, 1], u[3, 1], p.Ep), bipolar_density(p.zp, u[2, 2], u[3, 2], p.Ep)
    f[2] = -p.zp * p.mup * (bp * pL - bm * pK)
    return nothing
end

"Maxwell-Stefan mixture flux of DevEx005_Mixture.jl:74-104; the nspec x nspec system is solved in the callback by `inplace_linsolve!`"
struct MixtureFlux <: RegisteredFlux
    DKnudsen::Vector{Float64}
    DBinary::Matrix{Float64}
end
physics_id(::MixtureFlux) = 7
function physics_params(p::MixtureFlux, n)
    DB = copy(p.DBinary)
    for i in 1:n
        DB[i, i] = 1.0   # never read
    end
    return vcat(p.DKnudsen, vec(permutedims(DB)))
end
function (p::MixtureFlux)(f, u, edge, data)
    n = length(p.DKnudsen)
    T = eltype(u)
    M = zeros(T, n, n)
    du = zeros(T, n)
    au = zeros(T, n)
    ipiv = zeros(Int, n)
    for i in 1:n
        M[i, i] = 1.0 / p.DKnudsen[i]
        du[i] = u[i, 1] - u[i, 2]
        au[i] = 0.5 * (u[i, 1] + u[i, 2])
    end
    for i in 1:n, j in 1:n
        if i != j
            M[i, i] += au[j] / p.DBinary[i, j]
            M[i, j] = -au[i] / p.DBinary[i, j]
        end
    end
    VoronoiFVM.inplace_linsolve!(M, du, ipiv)
    for i in 1:n
        f[i] = du[i]
    end
    return nothing
end

# reaction(f,u,node,data)
"f_i = k_i u_i^{p_i}   Example207:32-34"
struct PowerReaction{T, S} <: RegisteredReaction
    k::T
    p::S
end
physics_id(::PowerReaction) = 1
physics_params(r::PowerReaction, n) = vcat(expand(r.k, n), expand(r.p, n))
function (r::PowerReaction)(f, u, node, data)
    k, p = expand(r.k, length(f)), expand(r.p, length(f))
    for i in eachindex(f)
        f[i] = k[i] * u[i]^p[i]
    end
    return nothing
end

"f_i = k_i (exp(u_i) - exp(-u_i))   Example105:55-58"
struct SinhReaction{T} <: RegisteredReaction
    k::T
end
physics_id(::SinhReaction) = 2
physics_params(r::SinhReaction, n) = expand(r.k, n)
function (r::SinhReaction)(f, u, node, data)
    k = expand(r.k, length(f))
    for i in eachindex(f)
        f[i] = k[i] * (exp(u[i]) - exp(-u[i]))
    end
    return nothing
end

"f = R u + r0   Example210:27-31, Example160 reaction! :60-66"
struct AffineReaction <: RegisteredReaction
    R::Matrix{Float64}
    r0::Vector{Float64}
end
physics_id(::AffineReaction) = 3
physics_params(r::AffineReaction, n) = vcat(vec(permutedims(r.R)), r.r0)  # row-major
function (r::AffineReaction)(f, u, node, data)
    for i in eachindex(f)
        f[i] = r.r0[i]
        for j in eachindex(f)
            f[i] += r.R[i, j] * u[j]
        end
    end
    return nothing
end

"f_1 = k u_1 u_2, f_2 = -k u_1 u_2   Example110:38-42"
struct BilinearReaction2 <: RegisteredReaction
    k::Float64
end
physics_id(::BilinearReaction2) = 4
physics_params(r::BilinearReaction2, n) = [r.k]
function (r::BilinearReaction2)(f, u, node, data)
    f[1] = r.k * u[1] * u[2]
    f[2] = -r.k * u[1] * u[2]
    return nothing
end

"f = R_r u + r0_r in cell region r   Example221:54-66"
struct RegionAffineReaction <: RegisteredReaction
    R::Vector{Matrix{Float64}}
    r0::Vector{Vector{Float64}}
end
physics_id(::RegionAffineReaction) = 6
function physics_params(r::RegionAffineReaction, n)
    out = Float64[length(r.R)]
    for (R, r0) in zip(r.R, r.r0)
        append!(out, vec(permutedims(R)))
        append!(out, r0)
    end
    return out
end
function (r::RegionAffineReaction)(f, u, node, data)
    R, r0 = r.R[node.region], r.r0[node.region]
    for i in eachindex(f)
        f[i] = r0[i]
        for j in eachindex(f)
            f[i] += R[i, j] * u[j]
        end
    end
    return nothing
end

"Example161 reaction! :109-132; doping[r] = C in cell region r"
Base.@kwdef struct BipolarReaction <: RegisteredReaction
    doping::Vector{Float64}
    zn::Float64 = -1.0
    zp::Float64 = 1.0
    En::Float64 = 1.0
    Ep::Float64 = 0.0
    r0::Float64 = 1.0
end
physics_id(::BipolarReaction) = 5
physics_params(r::BipolarReaction, n) = vcat([r.zn, r.zp, r.En, r.Ep, r.r0, 0.0, 1.0, 2.0, Float64(length(r.doping))], r.doping)
function (r::BipolarReaction)(f, u, node, data)
    n = bipolar_density(r.zn, u[1], u[3], r.En)
    p = bipolar_density(r.zp, u[2], u[3], r.Ep)
    f[3] = -(r.doping[node.region] + r.zn * n + r.zp * p)
    recomb = (r.r0 + 1.0 / (n + p)) * (n * p * (1.0 - exp(u[1] - u[2])))
    f[1] = r.zn * recomb
    f[2] = r.zp * recomb
    return nothing
end

# storage(f,u,node,data)
"f_i = c_i u_i   Example207:44-47, Example160 storage! :52-58"
struct LinearStorage{T} <: RegisteredStorage
    c::T
end
LinearStorage() = LinearStorage(1.0)
physics_id(::LinearStorage) = 1
physics_params(s::LinearStorage, n) = expand(s.c, n)
function (s::LinearStorage)(f, u, node, data)
    c = expand(s.c, length(f))
    for i in eachindex(f)
        f[i] = c[i] * u[i]
    end
    return nothing
end

"f_i = (eps_i + u_i)^(1/m_i)   Example107:52-55"
struct PowerStorage{T, S} <: RegisteredStorage
    eps::T
    m::S
end
physics_id(::PowerStorage) = 2
physics_params(s::PowerStorage, n) = vcat(expand(s.eps, n), expand(s.m, n))
function (s::PowerStorage)(f, u, node, data)
    e, m = expand(s.eps, length(f)), expand(s.m, length(f))
    for i in eachindex(f)
        f[i] = (e[i] + u[i])^(1.0 / m[i])
    end
    return nothing
end

"Example161 storage! :163-170"
Base.@kwdef struct BipolarStorage <: RegisteredStorage
    zn::Float64 = -1.0
    zp::Float64 = 1.0
    En::Float64 = 1.0
    Ep::Float64 = 0.0
end
physics_id(::BipolarStorage) = 3
physics_params(s::BipolarStorage, n) = [s.zn, s.zp, s.En, s.Ep, 0.0, 1.0, 2.0]
function (s::BipolarStorage)(f, u, node, data)
    f[1] = s.zn * bipolar_density(s.zn, u[1], u[3], s.En)
    f[2] = s.zp * bipolar_density(s.zp, u[2], u[3], s.Ep)
    return nothing
end

# source(f,node,data)
struct ConstSource{T} <: RegisteredSource
    s::T
end
physics_id(::ConstSource) = 1
physics_params(s::ConstSource, n) = expand(s.s, n)
function (s::ConstSource)(f, node, data)
    v = expand(s.s, length(f))
    for i in eachindex(f)
        f[i] = v[i]
    end
    return nothing
end

"f_sp = exp(-a |x - c|^2)   Example207:39-43, Example210:39-44"
struct GaussSource <: RegisteredSource
    species::Int
    a::Float64
    center::NTuple{3, Float64}
end
physics_id(::GaussSource) = 2
physics_params(s::GaussSource, n) = [s.species - 1, s.a, s.center...]
function (s::GaussSource)(f, node, data)
    r2 = 0.0
    for d in 1:length(node.coord[:, node.index])
        r2 += (node[d] - s.center[d])^2
    end
    f[s.species] = exp(-s.a * r2)
    return nothing
end

"f_sp = x sin(b y) exp(z)   Example301:22-26"
struct XSinYExpZSource <: RegisteredSource
    species::Int
    b::Float64
end
physics_id(::XSinYExpZSource) = 3
physics_params(s::XSinYExpZSource, n) = [s.species - 1, s.b]
function (s::XSinYExpZSource)(f, node, data)
    f[s.species] = node[1] * sin(s.b * node[2]) * exp(node[3])
    return nothing
end

"f_sp = x <= x0 ? lo : hi   Example105:45-52"
struct Step1DSource <: RegisteredSource
    species::Int
    x0::Float64
    lo::Float64
    hi::Float64
end
physics_id(::Step1DSource) = 4
physics_params(s::Step1DSource, n) = [s.species - 1, s.x0, s.lo, s.hi]
function (s::Step1DSource)(f, node, data)
    f[s.species] = node[1] <= s.x0 ? s.lo : s.hi
    return nothing
end

"f_i = a_i + b_i x   Example110:51-55"
struct AffineXSource <: RegisteredSource
    a::Vector{Float64}
    b::Vector{Float64}
end
physics_id(::AffineXSource) = 5
physics_params(s::AffineXSource, n) = vcat(s.a, s.b)
function (s::AffineXSource)(f, node, data)
    for i in eachindex(f)
        f[i] = s.a[i] + s.b[i] * node[1]
    end
    return nothing
end

"source given as an n x N table (any u-independent host callback can be tabulated once and uploaded)"
struct NodalSource <: RegisteredSource
    table::Matrix{Float64}
end
physics_id(::NodalSource) = 6
physics_params(s::NodalSource, n) = Float64[]
function (s::NodalSource)(f, node, data)
    for i in eachindex(f)
        f[i] = s.table[i, node.index]
    end
    return nothing
end

# bcondition(f,u,bnode,data)
"if bnode.region == region: f = R u   Example215:33-42"
struct LinearBoundaryReaction <: RegisteredBReaction
    region::Int
    R::Matrix{Float64}
end
physics_id(::LinearBoundaryReaction) = 1
physics_params(b::LinearBoundaryReaction, n) = vcat(Float64(b.region), vec(permutedims(b.R)))
function (b::LinearBoundaryReaction)(f, u, bnode, data)
    if bnode.region == b.region
        for i in eachindex(f), j in eachindex(f)
            f[i] += b.R[i, j] * u[j]
        end
    end
    return nothing
end

"if bnode.region == region: f_i = k_i u_i^{p_i}   Example226_BoundaryIntegral.jl:42-47"
struct PowerBoundaryReaction <: RegisteredBReaction
    region::Int
    k::Vector{Float64}
    p::Vector{Float64}
end
physics_id(::PowerBoundaryReaction) = 3
physics_params(b::PowerBoundaryReaction, n) = vcat(Float64(b.region), b.k, b.p)
function (b::PowerBoundaryReaction)(f, u, bnode, data)
    if bnode.region == b.region
        for i in eachindex(f)
            b.k[i] != 0 && (f[i] = b.k[i] * u[i]^b.p[i])
        end
    end
    return nothing
end

"Example115 breaction! :125-135 on boundary region `region` (species A, B in the bulk, C on the surface)"
Base.@kwdef struct CatalysisBoundaryReaction <: RegisteredBReaction
    region::Int
    S::Float64 = 0.01
    kp_AC::Float64 = 100.0
    km_AC::Float64 = 1.0
    kp_BC::Float64 = 0.1
    km_BC::Float64 = 1.0
    iA::Int = 1
    iB::Int = 2
    iC::Int = 3
end
physics_id(::CatalysisBoundaryReaction) = 2
physics_params(b::CatalysisBoundaryReaction, n) = [b.region, b.S, b.kp_AC, b.km_AC, b.kp_BC, b.km_BC, b.iA - 1, b.iB - 1, b.iC - 1]
function (b::CatalysisBoundaryReaction)(f, u, bnode, data)
    if bnode.region == b.region
        rac = b.kp_AC * u[b.iA] * (1 - u[b.iC]) - b.km_AC * u[b.iC]
        rbc = b.kp_BC * u[b.iB] * (1 - u[b.iC]) - b.km_BC * u[b.iC]
        f[b.iA] = b.S * rac
        f[b.iB] = b.S * rbc
        f[b.iC] = -rbc - rac
    end
    return nothing
end

# edgereaction(f,u,edge,data)
"f_i = c_i h^2 / (2 dim), h = meas(edge)   DevEx002_EdgeReaction.jl:83-87"
struct DiamondEdgeReaction{T} <: RegisteredEdgeReaction
    c::T
end
physics_id(::DiamondEdgeReaction) = 1
physics_params(r::DiamondEdgeReaction, n) = expand(r.c, n)
function (r::DiamondEdgeReaction)(f, u, edge, data)
    c = expand(r.c, length(f))
    h = meas(edge)
    dim = size(edge.coord, 1)
    for i in eachindex(f)
        f[i] = c[i] * h^2 / (2 * dim)
    end
    return nothing
end

"f_iT = -kappa (u_iphi,K - u_iphi,L)^2   Example206_JouleHeat.jl:83-86"
struct JouleHeatEdgeReaction <: RegisteredEdgeReaction
    kappa::Float64
    iphi::Int
    iT::Int
end
physics_id(::JouleHeatEdgeReaction) = 2
physics_params(r::JouleHeatEdgeReaction, n) = [r.kappa, r.iphi - 1, r.iT - 1]
function (r::JouleHeatEdgeReaction)(f, u, edge, data)
    f[r.iT] = -r.kappa * (u[r.iphi, 1] - u[r.iphi, 2]) * (u[r.iphi, 1] - u[r.iphi, 2])
    return nothing
end

# bstorage(f,u,bnode,data)
"if bnode.region == region: f_i = c_i u_i   Example115:138-143, Example311:78-83"
struct LinearBoundaryStorage <: RegisteredBStorage
    region::Int
    c::Vector{Float64}
end
physics_id(::LinearBoundaryStorage) = 1
physics_params(s::LinearBoundaryStorage, n) = vcat(Float64(s.region), s.c)
function (s::LinearBoundaryStorage)(f, u, bnode, data)
    if bnode.region == s.region
        for i in eachindex(f)
            s.c[i] != 0 && (f[i] = s.c[i] * u[i])
        end
    end
    return nothing
end

"C mirror of `vfvm_bc_entry` (include/vfvm_b200.h): one boundary_dirichlet!/neumann!/robin! call, src/vfvm_physics.jl:487-564"
struct BCEntry
    kind::Int32
    species::Int32   # 0-based
    region::Int32    # 0 = all boundary regions (region = bnode.region)
    has_ramp::Int32
    value::Float64
    factor::Float64
    t0::Float64
    t1::Float64
    v0::Float64
    v1::Float64
end

"""
A `bcondition` callback made of `boundary_dirichlet!` / `boundary_neumann!` / `boundary_robin!` calls, optionally after a registered
boundary reaction.  On the CPU path it performs exactly those calls; on the device its entries go to `vfvm_set_bc_entries`.
"""
struct BCondition <: RegisteredBReaction
    reaction::Union{Nothing, LinearBoundaryReaction, CatalysisBoundaryReaction, PowerBoundaryReaction}
    entries::Vector{BCEntry}
end
BCondition() = BCondition(nothing, BCEntry[])
BCondition(r::Union{LinearBoundaryReaction, CatalysisBoundaryReaction, PowerBoundaryReaction}) = BCondition(r, BCEntry[])
physics_id(b::BCondition) = b.reaction === nothing ? 0 : physics_id(b.reaction)
physics_params(b::BCondition, n) = b.reaction === nothing ? Float64[] : physics_params(b.reaction, n)
function push_entry!(b::BCondition, kind, species, region, value, factor, rmp)
    t0, t1, v0, v1 = rmp === nothing ? (0.0, 0.0, 0.0, 0.0) : (Float64(rmp.dt[1]), Float64(rmp.dt[2]), Float64(rmp.du[1]), Float64(rmp.du[2]))
    push!(b.entries, BCEntry(kind, species - 1, region === nothing ? 0 : region, rmp === nothing ? 0 : 1, Float64(value), Float64(factor), t0, t1, v0, v1))
    return b
end
"`ramp = (dt = (t0, t1), du = (v0, v1))` makes the value `ramp(bnode.time; dt, du)` (src/vfvm_physics.jl:516-525)"
dirichlet!(b::BCondition; species = 1, region = nothing, value = 0.0, ramp = nothing) = push_entry!(b, BC_DIRICHLET, species, region, value, 0.0, ramp)
neumann!(b::BCondition; species = 1, region = nothing, value = 0.0, ramp = nothing) = push_entry!(b, BC_NEUMANN, species, region, value, 0.0, ramp)
robin!(b::BCondition; species = 1, region = nothing, factor = 0.0, value = 0.0, ramp = nothing) = push_entry!(b, BC_ROBIN, species, region, value, factor, ramp)
function (b::BCondition)(f, u, bnode, data)
    b.reaction === nothing || b.reaction(f, u, bnode, data)
    for e in b.entries
        region = e.region == 0 ? bnode.region : Int(e.region)
        val = e.has_ramp == 1 ? ramp(bnode.time; dt = (e.t0, e.t1), du = (e.v0, e.v1)) : e.value
        if e.kind == BC_DIRICHLET
            boundary_dirichlet!(f, u, bnode, Int(e.species) + 1, region, val)
        elseif e.kind == BC_NEUMANN
            boundary_neumann!(f, u, bnode, Int(e.species) + 1, region, val)
        else
            boundary_robin!(f, u, bnode, Int(e.species) + 1, region, e.factor, val)
        end
    end
    return nothing
end

# ---- linear solver selection ----------------------------------------------------------------------------------------------------
struct JacobiPrecon end
struct BlockJacobiPrecon end
struct ILUZeroPrecon
    multicolor::Bool
end
ILUZeroPrecon() = ILUZeroPrecon(false)
"aggregation AMG (csrc/amg.cu); options as `vfvm_amg_set_options`: omega, alpha, theta, sweeps, coarse_sweeps, wdepth (NaN keeps a value)"
Base.@kwdef struct AMGPrecon
    omega::Float64 = NaN
    alpha::Float64 = NaN
    theta::Float64 = NaN
    sweeps::Float64 = NaN
    coarse_sweeps::Float64 = NaN
    wdepth::Float64 = NaN
end
precon_id(::Nothing) = PRECON_NONE
precon_id(::JacobiPrecon) = PRECON_JACOBI
precon_id(::BlockJacobiPrecon) = PRECON_BLOCKJACOBI
precon_id(p::ILUZeroPrecon) = p.multicolor ? PRECON_ILU0_MC : PRECON_ILU0
precon_id(::AMGPrecon) = PRECON_AMG

"explicit device Krylov selection for `SolverControl.method_linear`"
Base.@kwdef struct DeviceKrylov
    krylov::Int = KRYLOV_BICGSTAB
    precs::Any = BlockJacobiPrecon()
    restart::Int = 30
    direct_like::Bool = false   # stand-in for the default sparse LU: BiCGStab driven to 1e-13, non-convergence is an error
end

"maps `control.method_linear` (src/vfvm_solvercontrol.jl:106) to the device solver; LinearSolve's Krylov algorithms keep their meaning"
function device_method(method_linear)
    method_linear isa DeviceKrylov && return method_linear
    if method_linear === nothing || method_linear isa LinearSolve.AbstractFactorization
        return DeviceKrylov(; direct_like = true)   # the reference default is UMFPACK (src/vfvm_solver.jl:34-41)
    end
    if method_linear isa LinearSolve.KrylovJL
        name = string(method_linear.KrylovAlg)
        krylov = occursin("cg", name) && !occursin("bicg", name) ? KRYLOV_CG : (occursin("gmres", name) ? KRYLOV_GMRES : KRYLOV_BICGSTAB)
        precs = method_linear.precs
        precs = precs === nothing || precs isa Function ? BlockJacobiPrecon() : precs
        return DeviceKrylov(; krylov, precs, restart = method_linear.gmres_restart)
    end
    throw(ArgumentError("method_linear = $(typeof(method_linear)) has no device counterpart"))
end

# ---- the device-backed matrix and state -----------------------------------------------------------------------------------------
"""
Stands in the `matrix` field of `VoronoiFVM.SystemState` (type parameter `TMatrix`, src/vfvm_state.jl:16-19).  The Jacobian lives in
HBM behind the handle; `SparseMatrixCSC(A)` downloads it in the scalar pattern the reference would hold.
"""
mutable struct B200Matrix <: AbstractMatrix{Float64}
    h::Ptr{Cvoid}
    n::Int        # number of dofs
    nspecies::Int
    nnodes::Int
    physics_version::UInt64
    method_key::Any
end
Base.size(A::B200Matrix) = (A.n, A.n)
Base.getindex(A::B200Matrix, i::Integer, j::Integer) = SparseMatrixCSC(A)[i, j]

function SparseArrays.SparseMatrixCSC(A::B200Matrix)
    nrows, nnz = Ref{Int64}(0), Ref{Int64}(0)
    check(A.h, ccall((:vfvm_pattern_size, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), A.h, nrows, nnz))
    colptr, rowval, nzval = Vector{Int64}(undef, A.n + 1), Vector{Int64}(undef, nnz[]), Vector{Float64}(undef, nnz[])
    check(A.h, ccall((:vfvm_get_pattern_csc, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), A.h, colptr, rowval))
    check(A.h, ccall((:vfvm_get_nzval_csc, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), A.h, nzval, VFVM_HOST))
    return SparseMatrixCSC(A.n, A.n, colptr .+ 1, rowval .+ 1, nzval)
end

function destroy!(A::B200Matrix)
    if A.h != C_NULL
        ccall((:vfvm_destroy, LIB), Cvoid, (Ptr{Cvoid},), A.h)
        A.h = C_NULL
    end
    return nothing
end

const B200State = SystemState{Tv, Tp, B200Matrix} where {Tv, Tp}

coordsys_id(grid) = begin
    cs = grid[CoordinateSystem]
    cs <: Union{Polar1D, Cylindrical2D} ? 1 : (cs <: Spherical1D ? 2 : 0)
end

registered(cb, slot) = begin
    cb === VoronoiFVM.nofunc && return nothing
    cb isa RegisteredPhysics && physics_slot(cb) == slot && return cb
    throw(UnregisteredPhysicsError("$(typeof(cb)) is not a registered device callback; the B200 backend does not fall back to the CPU"))
end

"uploads ids, parameter blocks, boundary entries and the legacy boundary tables of `system` (enable_species!, boundary_dirichlet!, ...)"
function push_physics!(A::B200Matrix, system)
    h, n, ph = A.h, A.nspecies, system.physics
    for name in (:bflux, :bsource, :boutflow, :generic_operator)
        getproperty(ph, name) === VoronoiFVM.nofunc || throw(UnregisteredPhysicsError("physics callback `$name` is outside the device scope"))
    end
    for (slot, cb) in ((SLOT_FLUX, ph.flux), (SLOT_REACTION, ph.reaction), (SLOT_STORAGE, ph.storage), (SLOT_SOURCE, ph.source), (SLOT_BREACTION, ph.breaction),
                       (SLOT_EDGEREACTION, ph.edgereaction), (SLOT_BSTORAGE, ph.bstorage))
        r = registered(cb, slot)
        id = r === nothing ? 0 : physics_id(r)
        params = r === nothing ? Float64[] : Vector{Float64}(physics_params(r, n))
        check(h, ccall((:vfvm_set_physics, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Cint), h, slot, id, params, length(params)))
        if r isa NodalSource
            check(h, ccall((:vfvm_set_nodal_source, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, r.table))
        end
    end
    bc = registered(ph.breaction, SLOT_BREACTION)
    entries = bc isa BCondition ? bc.entries : BCEntry[]
    check(h, ccall((:vfvm_set_bc_entries, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{BCEntry}), h, length(entries), entries))
    nbreg = size(system.boundary_factors, 2)
    check(h, ccall((:vfvm_set_legacy_bc, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}), h, nbreg,
                   Matrix{Float64}(system.boundary_factors), Matrix{Float64}(system.boundary_values)))
    return nothing
end

# ---- unknown storage: the device always works on the dense n x N layout (dofs of undefined species are identity rows and stay zero).
# A SparseSolutionArray (unknown_storage = :sparse, src/vfvm_sparsesolution.jl:10-172) stores only the defined dofs in the CSC layout of
# `system.node_dof`; it is scattered on the way up and gathered on the way down.  `Matrix{Float64}(U)` would not do: reading an undefined
# dof of a sparse solution yields NaN (:156-166).
to_dense(U::AbstractMatrix) = Matrix{Float64}(U)
function to_dense(U::VoronoiFVM.SparseSolutionArray)
    A = U.u
    d = zeros(Float64, size(A)...)
    for K in 1:size(A, 2), k in A.colptr[K]:(A.colptr[K + 1] - 1)
        d[A.rowval[k], K] = A.nzval[k]
    end
    return d
end
from_dense!(U::AbstractMatrix, d::Matrix{Float64}) = (U .= d; U)
function from_dense!(U::VoronoiFVM.SparseSolutionArray, d::Matrix{Float64})
    A = U.u
    for K in 1:size(A, 2), k in A.colptr[K]:(A.colptr[K + 1] - 1)
        A.nzval[k] = d[A.rowval[k], K]
    end
    return U
end
dense_size(U) = size(U)

"""
    SystemState(B200(), system; data = system.physics.data)

Device twin of `SystemState(system)` (src/vfvm_state.jl:99-157): uploads the grid once, builds form factors (K1/K2) and the sparsity
pattern (K3) on the device, and returns a `VoronoiFVM.SystemState` whose `matrix` is a `B200Matrix`.
"""
function VoronoiFVM.SystemState(backend::B200, system::VoronoiFVM.AbstractSystem; data = system.physics.data, params = zeros(system.num_parameters))
    VoronoiFVM._complete!(system)
    system.num_parameters == 0 || throw(UnregisteredPhysicsError("parameter derivatives (dudp) are outside the device scope"))
    grid = system.grid
    dim = dim_space(grid)
    coord = Matrix{Float64}(grid[Coordinates])
    cellnodes = Matrix{Int32}(grid[CellNodes]) .- Int32(1)
    cellregions = Vector{Int32}(grid[CellRegions])
    bfacenodes = Matrix{Int32}(grid[BFaceNodes]) .- Int32(1)
    bfaceregions = Vector{Int32}(grid[BFaceRegions])
    nspec = num_species(system)
    nnodes = size(coord, 2)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:vfvm_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), backend.device, href)
    rc == VFVM_OK || throw(B200Error(rc, "vfvm_create failed: no CUDA device / driver (the device path has no CPU fallback)"))
    h = href[]
    check(h, ccall((:vfvm_set_grid, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                   h, dim, coordsys_id(grid), nnodes, size(cellnodes, 2), size(bfacenodes, 2), coord, cellnodes, cellregions, bfacenodes, bfaceregions))
    check(h, ccall((:vfvm_build_geometry, LIB), Cint, (Ptr{Cvoid},), h))
    region_species = UInt8[system.region_species[i, r] > 0 ? 1 : 0 for i in 1:nspec, r in 1:num_cellregions(grid)]
    check(h, ccall((:vfvm_set_system, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}), h, nspec, region_species))
    bregion_species = UInt8[system.bregion_species[i, r] > 0 ? 1 : 0 for i in 1:nspec, r in 1:num_bfaceregions(grid)]   # enable_boundary_species!
    if any(!=(0), bregion_species)
        check(h, ccall((:vfvm_set_boundary_species, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}), h, size(bregion_species, 2), bregion_species))
    end
    A = B200Matrix(h, nspec * nnodes, nspec, nnodes, UInt64(0), nothing)
    finalizer(destroy!, A)
    push_physics!(A, system)
    check(h, ccall((:vfvm_build_pattern, LIB), Cint, (Ptr{Cvoid},), h))
    solution, residual, update = unknowns(system), unknowns(system), unknowns(system)
    return SystemState(system, data, solution, A, nothing, typeof(solution)[], residual, update, nothing, Vector{Float64}(params), zero(UInt64), nothing)
end

"`solve(system, B200(); kwargs...)`: the reference's `solve(system; kwargs...)` (src/vfvm_solver.jl:665-668) on a device state"
function CommonSolve.solve(system::VoronoiFVM.AbstractSystem, backend::B200; data = system.physics.data, kwargs...)
    state = SystemState(backend, system; data)
    return CommonSolve.solve!(state; kwargs...)
end

# ---- eval_and_assemble: the reference's signature, dispatched on the matrix type (src/vfvm_assembly.jl:520-534) ---------------------
# U, UOld, F are dense arrays or SparseSolutionArrays (both are AbstractMatrix): to_dense / from_dense! above.
function VoronoiFVM.eval_and_assemble(
        system,
        U::AbstractMatrix{Tv},
        UOld::AbstractMatrix{Tv},
        F::AbstractMatrix{Tv},
        matrix::B200Matrix,
        generic_matrix::Union{AbstractMatrix, Nothing},
        dudp,
        time,
        tstep,
        λ,
        data,
        params::AbstractVector;
        edge_cutoff = 0.0,
    ) where {Tv}
    push_physics!(matrix, system)
    u, uold, f = to_dense(U), to_dense(UOld), Matrix{Float64}(undef, dense_size(F)...)
    rc = ccall((:vfvm_eval_res_jac, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cdouble, Cdouble, Cdouble),
               matrix.h, u, uold, f, VFVM_HOST, Float64(time), Float64(tstep), Float64(λ))
    rc == VFVM_ERR_NAN && error("trying to assemble NaN")   # src/vfvm_assembly.jl:10-12
    check(matrix.h, rc)
    from_dense!(F, f)
    return 0, 0, 1   # (ncalloc, nballoc, neval)
end

function linear_setup!(A::B200Matrix, control)
    m = device_method(control.method_linear)
    key = (m.krylov, precon_id(m.precs), m.restart, m.precs)
    fresh = A.method_key != key
    if fresh
        check(A.h, ccall((:vfvm_linsolve_setup, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint), A.h, m.krylov, precon_id(m.precs), m.restart))
        if m.precs isa AMGPrecon
            opts = Float64[m.precs.omega, m.precs.alpha, m.precs.theta, m.precs.sweeps, m.precs.coarse_sweeps, m.precs.wdepth]
            check(A.h, ccall((:vfvm_amg_set_options, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), A.h, opts, length(opts)))
        end
        A.method_key = key
    end
    tol = m.direct_like ? (0.0, 1.0e-13, 20000) : (control.abstol_linear, control.reltol_linear, control.maxiters_linear)
    return m, fresh, tol
end

function timings(h::Ptr{Cvoid})
    t = zeros(Float64, 8)
    ccall((:vfvm_timings, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, t)
    return t
end

"device solve of A * UPDATE = RESIDUAL with the resident vectors; returns the iteration count"
function device_linsolve!(A::B200Matrix, nlhistory, control, reuse_precs)
    m, fresh, (abstol, reltol, maxiters) = linear_setup!(A, control)
    reuse = reuse_precs && !fresh
    # nlu counts the set-ups that really happen: the device keeps only an ILU factorisation across solves (csrc/linsolve.cu)
    keeps = precon_id(m.precs) in (PRECON_ILU0, PRECON_ILU0_MC)
    (reuse && keeps) || (nlhistory.nlu += 1)
    iters, resnorm = Ref{Cint}(0), Ref{Cdouble}(0.0)
    rc = ccall((:vfvm_linsolve, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}),
               A.h, abstol, reltol, maxiters, reuse ? 1 : 0, iters, resnorm)
    t = timings(A.h)
    nlhistory.tlinsolve_setup += t[TIME_LINSOLVE_SETUP] * 1.0e-3
    nlhistory.tlinsolve_solve += t[TIME_LINSOLVE_SOLVE] * 1.0e-3
    nlhistory.nlin = iters[]
    rc == VFVM_ERR_LINSOLVE && throw(LinearSolverError())
    check(A.h, rc)
    if m.direct_like
        converged, rhsnorm = Ref{Cint}(0), Ref{Cdouble}(0.0)
        check(A.h, ccall((:vfvm_linsolve_status, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cdouble}), A.h, converged, rhsnorm))
        converged[] == 1 || throw(LinearSolverError())
    end
    return iters[]
end

# ---- _solve_linear!: the reference's signature, dispatched on the matrix type (src/vfvm_linsolve.jl:6) ------------------------------
# Reached when a caller other than the device `solve_step!` below holds host vectors: b goes up, the update comes back.
function VoronoiFVM._solve_linear!(u, state, nlhistory, control, method_linear, A::B200Matrix, b, reuse_precs)
    check(A.h, ccall((:vfvm_set_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), A.h, VEC_RESIDUAL, Vector{Float64}(b), VFVM_HOST))
    device_linsolve!(A, nlhistory, control, reuse_precs)
    upd = Vector{Float64}(undef, length(u))
    check(A.h, ccall((:vfvm_get_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), A.h, VEC_UPDATE, upd, VFVM_HOST))
    u .= upd
    return nothing
end

# ---- solve_step!: the reference's signature and control flow (src/vfvm_solver.jl:13-222), vectors resident on the device ------------
function VoronoiFVM.solve_step!(
        state::B200State,
        solution,
        oldsol,
        control,
        time,
        tstep,
        embedparam,
        params,
        istep_factorize
    )
    A = state.matrix
    h = A.h
    nlhistory = NewtonSolverHistory()
    tasm = 0.0
    tlinsolve = 0.0
    t = @elapsed begin
        push_physics!(A, state.system)
        # solution .= oldsol ; _initialize!(solution, ...)                                                     (:28, :31)
        check(h, ccall((:vfvm_set_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), h, VEC_OLDSOL, to_dense(oldsol), VFVM_HOST))
        check(h, ccall((:vfvm_copy_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Cint), h, VEC_SOLUTION, VEC_OLDSOL))
        check(h, ccall((:vfvm_init_dirichlet, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), h, Float64(time), Float64(embedparam)))
        oldnorm = 1.0
        converged = false
        damp = 1.0
        rnorm = 0.0
        ninf, n1 = Ref{Cdouble}(0.0), Ref{Cdouble}(0.0)
        if !state.system.is_linear
            damp = control.damp_initial
            check(h, ccall((:vfvm_vector_norms, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h, VEC_SOLUTION, ninf, n1))
            rnorm = n1[]                                                                                      # control.rnorm(solution) (:55)
        end
        nround = 0
        tolx = 0.0
        niter = 1
        norm = 0.0
        while niter <= control.maxiters
            # eval_and_assemble on the resident vectors                                                       (:67-96)
            rc = ccall((:vfvm_assemble, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble), h, Float64(time), Float64(tstep), Float64(embedparam))
            tasm += timings(h)[TIME_ASSEMBLE] * 1.0e-3
            rc == VFVM_ERR_NAN && throw(AssemblyError())
            check(h, rc)
            reuse_precs = (!control.factorize_every_newtonstep && niter > 1) || (istep_factorize % control.factorize_every_timestep != 0)   # :99
            if !control.updatecontrol
                check(h, ccall((:vfvm_vector_norms, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h, VEC_RESIDUAL, ninf, n1))
                norm = ninf[]
            end
            tlinsolve += @elapsed device_linsolve!(A, nlhistory, control, reuse_precs)                        # :105-114
            # dofs(solution) .-= damp * dofs(update), with ||update||_inf and ||solution||_1 from the same pass   (:116, :126, :132)
            check(h, ccall((:vfvm_newton_update, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}), h, damp, ninf, n1))
            if state.system.is_linear
                converged = true
                break
            end
            damp = min(damp * control.damp_growth, 1.0)
            if control.updatecontrol
                norm = ninf[]
            end
            if tolx == 0.0
                tolx = norm * control.reltol
            end
            dnorm = 1.0
            rnorm_new = n1[]
            if rnorm > 1.0e-50
                dnorm = abs((rnorm - rnorm_new) / rnorm)
            end
            nround = dnorm < control.tol_round ? nround + 1 : 0
            if control.log
                push!(nlhistory.l1normdiff, dnorm)
                push!(nlhistory.updatenorm, norm)
            end
            if niter > 1 && norm / oldnorm > 1.0 / control.tol_mono
                converged = false
                break
            end
            if norm < control.abstol || norm < tolx
                converged = true
                break
            end
            oldnorm = norm
            rnorm = rnorm_new
            if nround > control.max_round
                converged = true
                break
            end
            niter = niter + 1
        end
        converged || throw(ConvergenceError())
        sol = Matrix{Float64}(undef, dense_size(solution)...)
        check(h, ccall((:vfvm_get_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), h, VEC_SOLUTION, sol, VFVM_HOST))
        from_dense!(solution, sol)
    end
    if control.log
        nlhistory.time = t
        nlhistory.tlinsolve = tlinsolve
        nlhistory.tasm = tasm
    end
    solution.history = nlhistory
    return solution
end

end # module
