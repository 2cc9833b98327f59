#=
NOT EXECUTED in this repository's environment (no Julia in the image, no network for the package depot).

Times the reference's own CPU implementation of the hot path on the bench configuration cfg3 (Example301 physics on an
nx^3 tensor simplex grid), for anyone who has Julia + VoronoiFVM.jl:

    julia -t auto baseline/reference_timing.jl 65      # nx = 65 is the sample bench.py's CPU baseline uses

It prints the same two figures bench.py reports for its CPU arm: residual+Jacobian assembly throughput in Medges/s
(evaluate_residual_and_jacobian!, src/vfvm_solver.jl:224-243, which calls eval_and_assemble, src/vfvm_assembly.jl:520-643)
and the time of one Newton step (solve, src/vfvm_solver.jl:665-668; the problem is linear, so one step).
=#
using VoronoiFVM, ExtendableGrids, LinearAlgebra, Printf

function main(nx)
    X = range(0, 1; length = nx)
    grid = simplexgrid(X, X, X)
    # threaded assembly needs a partitioned grid (src/vfvm_assembly.jl:561-612)
    if Threads.nthreads() > 1
        grid = partition(grid, PlainMetisPartitioning(npart = 4 * Threads.nthreads()); nodes = true, edges = true)
    end
    flux(f, u, edge, data) = (f[1] = u[1, 1] - u[1, 2]; nothing)                 # examples/Example301_Laplace3D.jl:15-20
    source(f, node, data) = (f[1] = node[1] * sin(5.0 * node[2]) * exp(node[3]); nothing)  # :22-26
    sys = VoronoiFVM.System(grid; flux, source, species = [1], assembly = :edgewise)
    boundary_dirichlet!(sys, 1, 5, 0.0)
    boundary_dirichlet!(sys, 1, 6, 0.0)
    U = unknowns(sys; inival = 0.5)
    state = VoronoiFVM.SystemState(sys)               # the matrix pattern lives in the state: steady-state assemblies below
    VoronoiFVM.evaluate_residual_and_jacobian!(state, U)   # first call: pattern build + compilation (src/vfvm_solver.jl:224-243)
    nedges = num_edges(grid)
    t = minimum(@elapsed(VoronoiFVM.evaluate_residual_and_jacobian!(state, U)) for _ in 1:3)
    @printf("assembly: %d edges, %.3f ms, %.1f Medges/s on %d threads\n", nedges, 1e3 * t, nedges / t / 1e6, Threads.nthreads())
    solve(sys; inival = 0.0)                          # compilation
    t = @elapsed sol = solve(sys; inival = 0.0, log = true)
    h = history(sol)
    @printf("Newton step (default sparse LU): %.1f ms (assembly %.1f ms, linear solve %.1f ms)\n", 1e3 * t, 1e3 * h.tasm, 1e3 * h.tlinsolve)
    return nothing
end

main(length(ARGS) > 0 ? parse(Int, ARGS[1]) : 65)
