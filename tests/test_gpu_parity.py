"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle
on identical seeded inputs.

Bars (north_star): sparsity pattern and edge/node index maps bit-exact; residual and Jacobian entries <= 1e-12 relative;
Newton solutions <= 1e-10.
"""
import math

import numpy as np
import pytest

import vfvm_b200 as v
from vfvm_b200 import physics as ph
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL_ASM = 1.0e-12  # residual / Jacobian entries, relative
TOL_NEWTON = 1.0e-10  # Newton solutions, absolute (solutions are O(1))


def _grid(dim, nx):
    X = np.linspace(0, 1, nx)
    return v.simplexgrid(*([X] * dim))


def _rand_u(sys, seed=20261017, lo=0.1, hi=1.0):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.uniform(lo, hi, (sys.num_species, sys.grid.num_nodes)))


def _compare_assembly(sys, U, UOld=None, time=0.0, tstep=math.inf, embed=0.0):
    st = v.SystemState(sys)
    try:
        F = st.eval_res_jac(U, UOld, time=time, tstep=tstep, embed=embed)
        A = st.matrix("csc")
    finally:
        st.close()
    Fo, Ao = O.OracleSystem(sys).assemble(U, UOld, time=time, tstep=tstep, embed=embed)
    # pattern: bit-exact
    assert A.shape == Ao.shape
    assert np.array_equal(A.indptr, Ao.indptr), "CSC colptr differs"
    assert np.array_equal(A.indices, Ao.indices), "CSC rowval differs"
    # values: 1e-12 relative to the entry.  Entries that are sums (diagonal blocks, residuals, interface edges that carry
    # one form factor per cell region) may cancel, so on top of the relative bound they get an absolute allowance of a few
    # ulps of the magnitude of the summed terms (termscale = sum of |non-penalty entries| of the row).
    coo = Ao.tocoo()
    mag = np.where(np.abs(coo.data) < 1e29, np.abs(coo.data), 0.0)
    termscale = np.bincount(coo.row, weights=mag, minlength=Ao.shape[0])
    err = np.abs(A.data - Ao.data)
    bound = RTOL_ASM * np.abs(Ao.data) + 8 * np.finfo(float).eps * termscale[coo.row]
    assert np.all(err <= bound), f"Jacobian mismatch: max abs err {err.max():.3e}, max err/bound {(err / np.maximum(bound, 1e-300)).max():.3e}"
    f, fo = F.ravel(order="F"), Fo.ravel(order="F")
    fbound = RTOL_ASM * np.abs(fo) + 8 * np.finfo(float).eps * (termscale * max(1.0, np.abs(U).max()) + np.abs(fo))
    assert np.all(np.abs(f - fo) <= fbound), f"residual mismatch: max abs err {np.abs(f - fo).max():.3e}"
    return A, F


# ---------------------------------------------------------------------------------------------- geometry (K1, K2)
@pytest.mark.parametrize("dim,nx", [(1, 33), (2, 17), (3, 9)])
def test_geometry_index_maps_and_factors(dim, nx):
    sys = v.System(_grid(dim, nx), flux=ph.LinearDiffusion(), species=[1])
    st = v.SystemState(sys)
    o = O.OracleSystem(sys)
    try:
        assert st.num_edges == o.num_edges
        assert np.array_equal(st.edgenodes(), o.edgenodes()), "edge->node map differs"
        assert np.array_equal(st.celledges(), o.celledges()), "cell->edge map differs"
        for a, b in ((st.nodefactors(), o.nodefactors()), (st.edgefactors(), o.edgefactors())):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            np.testing.assert_allclose(a[2], b[2], rtol=1e-13, atol=1e-18)
        np.testing.assert_allclose(st.bfacefactors(), o.bfacefactors(), rtol=1e-14)
        # test/test120_norms.jl:95-100: node volumes of the unit cube sum to 1
        assert st.nodefactors()[2].sum() == pytest.approx(1.0, rel=1e-12)
    finally:
        st.close()


def test_geometry_random_tets_and_regions():
    """distorted 3D grid with three cell regions: factors per (region, node/edge), ragged CSC"""
    g = _grid(3, 7)
    rng = np.random.default_rng(7)
    h = 1.0 / 6
    interior = np.all((g.coord > 1e-9) & (g.coord < 1 - 1e-9), axis=0)
    g.coord[:, interior] += rng.uniform(-0.2 * h, 0.2 * h, (3, int(interior.sum())))
    v.cellmask(g, [0, 0, 0.33], [1, 1, 0.67], 2)
    v.cellmask(g, [0, 0, 0.66], [1, 1, 1.0], 3)
    sys = v.System(g, flux=ph.LinearDiffusion(), species=[1])
    st = v.SystemState(sys)
    o = O.OracleSystem(sys)
    try:
        assert np.array_equal(st.edgenodes(), o.edgenodes())
        for a, b in ((st.nodefactors(), o.nodefactors()), (st.edgefactors(), o.edgefactors())):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            np.testing.assert_allclose(a[2], b[2], rtol=1e-12, atol=1e-16)
    finally:
        st.close()


@pytest.mark.parametrize("coordsys", ["cyl1", "sph1", "cyl2"])
def test_geometry_curvilinear(coordsys):
    X = np.linspace(0.1, 1.3, 25)
    if coordsys == "cyl1":
        g = v.circular_symmetric(v.simplexgrid(X))
    elif coordsys == "sph1":
        g = v.spherical_symmetric(v.simplexgrid(X))
    else:
        g = v.circular_symmetric(v.simplexgrid(X, np.linspace(0, 1, 11)))
    sys = v.System(g, flux=ph.LinearDiffusion(), species=[1])
    st = v.SystemState(sys)
    o = O.OracleSystem(sys)
    try:
        np.testing.assert_allclose(st.nodefactors()[2], o.nodefactors()[2], rtol=1e-13)
        np.testing.assert_allclose(st.edgefactors()[2], o.edgefactors()[2], rtol=1e-13, atol=1e-16)
        np.testing.assert_allclose(st.bfacefactors(), o.bfacefactors(), rtol=1e-13)
    finally:
        st.close()


# ---------------------------------------------------------------------------------------------- assembly (K3-K6)
@pytest.mark.parametrize("dim,nx", [(1, 50), (2, 24), (3, 12)])
def test_assembly_laplace(dim, nx):
    """Example201 / Example301 physics: linear diffusion, legacy Dirichlet"""
    sys = v.System(_grid(dim, nx), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0) if dim == 3 else None)
    v.enable_species(sys, 1, [1])
    v.boundary_dirichlet(sys, 1, 1 if dim < 3 else 5, 0.0)
    v.boundary_dirichlet(sys, 1, 2 if dim == 1 else (3 if dim == 2 else 6), 1.0)
    _compare_assembly(sys, _rand_u(sys))


def test_assembly_example207_transient():
    """nonlinear diffusion + reaction + source + storage, implicit Euler step (Example207 physics)"""
    X = np.linspace(0, 1, 41)
    sys = v.System(v.simplexgrid(X, X), flux=ph.PowerDiffusion(1.0e-2, 2), reaction=ph.PowerReaction(1.0, 2.0), source=ph.GaussSource(1, 20.0, (0.5, 0.5)),
                   storage=ph.LinearStorage(1.0))
    v.enable_species(sys, 1, [1])
    v.boundary_dirichlet(sys, 1, 2, 0.1)
    v.boundary_dirichlet(sys, 1, 4, 0.1)
    U = _rand_u(sys)
    _compare_assembly(sys, U, _rand_u(sys, seed=5), tstep=0.01)
    _compare_assembly(sys, U)  # stationary: storage entries are not inserted (value-dependent pattern)


def test_assembly_two_species_coupled():
    """Example110 physics: coupled flux, bilinear reaction, x-dependent source, two species"""
    sys = v.System(v.simplexgrid(np.linspace(0, 1, 101)), reaction=ph.BilinearReaction2(1.0), flux=ph.CrossDiffusion2((0.25, 0.5), 0.01),
                   source=ph.AffineXSource([1.0e-4 * 0.01, 1.0e-4 * 1.01], [1.0e-4, -1.0e-4]), storage=ph.LinearStorage(1.0))
    v.enable_species(sys, 1, [1])
    v.enable_species(sys, 2, [1])
    for sp in (1, 2):
        v.boundary_dirichlet(sys, sp, 1, 1.0)
        v.boundary_dirichlet(sys, sp, 2, 0.0)
    _compare_assembly(sys, _rand_u(sys), _rand_u(sys, seed=3), tstep=0.1)


def test_assembly_unipolar_sedan_callback_bc():
    """Example160: sedanflux! with fbernoulli_pm + log1p, affine reaction, callback Dirichlet with ramp"""
    n = 40
    X = np.arange(0, n + 1) / n
    eps, z, V = 1.0e-3, -1.0, 5.0
    bc = ph.BCondition().dirichlet(species=1, region=1, value=0.0, ramp=((0, 1.0e-2), (0, V))).dirichlet(species=1, region=2, value=0).dirichlet(species=2, region=2, value=0.5)
    R = np.zeros((2, 2))
    R[0, 1] = -2 * z
    for flux in (ph.SedanFlux(eps, z, 1, 2), ph.UnipolarSGFlux(eps, 1, 2)):
        sys = v.System(v.simplexgrid(X), flux=flux, reaction=ph.AffineReaction(R, [z, 0.0]), bcondition=bc, storage=ph.LinearStorage([0.0, 1.0]), species=[1, 2])
        U = _rand_u(sys, lo=0.05, hi=0.9)
        U[0, :] *= 8.0  # potentials large enough to exercise all Bernoulli branches
        _compare_assembly(sys, U, _rand_u(sys, seed=9, lo=0.05, hi=0.9), time=0.004, tstep=1.0e-3)


def test_assembly_bipolar_drift_diffusion_3d():
    """Example161 physics (3 species, SG flux with exp densities, recombination, doping per cell region) on a 3D grid"""
    g = _grid(3, 8)
    v.cellmask(g, [0, 0, 0.3], [1, 1, 0.72], 2)
    v.cellmask(g, [0, 0, 0.7], [1, 1, 1.0], 3)
    bc = ph.BCondition()
    for sp, val in ((1, 0.0), (2, 0.0), (3, 0.5 + math.asinh(10.0 / (2 * math.sqrt(math.exp(-1.0)))))):
        bc.dirichlet(species=sp, region=5, value=val)
    for sp, val in ((1, 0.3), (2, 0.3), (3, 0.5 + math.asinh(-10.0 / (2 * math.sqrt(math.exp(-1.0)))) + 0.3)):
        bc.dirichlet(species=sp, region=6, value=val)
    sys = v.System(g, flux=ph.BipolarSGFlux(), reaction=ph.BipolarReaction([10.0, 0.0, -10.0]), storage=ph.BipolarStorage(), bcondition=bc, species=[1, 2, 3])
    U = _rand_u(sys, lo=-0.5, hi=0.5)
    U[2, :] = np.linspace(3.0, -3.0, g.num_nodes)
    _compare_assembly(sys, U, _rand_u(sys, seed=2, lo=-0.5, hi=0.5), tstep=1.0e-2)
    _compare_assembly(sys, U)


@pytest.mark.parametrize("ns", [4, 5, 10])
def test_assembly_many_species_decoupled(ns):
    """Example410 scaled to 3D: ns decoupled species -> only same-species couplings in the pattern"""
    sys = v.System(_grid(3, 7), flux=ph.LinearDiffusion(np.arange(1, ns + 1) * 0.5), storage=ph.LinearStorage(1.0))
    for i in range(1, ns + 1):
        v.enable_species(sys, i, [1])
        v.boundary_dirichlet(sys, i, 5, 0)
        v.boundary_dirichlet(sys, i, 6, 1)
    A, F = _compare_assembly(sys, _rand_u(sys), _rand_u(sys, seed=4), tstep=0.1)
    coo = A.tocoo()
    assert np.all(coo.row % ns == coo.col % ns)


def test_assembly_boundary_reaction_and_robin():
    """Example215 boundary reaction (off-diagonal species coupling only at boundary nodes of region 2) + Robin/Neumann entries"""
    X = np.arange(0, 11) / 10.0
    k = 1.0
    bc = ph.BCondition(ph.LinearBoundaryReaction(2, [[k, -k], [-k, k]])).robin(species=1, region=4, factor=2.0, value=0.5).neumann(species=2, region=3, value=0.25)
    sys = v.System(v.simplexgrid(X, X), bcondition=bc, flux=ph.LinearDiffusion(1.0e-2), storage=ph.LinearStorage(1.0), species=[1, 2])
    _compare_assembly(sys, _rand_u(sys), _rand_u(sys, seed=8), tstep=0.01)


def test_assembly_nan_is_reported():
    """src/vfvm_assembly.jl:10-12: a NaN Jacobian value is an error (AssemblyError), not a silent result"""
    X = np.linspace(0, 1, 9)
    sys = v.System(v.simplexgrid(X, X), flux=ph.PowerDiffusion(1.0, 0.5), species=[1])  # sqrt of a negative unknown
    st = v.SystemState(sys)
    try:
        U = _rand_u(sys)
        U[0, 10] = -1.0
        with pytest.raises(v._lib.VfvmError) as ei:
            st.eval_res_jac(U)
        assert ei.value.code == v._lib.ERR_NAN
        with pytest.raises(O.AssemblyError):
            O.OracleSystem(sys).assemble(U)
    finally:
        st.close()


def test_assembly_is_deterministic():
    sys = v.System(_grid(3, 10), flux=ph.PowerDiffusion(1.0, 2.0), reaction=ph.SinhReaction(0.1), species=[1])
    st = v.SystemState(sys)
    try:
        U = _rand_u(sys)
        F1 = st.eval_res_jac(U).copy()
        A1 = st.matrix("csr").data.copy()
        F2 = st.eval_res_jac(U)
        A2 = st.matrix("csr").data
        assert np.array_equal(F1, F2) and np.array_equal(A1, A2)  # no atomics, fixed summation order
    finally:
        st.close()


# ---------------------------------------------------------------------------------------------- linear algebra (K8-K11)
def test_spmv_matches_scipy():
    sys = v.System(_grid(3, 9), flux=ph.CrossDiffusion2((1.0, 2.0), 0.1), reaction=ph.BilinearReaction2(0.5), species=[1, 2])
    st = v.SystemState(sys)
    try:
        st.eval_res_jac(_rand_u(sys))
        A = st.matrix("csr")
        x = np.random.default_rng(1).standard_normal(A.shape[1])
        y = st.spmv(x)
        np.testing.assert_allclose(y, A @ x, rtol=1e-13, atol=1e-13 * np.abs(A @ x).max())
    finally:
        st.close()


@pytest.mark.parametrize("method", ["default", "bicgstab_jacobi", "cg_jacobi", "bicgstab_block", "gmres_jacobi", "gmres_block", "cg_none", "bicgstab_ilu0", "cg_ilu0",
                                    "bicgstab_ilu0mc", "gmres_ilu0mc", "cg_amg", "bicgstab_amg", "gmres_amg"])
def test_newton_example301(method):
    """Example301: solution[43] known answer, and agreement with the oracle's direct solve"""
    X = np.linspace(0, 1, 6)
    sys = v.System(v.simplexgrid(X, X, X), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0))
    v.enable_species(sys, 1, [1])
    v.boundary_dirichlet(sys, 1, 5, 0.0)
    v.boundary_dirichlet(sys, 1, 6, 0.0)
    ml = {"default": None, "bicgstab_jacobi": v.KrylovJL_BICGSTAB(precs=v.JacobiPreconBuilder()), "cg_jacobi": v.KrylovJL_CG(precs=v.JacobiPreconBuilder()),
          "bicgstab_block": v.KrylovJL_BICGSTAB(precs=v.BlockPreconBuilder()), "gmres_jacobi": v.KrylovJL_GMRES(precs=v.JacobiPreconBuilder(), restart=40),
          "gmres_block": v.KrylovJL_GMRES(precs=v.BlockPreconBuilder(), restart=25), "cg_none": v.KrylovJL_CG(),
          "bicgstab_ilu0": v.KrylovJL_BICGSTAB(precs=v.ILUZeroPreconBuilder()), "cg_ilu0": v.KrylovJL_CG(precs=v.ILUZeroPreconBuilder()),
          "bicgstab_ilu0mc": v.KrylovJL_BICGSTAB(precs=v.ILUZeroPreconBuilder(multicolor=True)),
          "gmres_ilu0mc": v.KrylovJL_GMRES(precs=v.ILUZeroPreconBuilder(multicolor=True)), "cg_amg": v.KrylovJL_CG(precs=v.AMGPreconBuilder()),
          "bicgstab_amg": v.KrylovJL_BICGSTAB(precs=v.AMGPreconBuilder()), "gmres_amg": v.KrylovJL_GMRES(precs=v.SmoothedAggregationPreconBuilder())}[method]
    if method == "cg_none":  # without a preconditioner the 1e30 Dirichlet penalty makes Krylov hopeless: use Robin-free pure Neumann + reaction instead
        sys = v.System(v.simplexgrid(X, X, X), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0), reaction=ph.PowerReaction(1.0, 1.0))
        v.enable_species(sys, 1, [1])
        sol = v.solve(sys, inival=0.0, method_linear=ml, reltol_linear=1e-13, abstol_linear=0.0, maxiters_linear=2000)
        ref = O.OracleSystem(sys).solve_step(v.unknowns(sys))
        assert np.max(np.abs(sol - ref)) < TOL_NEWTON
        return
    sol = v.solve(sys, inival=0.0, method_linear=ml, reltol_linear=1e-13, abstol_linear=0.0, maxiters_linear=2000)
    ref = O.OracleSystem(sys).solve_step(v.unknowns(sys))
    assert np.max(np.abs(sol - ref)) < TOL_NEWTON
    assert sol.ravel(order="F")[42] == pytest.approx(0.012234524449380824, rel=1e-9)


def test_newton_example207_time_steps():
    """Example207: 100 implicit Euler steps with a reused device state; U[15] known answer"""
    X = np.linspace(0, 1, 11)
    sys = v.System(v.simplexgrid(X, X), flux=ph.PowerDiffusion(1.0e-2, 2), reaction=ph.PowerReaction(1.0, 2.0), source=ph.GaussSource(1, 20.0, (0.5, 0.5)),
                   storage=ph.LinearStorage(1.0))
    v.enable_species(sys, 1, [1])
    v.boundary_dirichlet(sys, 1, 2, 0.1)
    v.boundary_dirichlet(sys, 1, 4, 0.1)
    st = v.SystemState(sys)
    o = O.OracleSystem(sys)
    try:
        u = v.unknowns(sys, 0.5)
        uo = u.copy()
        t, tstep = 0.0, 0.01
        while t < 1.0:
            t += tstep
            u = v.solve(sys, state=st, inival=u, tstep=tstep)
            uo = o.solve_step(uo, tstep=tstep)
        assert np.max(np.abs(u - uo)) < TOL_NEWTON
        assert u.ravel(order="F")[14] == pytest.approx(0.3554284760906605, rel=1e-9)
    finally:
        st.close()


def test_transient_example160():
    """Example160 evolution through the device transient driver: evolval 18.721369939565655 (rtol 1e-5)"""
    n = 20
    X = np.arange(0, n + 1) / n
    eps, z, V = 1.0e-3, -1.0, 5.0
    bc = ph.BCondition().dirichlet(species=1, region=1, value=0.0, ramp=((0, 1.0e-2), (0, V))).dirichlet(species=1, region=2, value=0).dirichlet(species=2, region=2, value=0.5)
    R = np.zeros((2, 2))
    R[0, 1] = -2 * z
    sys = v.System(v.simplexgrid(X), flux=ph.SedanFlux(eps, z, 1, 2), reaction=ph.AffineReaction(R, [z, 0.0]), bcondition=bc, storage=ph.LinearStorage([0.0, 1.0]),
                   species=[1, 2])
    inival = v.unknowns(sys)
    inival[1, :] = 0.5
    tstep = 1.0e-5
    control = v.SolverControl(Δt_min=tstep, Δt=tstep, Δt_grow=1.1, Δt_max=0.1, Δu_opt=0.1, damp_initial=0.5)
    tsol = v.solve(sys, inival=inival, times=[0.0, 10], control=control)
    assert tsol.u[-1].sum() == pytest.approx(18.721369939565655, rel=1e-5)


def test_newton_many_species_410():
    """Example410 with 10 species: norm(sol) = sqrt(10 * 3.85)"""
    sys = v.System(v.simplexgrid(np.linspace(0, 1, 11)), flux=ph.LinearDiffusion())
    for i in range(1, 11):
        v.enable_species(sys, i, [1])
        v.boundary_dirichlet(sys, i, 1, 0)
        v.boundary_dirichlet(sys, i, 2, 1)
    sol = v.solve(sys, inival=0.0)
    assert np.linalg.norm(sol) == pytest.approx(math.sqrt(10 * 3.85), rel=1e-10)


def test_unregistered_physics_id_is_an_error():
    import ctypes as C

    sys = v.System(_grid(2, 5), flux=ph.LinearDiffusion(), species=[1])
    st = v.SystemState(sys)
    try:
        p = np.zeros(1)
        rc = st.L.vfvm_set_physics(st.h, 0, 99, v._lib.dptr(p), 1)
        assert rc == v._lib.ERR_UNREGISTERED
    finally:
        st.close()


@pytest.mark.parametrize("ns", [1, 2])
def test_ilu0_is_exact_on_block_tridiagonal(ns):
    """on a 1D grid the node-block matrix is block tridiagonal, ILU(0) is the exact LU: the Krylov solver must converge in one step"""
    import ctypes as C

    X = np.linspace(0, 1, 200)
    if ns == 1:
        sys = v.System(v.simplexgrid(X), flux=ph.PowerDiffusion(1.0, 2), reaction=ph.SinhReaction(0.3), species=[1])
    else:
        sys = v.System(v.simplexgrid(X), flux=ph.CrossDiffusion2((1.0, 0.5), 0.01), reaction=ph.BilinearReaction2(1.0), storage=ph.LinearStorage(1.0), species=[1, 2])
    v.boundary_dirichlet(sys, 1, 1, 1.0)
    v.boundary_dirichlet(sys, ns, 2, 0.5)
    for precon in (v._lib.PRECON_ILU0, v._lib.PRECON_ILU0_MC):
        st = v.SystemState(sys)
        try:
            U = _rand_u(sys)
            st.set_vector(v._lib.VEC_SOLUTION, U)
            st.set_vector(v._lib.VEC_OLDSOL, U)
            assert st.assemble(tstep=0.1) == 0
            A = st.matrix("csr")
            F = st.get_vector(v._lib.VEC_RESIDUAL)
            v._lib.check(st.h, st.L.vfvm_linsolve_setup(st.h, v._lib.KRYLOV_BICGSTAB, precon, 0))
            it, rn = C.c_int(), C.c_double()
            v._lib.check(st.h, st.L.vfvm_linsolve(st.h, 0.0, 1e-12, 50, 0, C.byref(it), C.byref(rn)))
            x = st.get_vector(v._lib.VEC_UPDATE).ravel(order="F")
            r = A @ x - F.ravel(order="F")
            assert np.linalg.norm(r) <= 1e-10 * np.linalg.norm(F)
            if precon == v._lib.PRECON_ILU0:
                assert it.value <= 2, f"natural-order ILU(0) of a block tridiagonal matrix is exact, got {it.value} iterations"
        finally:
            st.close()


def test_newton_example207_cg_ilu0():
    """examples/Example207_NonlinearPoisson2D.jl:86: KrylovJL_CG(precs = ILUZeroPreconBuilder()) must reproduce U[15]"""
    X = np.linspace(0, 1, 11)
    sys = v.System(v.simplexgrid(X, X), flux=ph.PowerDiffusion(1.0e-2, 2), reaction=ph.PowerReaction(1.0, 2.0), source=ph.GaussSource(1, 20.0, (0.5, 0.5)),
                   storage=ph.LinearStorage(1.0))
    v.enable_species(sys, 1, [1])
    v.boundary_dirichlet(sys, 1, 2, 0.1)
    v.boundary_dirichlet(sys, 1, 4, 0.1)
    st = v.SystemState(sys)
    try:
        control = v.SolverControl(reltol_linear=1.0e-5, method_linear=v.KrylovJL_CG(precs=v.ILUZeroPreconBuilder()))
        u = v.unknowns(sys, 0.5)
        t, tstep = 0.0, 0.01
        while t < 1.0:
            t += tstep
            u = v.solve(sys, state=st, inival=u, control=control, tstep=tstep)
        assert u.ravel(order="F")[14] == pytest.approx(0.3554284760906605, rel=1.5e-8)  # the reference's isapprox default
    finally:
        st.close()


def test_ilu0_beats_jacobi_in_iterations():
    import ctypes as C

    sys = v.System(_grid(3, 17), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0), species=[1])
    v.boundary_dirichlet(sys, 1, 5, 0.0)
    v.boundary_dirichlet(sys, 1, 6, 0.0)
    st = v.SystemState(sys)
    try:
        st.set_vector(v._lib.VEC_SOLUTION, v.unknowns(sys, 0.0))
        st.set_vector(v._lib.VEC_OLDSOL, v.unknowns(sys, 0.0))
        assert st.assemble() == 0
        its = {}
        for name, pc in (("jacobi", v._lib.PRECON_JACOBI), ("ilu0", v._lib.PRECON_ILU0), ("ilu0mc", v._lib.PRECON_ILU0_MC)):
            v._lib.check(st.h, st.L.vfvm_linsolve_setup(st.h, v._lib.KRYLOV_CG, pc, 0))
            it, rn = C.c_int(), C.c_double()
            v._lib.check(st.h, st.L.vfvm_linsolve(st.h, 0.0, 1e-10, 2000, 0, C.byref(it), C.byref(rn)))
            its[name] = it.value
        assert its["ilu0"] < its["ilu0mc"] <= its["jacobi"], its
    finally:
        st.close()


# ---------------------------------------------------------------------------------------------- reference known answers on the device
def test_device_bernoulli_accuracy():
    """test/test010_bernoulli.jl:5-22 on the device function: |B(x) - x/(exp(x)-1)| < 1e-14 on both ranges, both halves of fbernoulli_pm"""
    import ctypes as C

    import mpmath

    mpmath.mp.prec = 256

    def big(x):
        bx = mpmath.mpf(float(x))
        return float(bx / (mpmath.exp(bx) - 1)) if x != 0 else 1.0

    h = C.c_void_p()
    L = v._lib.lib()
    assert L.vfvm_create(0, C.byref(h)) == 0
    try:
        for rng in (np.arange(-1, 1, 1.00001e-5)[::7], np.arange(-100, 100, 1.00001e-3)[::11]):
            x = np.ascontiguousarray(rng)
            bp, bm, dbp = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
            v._lib.check(h, L.vfvm_probe_bernoulli(h, x.size, v._lib.dptr(x), v._lib.dptr(bp), v._lib.dptr(bm), v._lib.dptr(dbp)))
            ref = np.array([big(t) for t in x])
            refm = np.array([big(-t) for t in x])
            assert np.max(np.abs(bp - ref)) < 1.0e-14
            assert np.max(np.abs(bm - refm)) < 1.0e-14
            obp, odbp, obm, odbm = O.fbernoulli_dual(x)
            np.testing.assert_allclose(dbp, odbp, rtol=1e-13, atol=1e-15)  # derivative through the dual-number path vs the oracle's
    finally:
        L.vfvm_destroy(h)


def test_example105_device():
    """examples/Example105_NonlinearPoisson1D.jl: sum(solution) == 1.5247901344230088"""
    sys = v.System(v.simplexgrid(np.arange(0, 11) / 10.0), flux=ph.LinearDiffusion(1.0e-3), source=ph.Step1DSource(1, 0.5, 1.0, -1.0), reaction=ph.SinhReaction())
    v.enable_species(sys, 1, [1])
    v.boundary_dirichlet(sys, 1, 1, 0.0)
    v.boundary_dirichlet(sys, 1, 2, 1.0)
    sol = v.solve(sys, inival=0.5)
    assert sol.sum() == pytest.approx(1.5247901344230088, rel=1e-10)


def test_example106_device_transient():
    """examples/Example106_NonlinearDiffusion1D.jl: fixed-step implicit Euler, sum(tsol.u[end]) == 46.66666666647518"""
    n, m, tend, tstep = 20, 2, 0.01, 0.0001
    h = 1.0 / (n / 2)
    X = np.arange(-1, 1 + h / 2, h)
    sys = v.System(v.simplexgrid(X), flux=ph.PowerDiffusion(1.0, m), storage=ph.LinearStorage(1.0))
    v.enable_species(sys, 1, [1])
    inival = v.unknowns(sys)
    t0 = 0.001
    tx = t0 ** (-1.0 / (m + 1.0))
    xx = np.maximum(1 - (X * tx) ** 2 * (m - 1) / (2.0 * m * (m + 1)), 0.0)
    inival[0, :] = tx * xx ** (1.0 / (m - 1.0))
    control = v.SolverControl(Δt_min=tstep, Δt_max=tstep, Δt=tstep, Δu_opt=1)
    tsol = v.solve(sys, inival=inival, times=[t0, tend], control=control)
    assert tsol.u[-1].sum() == pytest.approx(46.66666666647518, rel=1e-9)


def test_example110_device_parameter_continuation():
    """examples/Example110: seven solves with changing diffusion coefficients on ONE device state (parameters are re-uploaded,
    the pattern is kept); U[5] == 0.7117546972922056"""
    sys = v.System(v.simplexgrid(np.arange(0, 101) / 100.0), reaction=ph.BilinearReaction2(1.0), flux=ph.CrossDiffusion2((1.0, 1.0), 0.01),
                   source=ph.AffineXSource([1.0e-4 * 0.01, 1.0e-4 * 1.01], [1.0e-4, -1.0e-4]), storage=ph.LinearStorage(1.0))
    v.enable_species(sys, 1, [1])
    v.enable_species(sys, 2, [1])
    for sp in (1, 2):
        v.boundary_dirichlet(sys, sp, 1, 1.0)
        v.boundary_dirichlet(sys, sp, 2, 0.0)
    st = v.SystemState(sys)
    try:
        U = v.unknowns(sys, 0.0)
        control = v.SolverControl(damp_initial=0.1)
        for xeps in [1.0, 0.5, 0.25, 0.1, 0.05, 0.025, 0.01]:
            sys.physics.slots[0].eps = (xeps, xeps)
            sys._version += 1
            U = v.solve(sys, state=st, inival=U, control=control)
        assert U.ravel(order="F")[4] == pytest.approx(0.7117546972922056, rel=1e-8)
    finally:
        st.close()


def test_example210_device_transient_two_species():
    """examples/Example210_NonlinearPoisson2D_Reaction.jl: sum(tsol.u[end]) == 16.01812472041518"""
    X = np.linspace(0, 1, 11)
    k, eps = 1.0, 1.0e-2
    sys = v.System(v.simplexgrid(X, X), flux=ph.LinearDiffusion(eps), storage=ph.LinearStorage(1.0), reaction=ph.AffineReaction([[k, -k], [-k, k]]),
                   source=ph.GaussSource(1, 20.0, (0.5, 0.5)))
    v.enable_species(sys, 1, [1])
    v.enable_species(sys, 2, [1])
    tstep = 0.01
    control = v.SolverControl(Δt=tstep, Δt_min=tstep, Δt_max=tstep, Δu_opt=1.0e5, method_linear=v.KrylovJL_BICGSTAB(precs=v.ILUZeroPreconBuilder()),
                              reltol_linear=1e-12, abstol_linear=0.0, maxiters_linear=500, factorize_every_timestep=2)
    tsol = v.solve(sys, inival=v.unknowns(sys, 0.0), times=(0, 1), control=control)
    assert tsol.u[-1].sum() == pytest.approx(16.01812472041518, rel=1e-8)


def test_example215_device_boundary_reaction():
    """examples/Example215_NonlinearPoisson2D_BoundaryReaction.jl: U[25] == 0.2760603343272377"""
    X = np.arange(0, 11) / 10.0
    g = v.simplexgrid(X, X)
    k = 1.0
    sys = v.System(g, breaction=ph.LinearBoundaryReaction(2, [[k, -k], [-k, k]]), flux=ph.LinearDiffusion(1.0e-2), storage=ph.LinearStorage(1.0))
    v.enable_species(sys, 1, [1])
    v.enable_species(sys, 2, [1])
    inival = v.unknowns(sys)
    inival[0, :] = np.exp(-5.0 * ((g.coord[0] - 0.5) ** 2 + (g.coord[1] - 0.5) ** 2))
    st = v.SystemState(sys)
    try:
        tstep, time, u25 = 0.01, 0.0, 0.0
        while time < 100:
            time += tstep
            U = v.solve(sys, state=st, inival=inival, tstep=tstep)
            inival = U
            tstep *= 1.2
            u25 = U.ravel(order="F")[24]
        assert u25 == pytest.approx(0.2760603343272377, rel=1e-8)
    finally:
        st.close()


def test_amg_multilevel_matches_direct_solve_and_beats_jacobi():
    """aggregation AMG (csrc/amg.cu) on a problem with several coarse levels: Newton solution = oracle's direct solve to 1e-10,
    and far fewer Krylov iterations than point Jacobi"""
    X = np.linspace(0, 1, 33)
    sys = v.System(v.simplexgrid(X, X, X), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0))
    v.enable_species(sys, 1, [1])
    v.boundary_dirichlet(sys, 1, 5, 0.0)
    v.boundary_dirichlet(sys, 1, 6, 0.0)
    ref = O.OracleSystem(sys).solve_step(v.unknowns(sys))
    its = {}
    for name, pc in (("jacobi", v.JacobiPreconBuilder()), ("amg", v.AMGPreconBuilder()), ("amg_w", v.AMGPreconBuilder(wdepth=2, alpha=2.0))):
        st = v.SystemState(sys)
        try:
            sol = v.solve(sys, state=st, inival=0.0, method_linear=v.KrylovJL_CG(precs=pc), reltol_linear=1e-13, abstol_linear=0.0, maxiters_linear=3000)
            its[name] = st.history.nlin
        finally:
            st.close()
        assert np.max(np.abs(sol - ref)) < TOL_NEWTON, name
    assert 0 < its["amg"] * 3 < its["jacobi"], its
    assert 0 < its["amg_w"] <= its["amg"], its


def test_amg_block_system_bipolar_newton():
    """AMG on a 3-species block system with non-symmetric coupled blocks (bipolar drift-diffusion, BiCGStab): same Newton
    iterates as BiCGStab + block-Jacobi"""
    g = _grid(3, 14)
    bc = ph.BCondition()
    for sp, val in ((1, 0.0), (2, 0.0), (3, 0.5)):
        bc.dirichlet(species=sp, region=5, value=val)
    for sp, val in ((1, 0.1), (2, 0.1), (3, 0.2)):
        bc.dirichlet(species=sp, region=6, value=val)
    sys = v.System(g, flux=ph.BipolarSGFlux(), reaction=ph.BipolarReaction([1.0]), storage=ph.BipolarStorage(), bcondition=bc, species=[1, 2, 3])
    sols = []
    for pc in (v.BlockPreconBuilder(), v.AMGPreconBuilder()):
        sols.append(v.solve(sys, inival=0.1, tstep=1.0e-2, method_linear=v.KrylovJL_BICGSTAB(precs=pc), reltol_linear=1e-13, abstol_linear=0.0, maxiters_linear=3000))
    assert np.max(np.abs(sols[0] - sols[1])) < TOL_NEWTON


# ---- species enabled per cell region (enable_species!(sys, i, regions)), inactive dofs ---------------------------------------
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_example221_species_per_region_device(dim):
    """examples/Example221_EquationBlockPrecon.jl: species 1 in region 1, species 2 everywhere, species 3 in region 3.
    Pattern bit-exact and values 1e-12 against the oracle (which reproduces the reference's known answers), Newton solution 1e-10,
    inactive dofs identically zero"""
    from test_oracle_golden import example221_system

    s = example221_system(dim)
    U = _rand_u(s)
    U[~s.node_dof()] = 0.0
    _compare_assembly(s, U)
    _compare_assembly(s, U, _rand_u(s, seed=4) * s.node_dof(), tstep=0.1)
    ref = O.OracleSystem(s).solve_step(v.unknowns(s, inival=0.0))
    for ml in (None, v.KrylovJL_BICGSTAB(precs=v.BlockPreconBuilder()), v.KrylovJL_BICGSTAB(precs=v.AMGPreconBuilder())):
        sol = v.solve(s, inival=0.0, method_linear=ml, reltol_linear=1e-13, abstol_linear=0.0, maxiters_linear=5000)
        assert np.max(np.abs(sol - ref)) < TOL_NEWTON
        assert np.all(sol[~s.node_dof()] == 0.0)
    expected = {1: 0.014101758266210086, 2: 0.12691582439590407, 3: 1.1422561017685693}[dim]
    assert sol[1].sum() == pytest.approx(expected, rel=1e-9 if dim < 3 else 5e-5)


def test_masked_coupled_flux_two_species():
    """cross-diffusion flux (full 2x2 coupling) with species 2 restricted to one of two cell regions: couplings exist only where
    both species share a region"""
    g = _grid(2, 13)
    v.cellmask(g, [0.0, 0.0], [0.5, 1.0], 2)
    s = v.System(g, flux=ph.CrossDiffusion2([1.0, 0.5], 0.1), reaction=ph.BilinearReaction2(0.3), storage=ph.LinearStorage([1.0, 2.0]))
    v.enable_species(s, 1, [1, 2])
    v.enable_species(s, 2, [2])
    v.boundary_dirichlet(s, 1, 2, 1.0)
    v.boundary_dirichlet(s, 2, 4, 0.5)
    U = _rand_u(s)
    U[~s.node_dof()] = 0.0
    _compare_assembly(s, U, _rand_u(s, seed=11) * s.node_dof(), tstep=0.05)


def test_enable_species_after_state_creation():
    """enable_species! on a system whose SystemState already exists: the device twin takes the new region masks at the next sync (new
    pattern, all physics pushed again) and assembles what a fresh state assembles; transient callbacks receive the solution arrays"""
    g = _grid(2, 13)
    v.cellmask(g, [0.0, 0.0], [0.5, 1.0], 2)
    s = v.System(g, flux=ph.LinearDiffusion([1.0, 0.5]), reaction=ph.AffineReaction([[1.0, -0.2], [-0.1, 2.0]]), storage=ph.LinearStorage([1.0, 2.0]))
    v.enable_species(s, 1, [1, 2])
    v.enable_species(s, 2, [2])
    v.boundary_dirichlet(s, 1, 2, 1.0)
    st = v.SystemState(s)
    try:
        U = _rand_u(s)
        st.eval_res_jac(U * s.node_dof(), None, tstep=0.05)
        v.enable_species(s, 2, [1])  # species 2 everywhere now
        st.sync()
        U2 = _rand_u(s, seed=3)
        F = st.eval_res_jac(U2, None, tstep=0.05)
        A = st.matrix("csc")
        Fo, Ao = O.OracleSystem(s).assemble(U2, None, tstep=0.05)
        assert np.array_equal(A.indptr, Ao.indptr) and np.array_equal(A.indices, Ao.indices)
        np.testing.assert_allclose(A.data, Ao.data, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(F, Fo, rtol=1e-12, atol=1e-13)
        seen = []
        tsol = v.solve_state(st, inival=0.5, times=[0.0, 0.1], Δt=0.05, Δt_min=0.05, Δt_max=0.05, Δu_opt=1.0e5,
                             pre=lambda sol, t: seen.append(("pre", None if sol is None else sol.shape, t)),
                             post=lambda sol, oldsol, t, dt: seen.append(("post", sol.shape, oldsol.shape, float(np.abs(sol - oldsol).max()) > 0.0)),
                             sample=lambda sol, t: seen.append(("sample", sol.shape, t)))
        assert ("pre", (2, g.num_nodes), 0.05) in seen and any(e[0] == "sample" and e[1] == (2, g.num_nodes) for e in seen)
        assert ("post", (2, g.num_nodes), (2, g.num_nodes), True) in seen
        assert len(tsol.t) >= 2
    finally:
        st.close()


def test_sparse_unknown_storage_solve():
    """unknown_storage = :sparse changes the host container only (test/ of the reference runs every example with both storages and
    expects the same numbers): solve() hands back a SparseSolutionArray with the dense solve's values at the defined dofs"""
    from vfvm_b200.sparsesolution import SparseSolutionArray

    sols = {}
    for storage in ("dense", "sparse"):
        g = _grid(2, 13)
        v.cellmask(g, [0.0, 0.0], [0.5, 1.0], 2)
        s = v.System(g, flux=ph.LinearDiffusion([1.0, 0.5]), reaction=ph.AffineReaction([[1.0, -0.2], [-0.1, 2.0]]), unknown_storage=storage)
        v.enable_species(s, 1, [1, 2])
        v.enable_species(s, 2, [2])
        v.boundary_dirichlet(s, 1, 2, 1.0)
        v.boundary_dirichlet(s, 2, 4, 0.5)
        sols[storage] = (v.solve(s, inival=v.unknowns(s, inival=0.1)), s.node_dof())
    sp, mask = sols["sparse"]
    assert isinstance(sp, SparseSolutionArray) and isinstance(sols["dense"][0], np.ndarray)
    assert len(sp) == int(mask.sum())
    # (the two solves start from different values at the undefined dofs -- 0.1 in the dense array, nothing in the sparse one -- so their Krylov
    # right-hand sides differ there and the defined dofs agree to solver accuracy, not bitwise)
    np.testing.assert_allclose(sp.dense()[mask], sols["dense"][0][mask], rtol=0, atol=1e-11)
    assert np.all(sp.dense()[~mask] == 0.0)
    assert sp.history is not None


# ---- boundary species, bstorage, edgereaction ---------------------------------------------------------------------
@pytest.mark.parametrize("storage", ["dense", "sparse"])
def test_state_reuse_and_restart_test060(storage):
    """test/test060_state.jl:16-44: a solve through a reused SystemState equals the solve that builds its own; a transient run restarted
    from the end of a previous one equals the run over both intervals, with and without a shared state, for both unknown storages"""
    def dense(u):
        return u.dense() if hasattr(u, "dense") else u

    bc = ph.BCondition().robin(species=1, region=1, value=0.0, factor=0.1).dirichlet(species=1, region=2, value=1.0)
    s = v.System(v.simplexgrid(np.arange(0, 1.05, 0.1)), flux=ph.LinearDiffusion(1.0), bcondition=bc, species=[1], unknown_storage=storage)
    sol1 = dense(v.solve(s))
    st = v.SystemState(s)
    try:
        sol2 = dense(v.solve_state(st))
        np.testing.assert_allclose(sol1, sol2, rtol=1.5e-8)
        control = v.SolverControl()
        v.fixed_timesteps(control, 0.025)
        tsol1 = v.solve(s, inival=0.0, times=(0, 0.1), control=control)
        tsol2 = v.solve(s, inival=tsol1.u[-1], times=(0.1, 0.2), control=control)
        tsol3 = v.solve(s, inival=0.0, times=[0.0, 0.1, 0.2], control=control)
        np.testing.assert_allclose(tsol3.u[-1], tsol2.u[-1], rtol=1.5e-8)
        xsol1 = v.solve_state(st, inival=0.0, times=(0, 0.1), control=control)
        xsol2 = v.solve_state(st, inival=tsol1.u[-1], times=(0.1, 0.2), control=control)
        xsol3 = v.solve_state(st, inival=0.0, times=[0.0, 0.1, 0.2], control=control)
        np.testing.assert_allclose(xsol3.u[-1], xsol2.u[-1], rtol=1.5e-8)
        np.testing.assert_allclose(xsol3.u[-1], tsol3.u[-1], rtol=1.5e-8)
        assert len(xsol1.t) == 5 and xsol3.t[-1] == pytest.approx(0.2)
    finally:
        st.close()


def test_example220_boundary_species_2d_device():
    """examples/Example220_NonlinearPoisson2D_BoundarySpecies.jl:65-101 on the device: two bulk species, one species on boundary region 2, linear
    exchange boundary reaction, bstorage; 100 implicit Euler steps, U_bound[5] == 0.0020781361856598"""
    from test_oracle_golden import _example220_system

    s, bnodes = _example220_system()
    st = v.SystemState(s)
    try:
        U = v.unknowns(s)
        for _ in range(100):
            U = v.solve_state(st, inival=U, tstep=0.01, reltol=1.0e-5)
        assert U[2, bnodes[4]] == pytest.approx(0.0020781361856598, rel=1e-8)
        assert np.all(U[2][~s.node_dof()[2]] == 0.0)
    finally:
        st.close()


@pytest.mark.parametrize("switchbc", [False, True])
def test_example115_boundary_species_bstorage_device(switchbc):
    """examples/Example115_HeterogeneousCatalysis1D.jl: surface species C (enable_boundary_species!) with bstorage and the nonlinear
    catalysis breaction; assembly against the oracle (pattern bit-exact, values 1e-12) and the known answer
    tsol[iC, inodeCat, end] == 0.87544440641274 after 100 implicit Euler steps on the device"""
    from test_oracle_golden import example115_system

    sys, inode = example115_system(switchbc)
    U = _rand_u(sys)
    U[~sys.node_dof()] = 0.0
    Uold = _rand_u(sys, seed=4) * sys.node_dof()
    _compare_assembly(sys, U, Uold, tstep=0.01)
    control = v.fixed_timesteps(v.SolverControl(), 0.01)
    tsol = v.solve(sys, inival=0.0, times=[0.0, 1.0], control=control)
    assert tsol.u[-1][2, inode] == pytest.approx(0.87544440641274, rel=1e-10)
    assert np.all(tsol.u[-1][2, np.arange(11) != inode] == 0.0)


def test_boundary_species_2d_parity():
    """a surface species on one boundary region of a 2D grid, coupled to two bulk species: pattern bit-exact, values 1e-12, Newton 1e-10"""
    X = np.linspace(0, 1, 9)
    g = v.simplexgrid(X, X)
    sys = v.System(g, flux=ph.LinearDiffusion([1.0, 1.0e-1, 0.0]), storage=ph.LinearStorage([1.0, 1.0, 0.0]), source=ph.GaussSource(1, 20.0, (0.5, 0.5)),
                   breaction=ph.CatalysisBoundaryReaction(1, S=0.05, kp_AC=10.0, km_AC=1.0, kp_BC=0.5, km_BC=1.0), bstorage=ph.LinearBoundaryStorage(1, [0.0, 0.0, 1.0]))
    v.enable_species(sys, 1, [1])
    v.enable_species(sys, 2, [1])
    v.enable_boundary_species(sys, 3, [1])
    v.boundary_dirichlet(sys, 2, 3, 0.0)
    U = _rand_u(sys)
    U[~sys.node_dof()] = 0.0
    Uold = _rand_u(sys, seed=5) * sys.node_dof()
    _compare_assembly(sys, U, Uold, tstep=0.05)
    ref = O.OracleSystem(sys).solve_step(v.unknowns(sys, 0.1) * sys.node_dof(), tstep=0.05)
    sol = v.solve(sys, inival=v.unknowns(sys, 0.1) * sys.node_dof(), tstep=0.05)
    assert np.max(np.abs(sol - ref)) < TOL_NEWTON


def test_devex002_edge_reaction_device():
    """examples/DevEx002_EdgeReaction.jl (3D case): node reaction and edge reaction (times the half diamond volume) give the same solution;
    the edge-reaction assembly agrees with the oracle"""
    X = np.linspace(0, 1, 7)
    g = v.simplexgrid(X, X, X)
    bc = ph.BCondition()
    for r in range(1, 7):
        bc.dirichlet(species=1, region=r, value=0.0)
    s_node = v.System(g, flux=ph.LinearDiffusion(), reaction=ph.AffineReaction([[0.0]], [-1.0]), storage=ph.LinearStorage(1.0), bcondition=bc, species=[1], is_linear=True)
    s_edge = v.System(g, flux=ph.LinearDiffusion(), edgereaction=ph.DiamondEdgeReaction(-1.0), storage=ph.LinearStorage(1.0), bcondition=bc, species=[1], is_linear=True)
    _compare_assembly(s_edge, _rand_u(s_edge))
    u_node = v.solve(s_node, inival=0.0)
    u_edge = v.solve(s_edge, inival=0.0)
    assert np.abs(u_node).max() > 1e-3
    assert np.abs(u_node - u_edge).max() <= 1e-10 * np.abs(u_node).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_joule_heat_edge_reaction_parity(dim):
    """Example206 physics (potential + temperature, Joule heat as an edge reaction that depends on the solution) on a tensor grid:
    residual and Jacobian (with the reference's own sign convention, src/vfvm_assembly.jl:213-230) against the oracle, Newton 1e-10"""
    g = _grid(dim, 9 if dim == 2 else 6)
    kappa = math.exp(-0.5)
    bc = ph.BCondition()
    bc.dirichlet(species=1, region=2 if dim == 2 else 5, value=-1.0).dirichlet(species=1, region=4 if dim == 2 else 6, value=1.0)
    for r in range(1, 2 * dim + 1):
        bc.robin(species=2, region=r, factor=0.5, value=0.5)
    sys = v.System(g, flux=ph.LinearDiffusion([kappa, 1.0]), edgereaction=ph.JouleHeatEdgeReaction(kappa, 1, 2), storage=ph.LinearStorage([0.0, 1.0]), bcondition=bc, species=[1, 2])
    _compare_assembly(sys, _rand_u(sys))
    _compare_assembly(sys, _rand_u(sys), _rand_u(sys, seed=3), tstep=0.1)
    ref = O.OracleSystem(sys).solve_step(v.unknowns(sys, 0.0))
    sol = v.solve(sys, inival=0.0)
    assert np.max(np.abs(sol - ref)) < TOL_NEWTON


# ---- small dense LU inside a callback (inplace_linsolve!), DevEx005 mixture flux ----------------------------------------
def test_inplace_linsolve_device_probe():
    """test/test040_inplacelu.jl:16-38 on the device: A = -rand + 100 I, x = 1, b = A x, error < 100 eps for N = 2..10 (non-pivoting Doolittle and
    pivoting LU); both agree with the oracle's restatement to rounding"""
    import ctypes as C

    sys = v.System(_grid(1, 5), flux=ph.LinearDiffusion(), species=[1])
    st = v.SystemState(sys)
    rng = np.random.default_rng(40)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    try:
        for n in range(2, 11):
            A = np.ascontiguousarray(-rng.uniform(size=(64, n, n)) + 100.0 * np.eye(n))
            b = np.ascontiguousarray(A.sum(axis=2))
            for piv in (0, 1):
                x, xo = np.zeros((64, n)), np.zeros((64, n))
                v._lib.check(st.h, st.L.vfvm_probe_inplace_linsolve(st.h, n, 64, piv, dp(A), dp(b), dp(x)))
                assert np.all(np.sqrt(((x - 1.0) ** 2).sum(axis=1)) / n < 100 * np.finfo(float).eps)
                assert O.lib().vo_probe_inplace_linsolve(n, 64, piv, dp(A), dp(b), dp(xo)) == 0
                assert np.max(np.abs(x - xo)) < 1e-14
    finally:
        st.close()


@pytest.mark.parametrize("dim,expected", [(1, 4.788926530387466), (2, 15.883072449873742), (3, 52.67819183426213)])
def test_devex005_mixture_device(dim, expected):
    """examples/DevEx005_Mixture.jl: five-species Maxwell-Stefan flux with a 5 x 5 pivoting LU inside the flux callback, evaluated in Dual<10>
    on the device: assembly against the oracle (1e-12) and the reference's known answers norm(u) (atol 1e-5)"""
    from test_oracle_golden import devex005_system

    sys = devex005_system(dim)
    _compare_assembly(sys, _rand_u(sys))
    u = v.solve(sys, inival=0.0, damp_initial=0.5, tol_mono=1.0e-10, tol_round=1.0e-15, max_round=3, maxiters=500)
    assert np.linalg.norm(u) == pytest.approx(expected, abs=1.0e-5)
