"""Post-processing integrals on the device: the reference's known answers of test/test120_norms.jl
and entry-by-entry agreement with the CPU oracle's restatement of integrate / edgeintegrate."""
import math

import numpy as np
import pytest

import vfvm_b200 as v
from vfvm_b200 import physics as ph
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _grid(dim, h=0.1):
    X = np.linspace(0, 1, int(round(1 / h)) + 1)
    return v.simplexgrid(*([X] * dim))


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("c", [0.5, 1.0, 2.0])
def test_norms_known_answers(dim, c):
    """test/test120_norms.jl:103-133: int c = c, edge integral of c = c, l2norm(c) = c, h1seminorm(c) = 0, h1seminorm(c sum(x)/sqrt(dim)) = c,
    sum of the node volumes = 1"""
    g = _grid(dim)
    sys = v.System(g, species=[1])
    st = v.SystemState(sys)
    try:
        F = np.full((1, g.num_nodes), c)
        assert v.integrate(sys, F, state=st)[0, 0] == pytest.approx(c, rel=1e-12)                               # test_solint
        assert v.edgeintegrate(sys, v.postprocess.EdgeAverage(), F, state=st)[0, 0] == pytest.approx(c, rel=1e-12)  # test_edgeint: f = (u_K + u_L) / 2
        assert v.l2norm(sys, F, state=st) == pytest.approx(c, rel=1e-12)
        assert v.h1seminorm(sys, F, state=st) == pytest.approx(0.0, abs=1e-12)
        lin = (c * g.coord.sum(axis=0) / math.sqrt(dim))[None, :]
        assert v.h1seminorm(sys, lin, state=st) == pytest.approx(c, rel=1e-10)
        assert v.nodevolumes(sys, state=st).sum() == pytest.approx(1.0, rel=1e-12)
    finally:
        st.close()


def test_integrals_match_the_oracle_multiregion_multispecies():
    X = np.linspace(0, 1, 13)
    g = v.simplexgrid(X, X, X)
    v.cellmask(g, [0, 0, 0.3], [1, 1, 0.72], 2)
    v.cellmask(g, [0, 0, 0.7], [1, 1, 1.0], 3)
    sys = v.System(g, flux=ph.BipolarSGFlux(), reaction=ph.BipolarReaction([10.0, 0.0, -10.0]), storage=ph.BipolarStorage(), species=[1, 2, 3])
    rng = np.random.default_rng(3)
    U = np.asfortranarray(rng.uniform(-0.5, 0.5, (3, g.num_nodes)))
    o = O.OracleSystem(sys)
    st = v.SystemState(sys)
    try:
        for F in (None, sys.physics.reaction, sys.physics.storage, ph.PowerReaction(1.0, 2.0)):
            dev = v.integrate(sys, U, state=st) if F is None else v.integrate(sys, F, U, state=st)
            ref = o.integrate(U) if F is None else o.integrate(U, F.slot, F.id, F.params(3))
            assert dev.shape == (3, 3)
            assert np.all(np.abs(dev - ref) <= 1e-12 * np.abs(ref) + 1e-14), (F, dev, ref)
        for F in (sys.physics.flux, ph.LinearDiffusion([1.0, 2.0, 3.0]), v.postprocess.W1pIntegrand(2.0), v.postprocess.EdgeAverage()):
            dev = v.edgeintegrate(sys, F, U, state=st)
            ref = o.edgeintegrate(U, F.id, F.params(3))
            assert np.all(np.abs(dev - ref) <= 1e-12 * np.abs(ref) + 1e-13), (F, dev, ref)
    finally:
        st.close()


def test_integrals_respect_species_enabled_per_region():
    """enable_species!(sys, i, regions): a species contributes to integral[i, region] only where it is enabled
    (assemble_res -> isregionspecies, src/vfvm_assemblydata.jl:310-332, 385-405); affine node function with F(0) != 0"""
    X = np.linspace(0, 1, 11)
    g = v.simplexgrid(X, X, X)
    v.cellmask(g, [0, 0, 0.3], [1, 1, 0.7], 2)
    v.cellmask(g, [0, 0, 0.7], [1, 1, 1.0], 3)
    R = np.array([[1.0, 0.2, 0.0], [0.1, 2.0, 0.3], [0.0, 0.4, 3.0]])
    sys = v.System(g, flux=ph.LinearDiffusion([1.0, 2.0, 3.0]), reaction=ph.AffineReaction(R, [0.5, -0.25, 1.5]))
    v.enable_species(sys, 1, [1])
    v.enable_species(sys, 2, [1, 2, 3])
    v.enable_species(sys, 3, [3])
    U = np.asfortranarray(np.random.default_rng(5).uniform(0.1, 1.0, (3, g.num_nodes)))
    o = O.OracleSystem(sys)
    st = v.SystemState(sys)
    try:
        for F in (None, sys.physics.reaction):
            dev = v.integrate(sys, U, state=st) if F is None else v.integrate(sys, F, U, state=st)
            ref = o.integrate(U) if F is None else o.integrate(U, F.slot, F.id, F.params(3))
            assert np.all(np.abs(dev - ref) <= 1e-12 * np.abs(ref) + 1e-14), (F, dev, ref)
            assert dev[0, 1] == 0.0 and dev[0, 2] == 0.0 and dev[2, 0] == 0.0 and dev[2, 1] == 0.0  # disabled (species, region) pairs stay empty
        for F in (sys.physics.flux, v.postprocess.W1pIntegrand(2.0), v.postprocess.EdgeAverage()):
            dev = v.edgeintegrate(sys, F, U, state=st)
            ref = o.edgeintegrate(U, F.id, F.params(3))
            assert np.all(np.abs(dev - ref) <= 1e-12 * np.abs(ref) + 1e-13), (F, dev, ref)
            assert dev[0, 1] == 0.0 and dev[2, 0] == 0.0
    finally:
        st.close()


def test_integrate_rejects_unregistered_functions():
    sys = v.System(_grid(2), species=[1])
    with pytest.raises(v.UnregisteredPhysicsError):
        v.integrate(sys, lambda y, u, node, data=None: None, np.zeros((1, sys.grid.num_nodes)))


def test_ode_interface_entry_points():
    """eval_rhs! / eval_jacobian! / mass_matrix (src/vfvm_diffeq_interface.jl:27-101): -F, -J of the stationary operator and the
    storage Jacobian at 0 times the node volumes, against the oracle"""
    X = np.linspace(0, 1, 9)
    g = v.simplexgrid(X, X)
    v.cellmask(g, [0.0, 0.0], [0.5, 1.0], 2)
    s = v.System(g, flux=ph.CrossDiffusion2([1.0, 0.5], 0.1), reaction=ph.BilinearReaction2(0.3), storage=ph.LinearStorage([1.0, 2.0]), species=[1, 2])
    v.boundary_dirichlet(s, 1, 2, 1.0)
    rng = np.random.default_rng(2)
    U = np.asfortranarray(rng.uniform(0.1, 1.0, (2, g.num_nodes)))
    o = O.OracleSystem(s)
    Fo, Ao = o.assemble(U, U)
    st = v.SystemState(s)
    try:
        du = v.eval_rhs(st, U.ravel(order="F"))
        J = v.eval_jacobian(st, U.ravel(order="F"))
        assert np.allclose(du, -Fo.ravel(order="F"), rtol=1e-12, atol=1e-13)
        assert np.array_equal(J.indices, Ao.indices) and np.allclose(J.data, -Ao.data, rtol=1e-12, atol=1e-13)
        M = v.mass_matrix(st)
        Mo = o.mass_matrix()
        assert M.ndim == 1  # linear storage: diagonal, like the reference's Diagonal
        assert np.allclose(M.reshape(g.num_nodes, 2), Mo[:, [0, 1], [0, 1]], rtol=1e-14)
        assert M.reshape(g.num_nodes, 2)[:, 0].sum() == pytest.approx(1.0, rel=1e-12)  # storage coefficient 1 x volume of the unit square
    finally:
        st.close()
    # non-diagonal storage (bipolar) with three cell regions -> sparse block-diagonal matrix
    g3 = v.simplexgrid(X, X, X)
    v.cellmask(g3, [0, 0, 0.3], [1, 1, 0.72], 2)
    s3 = v.System(g3, flux=ph.BipolarSGFlux(), reaction=ph.BipolarReaction([10.0, 0.0, -10.0]), storage=ph.BipolarStorage(), species=[1, 2, 3])
    st3 = v.SystemState(s3)
    try:
        M3 = v.mass_matrix(st3)
        Mo3 = O.OracleSystem(s3).mass_matrix()
        assert not isinstance(M3, np.ndarray)
        B = M3.toarray() if g3.num_nodes * 3 < 3000 else None
        d = np.array([M3[3 * K : 3 * K + 3, 3 * K : 3 * K + 3].toarray() for K in range(0, g3.num_nodes, 37)])
        assert np.allclose(d, Mo3[::37], rtol=1e-13, atol=0)
    finally:
        st3.close()


def test_example201_nodeflux_device():
    """examples/Example201_Laplace2D.jl:33-47 through the device path: solve, nodeflux (flux callback on the device), known answer;
    and the device edge fluxes against the oracle's on a non-uniform grid with a nonlinear flux"""
    from test_oracle_golden import example201_system

    s = example201_system()
    st = v.SystemState(s)
    try:
        sol = v.solve(s, state=st, inival=0.0)
        nf = v.nodeflux(s, sol, state=st)
        assert nf.shape == (2, 1, s.grid.num_nodes)
        assert np.linalg.norm(sol) + np.linalg.norm(nf) == pytest.approx(9.63318042491699, rel=1e-12)
    finally:
        st.close()
    X = np.linspace(0, 1, 8) ** 1.5
    g = v.simplexgrid(X, X)
    s2 = v.System(g, flux=ph.PowerDiffusion([0.3, 2.0], 3.0), species=[1, 2])
    U = np.asfortranarray(np.random.default_rng(8).uniform(0.2, 1.0, (2, g.num_nodes)))
    st2 = v.SystemState(s2)
    try:
        prm = np.ascontiguousarray(s2.physics.flux.params(2))
        dev = np.zeros(2 * st2.num_edges)
        st2.set_vector(v._lib.VEC_UPDATE, U)
        v._lib.check(st2.h, st2.L.vfvm_edgeflux(st2.h, s2.physics.flux.id, prm.ctypes.data, prm.size, v._lib.VEC_UPDATE, dev.ctypes.data))
        ref = O.OracleSystem(s2).edgeflux(U, s2.physics.flux.id, prm)
        assert np.allclose(dev.reshape((2, st2.num_edges), order="F"), ref, rtol=1e-13, atol=1e-15)
        nf2 = v.nodeflux(s2, U, state=st2)
        assert nf2.shape == (2, 2, g.num_nodes) and np.all(np.isfinite(nf2))
    finally:
        st2.close()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_boundary_integrals_device(dim):
    """integrate(system, F, U; boundary = true) on the device: known measures of the boundary regions, and agreement with the oracle for the
    registered boundary / node functions on a system with a boundary species"""
    X = np.linspace(0, 1, 9)
    g = v.simplexgrid(*([X] * dim))
    sys = v.System(g, flux=ph.LinearDiffusion([1.0, 1.0, 0.0]), breaction=ph.CatalysisBoundaryReaction(2, S=0.05), bstorage=ph.LinearBoundaryStorage(2, [0.0, 0.0, 1.0]))
    v.enable_species(sys, 1, [1])
    v.enable_species(sys, 2, [1])
    v.enable_boundary_species(sys, 3, [2])
    U = np.asfortranarray(np.random.default_rng(9).uniform(0.1, 1.0, (3, g.num_nodes))) * sys.node_dof()
    o = O.OracleSystem(sys)
    st = v.SystemState(sys)
    try:
        one = np.asfortranarray(np.ones((3, g.num_nodes)) * sys.node_dof())
        B = v.integrate(sys, one, state=st, boundary=True)
        np.testing.assert_allclose(B[0], 1.0, rtol=1e-12)  # |Gamma_r| = 1 for every region of the unit square / cube
        # the surface species is defined at the NODES of region 2 (isnodespecies, src/vfvm_assemblydata.jl:286-292): region 2 itself has measure 1,
        # the faces that share an edge / corner with it see those nodes with their own boundary factors (a strip of width h/2), the opposite face nothing
        h2 = 0.5 * (X[1] - X[0])
        expected = {1: [0.0, 1.0], 2: [h2, 1.0, h2, 0.0], 3: [h2, 1.0, h2, 0.0, h2, h2]}[dim]
        np.testing.assert_allclose(B[2], expected, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(B, o.integrate_boundary(one), rtol=1e-12, atol=1e-15)
        for F in (None, sys.physics.breaction, sys.physics.bstorage, ph.PowerBoundaryReaction(2, [1.0, 0.5, 2.0], 2.0), ph.PowerReaction(1.0, 2.0)):
            dev = v.integrate(sys, U, state=st, boundary=True) if F is None else v.integrate(sys, F, U, state=st, boundary=True)
            ref = o.integrate_boundary(U) if F is None else o.integrate_boundary(U, F.slot, F.id, F.params(3))
            assert dev.shape == ref.shape == (3, 2 * dim)
            assert np.all(np.abs(dev - ref) <= 1e-12 * np.abs(ref) + 1e-14), (F, dev, ref)
    finally:
        st.close()
