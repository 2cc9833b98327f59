"""Post-processing on the device twin: `integrate`, `edgeintegrate`, the discrete norms and `nodevolumes`
(src/vfvm_postprocess.jl:18-67, 94-98, 109-146, 278-343, 397-411).

`F` must be a registered physics object (the reference takes any callback with the reaction / flux signature): node functions
are the registered reaction and storage objects, edge functions the registered fluxes.  Everything runs on the resident
device vectors; only the n x ncellregions result comes back.
"""
from __future__ import annotations


import numpy as np

from . import _lib
from ._lib import check
from .physics import RegisteredPhysics, UnregisteredPhysicsError, BCondition, SLOT_BREACTION, SLOT_BSTORAGE, SLOT_FLUX, SLOT_REACTION, SLOT_STORAGE
from .state import SystemState
from .system import System


def _with_state(system, state):
    if state is not None:
        return state, False
    return SystemState(system), True


def _call(state: SystemState, U, fn, nregions=None):
    state.sync()
    state.set_vector(_lib.VEC_UPDATE, np.asfortranarray(np.asarray(U, dtype=np.float64)))  # scratch vector of the Newton loop
    nreg = state.system.grid.num_cellregions if nregions is None else nregions
    out = np.zeros(state.n * nreg)
    check(state.h, fn(out))
    return out.reshape((state.n, nreg), order="F")


def integrate(system: System, F, U=None, state: SystemState | None = None, boundary=False):
    """`integrate(system, F, U)`: region-wise integrals of a node function -> (nspecies, ncellregions);
    `integrate(system, U)` integrates the solution itself (src/vfvm_postprocess.jl:94-98)"""
    if U is None:
        F, U = None, F
    if boundary:  # src/vfvm_postprocess.jl:29-46 -> (nspecies, nbfaceregions)
        if isinstance(F, BCondition):
            F = F.reaction  # the boundary reaction function itself; boundary_dirichlet!/... helper calls are not integrands
        if F is not None and not (isinstance(F, RegisteredPhysics) and F.slot in (SLOT_BREACTION, SLOT_BSTORAGE, SLOT_REACTION, SLOT_STORAGE)):
            raise UnregisteredPhysicsError("integrate(boundary=True): F must be a registered boundary reaction / boundary storage / reaction / storage object")
        st, own = _with_state(system, state)
        try:
            n = system.num_species
            slot, pid, prm = (SLOT_BREACTION, 0, np.zeros(0)) if F is None else (F.slot, F.id, np.ascontiguousarray(F.params(n), dtype=np.float64))
            return _call(st, U, lambda out: st.L.vfvm_integrate_boundary(st.h, slot, pid, prm.ctypes.data if prm.size else None, prm.size, _lib.VEC_UPDATE, out.ctypes.data),
                         nregions=system.grid.num_bfaceregions)
        finally:
            if own:
                st.close()
    if F is not None and not (isinstance(F, RegisteredPhysics) and F.slot in (SLOT_REACTION, SLOT_STORAGE)):
        raise UnregisteredPhysicsError("integrate: F must be a registered reaction or storage object")
    st, own = _with_state(system, state)
    try:
        n = system.num_species
        if F is None:
            slot, pid, prm = SLOT_REACTION, 0, np.zeros(0)
        else:
            slot, pid, prm = F.slot, F.id, np.ascontiguousarray(F.params(n), dtype=np.float64)
        return _call(st, U, lambda out: st.L.vfvm_integrate(st.h, slot, pid, prm.ctypes.data if prm.size else None, prm.size, _lib.VEC_UPDATE, out.ctypes.data))
    finally:
        if own:
            st.close()


class W1pIntegrand:
    """edge function of `w1pseminorm`: y_i = dim ((u_iK - u_iL) / h)^p (src/vfvm_postprocess.jl:304-310)"""

    slot, id = SLOT_FLUX, -1

    def __init__(self, p=2.0):
        self.p = float(p)

    def params(self, n):
        return np.array([self.p])


class EdgeAverage:
    """edge function y_i = (u_iK + u_iL) / 2 (test/test120_norms.jl:35-38)"""

    slot, id = SLOT_FLUX, -2

    def params(self, n):
        return np.zeros(0)


def edgeintegrate(system: System, F, U, state: SystemState | None = None):
    """`edgeintegrate(system, F, U)`: region-wise integrals of an edge function (diamond volumes h^2 sigma/h / dim)"""
    if not ((isinstance(F, RegisteredPhysics) and F.slot == SLOT_FLUX) or isinstance(F, (W1pIntegrand, EdgeAverage))):
        raise UnregisteredPhysicsError("edgeintegrate: F must be a registered flux object")
    st, own = _with_state(system, state)
    try:
        prm = np.ascontiguousarray(F.params(system.num_species), dtype=np.float64)
        return _call(st, U, lambda out: st.L.vfvm_edgeintegrate(st.h, F.id, prm.ctypes.data if prm.size else None, prm.size, _lib.VEC_UPDATE, out.ctypes.data))
    finally:
        if own:
            st.close()


def lpnorm(system, u, p, species_weights=None, state=None):
    from .physics import PowerReaction

    w = np.ones(system.num_species) if species_weights is None else np.asarray(species_weights, dtype=float)
    II = integrate(system, PowerReaction(1.0, float(p)), u, state=state)
    return float((II.sum(axis=1) * w).sum() ** (1.0 / p))


def l2norm(system, u, species_weights=None, state=None):
    return lpnorm(system, u, 2, species_weights, state)


def w1pseminorm(system, u, p, species_weights=None, state=None):
    w = np.ones(system.num_species) if species_weights is None else np.asarray(species_weights, dtype=float)
    II = edgeintegrate(system, W1pIntegrand(p), u, state=state)
    return float((II.sum(axis=1) * w).sum() ** (1.0 / p))


def h1seminorm(system, u, species_weights=None, state=None):
    return w1pseminorm(system, u, 2, species_weights, state)


def w1pnorm(system, u, p, species_weights=None, state=None):
    return lpnorm(system, u, p, species_weights, state) + w1pseminorm(system, u, p, species_weights, state)


def h1norm(system, u, species_weights=None, state=None):
    return w1pnorm(system, u, 2, species_weights, state)


def nodevolumes(system: System, state: SystemState | None = None):
    """volumes of the Voronoi cells (src/vfvm_postprocess.jl:397-411): node factors summed over the cell regions"""
    st, own = _with_state(system, state)
    try:
        colptr, _, fac = st.nodefactors()
        return np.add.reduceat(fac, colptr[:-1])
    finally:
        if own:
            st.close()


# ---- ODE interface entry points (src/vfvm_diffeq_interface.jl:10-101) on the device twin ---------------------------------
def eval_rhs(state: SystemState, u, t=0.0):
    """`eval_rhs!(du, u, state, t)`: du = -residual of the stationary operator (tstep = Inf; the ODE solver owns the time derivative)"""
    U = np.asarray(u, dtype=np.float64).reshape((state.n, state.N), order="F")
    return -state.eval_res_jac(U, time=float(t)).ravel(order="F")


def eval_jacobian(state: SystemState, u, t=0.0):
    """`eval_jacobian!(J, u, state, t)`: J = -Jacobian of the stationary operator (CSC, the reference's pattern)"""
    U = np.asarray(u, dtype=np.float64).reshape((state.n, state.N), order="F")
    state.eval_res_jac(U, time=float(t))
    return -state.matrix("csc")


def mass_matrix(state: SystemState):
    """`mass_matrix(state)`: storage Jacobian at U = 0 times the node volumes; a 1D array (the diagonal) if it is diagonal, like the
    reference's `Diagonal`, else a sparse block-diagonal matrix"""
    import scipy.sparse as sp

    state.sync()
    n, N = state.n, state.Nown
    out = np.zeros(N * n * n)
    check(state.h, state.L.vfvm_mass_matrix(state.h, out.ctypes.data))
    blocks = out.reshape((N, n, n))
    off = blocks.copy()
    off[:, np.arange(n), np.arange(n)] = 0.0
    if not off.any():
        return blocks[:, np.arange(n), np.arange(n)].ravel()
    return sp.bsr_matrix((blocks, np.arange(N), np.arange(N + 1)), shape=(N * n, N * n)).tocsc()


# ---- nodeflux (src/vfvm_postprocess.jl:167-252) ------------------------------------------------------------------------------
def voronoi_face_centers(grid, edgenodes, celledges):
    """x_sigma per edge, the point of the Voronoi face the "magic formula" (Eymard/Gallouet/Herbin 2006, Lemma 2.4) is evaluated at.
    ExtendableGrids' `VoronoiFaceCenters` is not under /root/reference; restated here as: 1D the edge midpoint; 2D the midpoint of the
    two adjacent triangles' circumcentres for interior edges and the edge midpoint for boundary edges.  PARITY UNPINNED beyond the
    reference's known answer of Example201 (uniform grid, where every consistent choice coincides); 3D is not provided."""
    x = grid.coord
    mid = 0.5 * (x[:, edgenodes[0]] + x[:, edgenodes[1]])
    if grid.dim == 1:
        return mid
    if grid.dim != 2:
        raise NotImplementedError("nodeflux: Voronoi face centres are restated for 1D and 2D grids only")
    a, b, c = (x[:, grid.cellnodes[k]] for k in range(3))
    d = 2.0 * (a[0] * (b[1] - c[1]) + b[0] * (c[1] - a[1]) + c[0] * (a[1] - b[1]))
    a2, b2, c2 = (a * a).sum(0), (b * b).sum(0), (c * c).sum(0)
    cc = np.stack([(a2 * (b[1] - c[1]) + b2 * (c[1] - a[1]) + c2 * (a[1] - b[1])) / d, (a2 * (c[0] - b[0]) + b2 * (a[0] - c[0]) + c2 * (b[0] - a[0])) / d])
    E = edgenodes.shape[1]
    acc, cnt = np.zeros((2, E)), np.zeros(E)
    for k in range(3):
        np.add.at(acc[0], celledges[k], cc[0])
        np.add.at(acc[1], celledges[k], cc[1])
        np.add.at(cnt, celledges[k], 1.0)
    return np.where(cnt >= 2, acc / np.maximum(cnt, 1.0), mid)


def _nodeflux_from_edgeflux(grid, edgenodes, xsigma, efac, nodevol, flux):
    """nodeflux[:, i, K] += fac f_i (x_sigma - x_K), nodeflux[:, i, L] -= fac f_i (x_sigma - x_L), then / nodevol (:204-217)"""
    n, dim, N = flux.shape[0], grid.dim, grid.num_nodes
    K, L = edgenodes[0], edgenodes[1]
    out = np.zeros((dim, n, N))
    dK, dL = xsigma - grid.coord[:, K], xsigma - grid.coord[:, L]
    for d in range(dim):
        for i in range(n):
            w = efac * flux[i]
            np.add.at(out[d, i], K, w * dK[d])
            np.add.at(out[d, i], L, -w * dL[d])
    return out / nodevol[None, None, :]


def nodeflux(system: System, U, F=None, state: SystemState | None = None):
    """`nodeflux(system, U)` / `nodeflux(system, F, U)`: reconstruction of the edge flux as a vector field on the nodes -> (dim, nspecies, nnodes).
    The flux callback runs on the device (`vfvm_edgeflux`); the accumulation with the Voronoi face centres is host-side post-processing."""
    F = system.physics.flux if F is None else F
    if not (isinstance(F, RegisteredPhysics) and F.slot == SLOT_FLUX):
        raise UnregisteredPhysicsError("nodeflux: the flux must be a registered flux object")
    st, own = _with_state(system, state)
    try:
        st.sync()
        st.set_vector(_lib.VEC_UPDATE, np.asfortranarray(np.asarray(U, dtype=np.float64)))
        prm = np.ascontiguousarray(F.params(system.num_species), dtype=np.float64)
        flux = np.zeros(st.n * st.num_edges)
        check(st.h, st.L.vfvm_edgeflux(st.h, F.id, prm.ctypes.data if prm.size else None, prm.size, _lib.VEC_UPDATE, flux.ctypes.data))
        en = st.edgenodes()
        cp, _, ef = st.edgefactors()
        efac = np.add.reduceat(np.append(ef, 0.0), cp[:-1]) * (cp[1:] > cp[:-1])
        np_, _, nf = st.nodefactors()
        nodevol = np.add.reduceat(nf, np_[:-1])
        xs = voronoi_face_centers(system.grid, en, st.celledges())
        return _nodeflux_from_edgeflux(system.grid, en, xs, efac, nodevol, flux.reshape((st.n, st.num_edges), order="F"))
    finally:
        if own:
            st.close()
