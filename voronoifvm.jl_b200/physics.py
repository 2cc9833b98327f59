"""Registered physics library -- host side.

In the reference the constitutive callbacks are arbitrary Julia closures `flux(f,u,edge,data)`,
`reaction/storage(f,u,node,data)`, `source(f,node,data)`, `bcondition(f,u,bnode,data)` stored in
`VoronoiFVM.Physics` (src/vfvm_physics.jl:67-238) and differentiated by ForwardDiff
(src/vfvm_physics.jl:394-466).  On the B200 path a callback must be one of the *registered* objects below:
each carries the id of a device function (include/vfvm_b200.h) plus a parameter block, and is evaluated on
the GPU in fp64 forward-mode dual numbers.  Passing anything else raises `UnregisteredPhysicsError`
-- there is no CPU fallback.

Species numbers and region numbers are labels and stay 1-based as in the reference.
"""
from __future__ import annotations

import numpy as np

# ids: keep in sync with include/vfvm_b200.h
SLOT_FLUX, SLOT_REACTION, SLOT_STORAGE, SLOT_SOURCE, SLOT_BREACTION, SLOT_EDGEREACTION, SLOT_BSTORAGE = range(7)
SLOT_NAMES = ("flux", "reaction", "storage", "source", "breaction", "edgereaction", "bstorage")

FLUX_DIFFUSION, FLUX_POWDIFF, FLUX_CROSSDIFF2, FLUX_SG_UNIPOLAR, FLUX_SEDAN, FLUX_SG_BIPOLAR, FLUX_MIXTURE = 1, 2, 3, 4, 5, 6, 7
REACTION_POW, REACTION_SINH, REACTION_AFFINE, REACTION_BILINEAR2, REACTION_BIPOLAR, REACTION_REGION_AFFINE = 1, 2, 3, 4, 5, 6
STORAGE_LINEAR, STORAGE_POW, STORAGE_BIPOLAR = 1, 2, 3
SOURCE_CONST, SOURCE_GAUSS, SOURCE_XSINYEXPZ, SOURCE_STEP1D, SOURCE_AFFINE_X, SOURCE_NODAL = 1, 2, 3, 4, 5, 6
BREACTION_LINEAR, BREACTION_CATALYSIS, BREACTION_POW = 1, 2, 3
EDGEREACTION_DIAMOND, EDGEREACTION_JOULE = 1, 2
BSTORAGE_LINEAR = 1
BC_DIRICHLET, BC_NEUMANN, BC_ROBIN = 1, 2, 3


class UnregisteredPhysicsError(TypeError):
    """north_star: unregistered arbitrary callbacks raise an error instead of silently falling back to the CPU."""


def _vec(x, n):
    a = np.asarray(x, dtype=np.float64).ravel()
    if a.size == 1:
        a = np.full(n, a[0])
    if a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


class RegisteredPhysics:
    slot: int = -1
    id: int = 0
    min_species: int = 1

    def params(self, n: int) -> np.ndarray:  # parameter block for n species
        raise NotImplementedError

    def __call__(self, *args, **kw):
        raise UnregisteredPhysicsError(f"{type(self).__name__} is evaluated on the device; it cannot be called on the host")


# ------------------------------------------------------------------------------------------- flux
class LinearDiffusion(RegisteredPhysics):
    """f_i = D_i (u_i,K - u_i,L)   (examples/Example201_Laplace2D.jl:17-20, Example301:17-20, Example410:20-25)"""

    slot, id = SLOT_FLUX, FLUX_DIFFUSION

    def __init__(self, D=1.0):
        self.D = D

    def params(self, n):
        return _vec(self.D, n)


class PowerDiffusion(RegisteredPhysics):
    """f_i = D_i (u_i,K^m - u_i,L^m)   (Example207:35-38 with m=2, Example106:49-52)"""

    slot, id = SLOT_FLUX, FLUX_POWDIFF

    def __init__(self, D=1.0, m=2.0):
        self.D, self.m = D, float(m)

    def params(self, n):
        return np.concatenate([_vec(self.D, n), [self.m]])


class CrossDiffusion2(RegisteredPhysics):
    """Example110_ReactionDiffusion1D_TwoSpecies.jl:43-50"""

    slot, id, min_species = SLOT_FLUX, FLUX_CROSSDIFF2, 2

    def __init__(self, eps=(1.0, 1.0), c=0.01):
        self.eps, self.c = eps, c

    def params(self, n):
        return np.array([self.eps[0], self.eps[1], self.c], dtype=np.float64)


class UnipolarSGFlux(RegisteredPhysics):
    """Example160 `classflux!` :43-50 (Scharfetter-Gummel with fbernoulli_pm)"""

    slot, id, min_species = SLOT_FLUX, FLUX_SG_UNIPOLAR, 2

    def __init__(self, eps, iphi=1, ic=2):
        self.eps, self.iphi, self.ic = eps, iphi, ic

    def params(self, n):
        return np.array([self.eps, self.iphi - 1, self.ic - 1], dtype=np.float64)


class SedanFlux(RegisteredPhysics):
    """Example160 `sedanflux!` :68-77"""

    slot, id, min_species = SLOT_FLUX, FLUX_SEDAN, 2

    def __init__(self, eps, z, iphi=1, ic=2, eps_reg=1.0e-10):
        self.eps, self.z, self.iphi, self.ic, self.eps_reg = eps, z, iphi, ic, eps_reg

    def params(self, n):
        return np.array([self.eps, self.z, self.iphi - 1, self.ic - 1, self.eps_reg], dtype=np.float64)


class BipolarSGFlux(RegisteredPhysics):
    """Example161 `flux!` :134-150"""

    slot, id, min_species = SLOT_FLUX, FLUX_SG_BIPOLAR, 3

    def __init__(self, lam=0.1, mun=1.0, mup=10.0, zn=-1.0, zp=1.0, En=1.0, Ep=0.0, iphin=1, iphip=2, ipsi=3):
        self.p = [lam, mun, mup, zn, zp, En, Ep, iphin - 1, iphip - 1, ipsi - 1]

    def params(self, n):
        return np.array(self.p, dtype=np.float64)


class MixtureFlux(RegisteredPhysics):
    """Maxwell-Stefan mixture flux of DevEx005_Mixture.jl:74-104: f = M(u)^{-1} (u_K - u_L), M_ii = 1/DK_i + sum_{j != i} au_j / DB_ij,
    M_ij = -au_i / DB_ij, au = (u_K + u_L) / 2; the n x n system is solved inside the callback (`inplace_linsolve!`)"""

    slot, id, min_species = SLOT_FLUX, FLUX_MIXTURE, 2

    def __init__(self, DKnudsen, DBinary):
        self.DK = np.asarray(DKnudsen, dtype=np.float64).ravel()
        self.DB = np.asarray(DBinary, dtype=np.float64)

    def params(self, n):
        assert self.DK.size == n and self.DB.shape == (n, n)
        DB = self.DB.copy()
        np.fill_diagonal(DB, 1.0)  # never read
        return np.concatenate([self.DK, DB.ravel(order="C")])


# ------------------------------------------------------------------------------------------- reaction
class PowerReaction(RegisteredPhysics):
    """f_i = k_i u_i^p_i   (Example207:32-34: u^2)"""

    slot, id = SLOT_REACTION, REACTION_POW

    def __init__(self, k=1.0, p=2.0):
        self.k, self.p = k, p

    def params(self, n):
        return np.concatenate([_vec(self.k, n), _vec(self.p, n)])


class SinhReaction(RegisteredPhysics):
    """f_i = k_i (exp(u_i) - exp(-u_i))   (Example105:55-58)"""

    slot, id = SLOT_REACTION, REACTION_SINH

    def __init__(self, k=1.0):
        self.k = k

    def params(self, n):
        return _vec(self.k, n)


class AffineReaction(RegisteredPhysics):
    """f = R u + r0   (Example210:27-31, Example160 reaction! :60-66)"""

    slot, id = SLOT_REACTION, REACTION_AFFINE

    def __init__(self, R, r0=None):
        self.R = np.atleast_2d(np.asarray(R, dtype=np.float64))
        self.r0 = r0

    def params(self, n):
        assert self.R.shape == (n, n)
        r0 = np.zeros(n) if self.r0 is None else _vec(self.r0, n)
        return np.concatenate([self.R.ravel(order="C"), r0])


class BilinearReaction2(RegisteredPhysics):
    """f_1 = k u_1 u_2, f_2 = -k u_1 u_2   (Example110:38-42)"""

    slot, id, min_species = SLOT_REACTION, REACTION_BILINEAR2, 2

    def __init__(self, k=1.0):
        self.k = k

    def params(self, n):
        return np.array([self.k], dtype=np.float64)


class RegionAffineReaction(RegisteredPhysics):
    """f = R_r u + r0_r in cell region r: one affine map per region (Example221 reaction :54-66)"""

    slot, id = SLOT_REACTION, REACTION_REGION_AFFINE

    def __init__(self, R, r0=None):
        self.R = [np.asarray(m, dtype=np.float64) for m in R]
        self.r0 = [None] * len(self.R) if r0 is None else list(r0)

    def params(self, n):
        out = [np.array([float(len(self.R))])]
        for m, c in zip(self.R, self.r0):
            assert m.shape == (n, n)
            out += [m.ravel(), np.zeros(n) if c is None else _vec(c, n)]
        return np.concatenate(out)


class BipolarReaction(RegisteredPhysics):
    """Example161 `reaction!` :109-132; `doping[r-1]` is C in cell region r (Cn, Ca, -Cp there)"""

    slot, id, min_species = SLOT_REACTION, REACTION_BIPOLAR, 3

    def __init__(self, doping, zn=-1.0, zp=1.0, En=1.0, Ep=0.0, r0=1.0, iphin=1, iphip=2, ipsi=3):
        self.doping = np.asarray(doping, dtype=np.float64).ravel()
        self.p = [zn, zp, En, Ep, r0, iphin - 1, iphip - 1, ipsi - 1]

    def params(self, n):
        return np.concatenate([self.p, [self.doping.size], self.doping])


# ------------------------------------------------------------------------------------------- storage
class LinearStorage(RegisteredPhysics):
    """f_i = c_i u_i   (Example207:44-47; Example160 storage! :52-58 with c = (0, 1))"""

    slot, id = SLOT_STORAGE, STORAGE_LINEAR

    def __init__(self, c=1.0):
        self.c = c

    def params(self, n):
        return _vec(self.c, n)


class PowerStorage(RegisteredPhysics):
    """f_i = (eps_i + u_i)^(1/m_i)   (Example107:52-55)"""

    slot, id = SLOT_STORAGE, STORAGE_POW

    def __init__(self, eps=1.0e-10, m=2.0):
        self.eps, self.m = eps, m

    def params(self, n):
        return np.concatenate([_vec(self.eps, n), _vec(self.m, n)])


class BipolarStorage(RegisteredPhysics):
    """Example161 `storage!` :163-170"""

    slot, id, min_species = SLOT_STORAGE, STORAGE_BIPOLAR, 3

    def __init__(self, zn=-1.0, zp=1.0, En=1.0, Ep=0.0, iphin=1, iphip=2, ipsi=3):
        self.p = [zn, zp, En, Ep, iphin - 1, iphip - 1, ipsi - 1]

    def params(self, n):
        return np.array(self.p, dtype=np.float64)


# ------------------------------------------------------------------------------------------- source
class ConstSource(RegisteredPhysics):
    slot, id = SLOT_SOURCE, SOURCE_CONST

    def __init__(self, s=1.0):
        self.s = s

    def params(self, n):
        return _vec(self.s, n)


class GaussSource(RegisteredPhysics):
    """f_sp = exp(-a |x - c|^2)   (Example207:39-43, Example210:39-44)"""

    slot, id = SLOT_SOURCE, SOURCE_GAUSS

    def __init__(self, species=1, a=20.0, center=(0.5, 0.5, 0.5)):
        c = list(center) + [0.0] * (3 - len(center))
        self.p = [species - 1, a] + c[:3]

    def params(self, n):
        return np.array(self.p, dtype=np.float64)


class XSinYExpZSource(RegisteredPhysics):
    """f_sp = x sin(b y) exp(z)   (Example301:22-26 with b = 5)"""

    slot, id = SLOT_SOURCE, SOURCE_XSINYEXPZ

    def __init__(self, species=1, b=5.0):
        self.p = [species - 1, b]

    def params(self, n):
        return np.array(self.p, dtype=np.float64)


class Step1DSource(RegisteredPhysics):
    """f_sp = x <= x0 ? lo : hi   (Example105:45-52)"""

    slot, id = SLOT_SOURCE, SOURCE_STEP1D

    def __init__(self, species=1, x0=0.5, lo=1.0, hi=-1.0):
        self.p = [species - 1, x0, lo, hi]

    def params(self, n):
        return np.array(self.p, dtype=np.float64)


class AffineXSource(RegisteredPhysics):
    """f_i = a_i + b_i x   (Example110:51-55)"""

    slot, id = SLOT_SOURCE, SOURCE_AFFINE_X

    def __init__(self, a, b):
        self.a, self.b = a, b

    def params(self, n):
        return np.concatenate([_vec(self.a, n), _vec(self.b, n)])


class NodalSource(RegisteredPhysics):
    """Source given as an n x N table.  The source callback does not depend on u (src/vfvm_physics.jl:335-357,
    ResEvaluator without AD), so a host closure `fn(x) -> (n,)` may be tabulated once and uploaded."""

    slot, id = SLOT_SOURCE, SOURCE_NODAL

    def __init__(self, table=None, fn=None):
        self.table, self.fn = table, fn

    def tabulate(self, grid, n):
        if self.table is not None:
            t = np.asfortranarray(self.table, dtype=np.float64)
        else:
            t = np.zeros((n, grid.num_nodes), order="F")
            for k in range(grid.num_nodes):
                t[:, k] = self.fn(grid.coord[:, k])
        assert t.shape == (n, grid.num_nodes)
        return t

    def params(self, n):
        return np.zeros(0)


# ------------------------------------------------------------------------------------------- boundary
class LinearBoundaryReaction(RegisteredPhysics):
    """if bnode.region == region: f = R u   (Example215:33-42)"""

    slot, id = SLOT_BREACTION, BREACTION_LINEAR

    def __init__(self, region, R):
        self.region = region
        self.R = np.atleast_2d(np.asarray(R, dtype=np.float64))

    def params(self, n):
        assert self.R.shape == (n, n)
        return np.concatenate([[self.region], self.R.ravel(order="C")])


class PowerBoundaryReaction(RegisteredPhysics):
    """if bnode.region == region: f_i = k_i u_i^p_i   (Example226_BoundaryIntegral.jl:42-47: u^2)"""

    slot, id = SLOT_BREACTION, BREACTION_POW

    def __init__(self, region, k=1.0, p=2.0):
        self.region, self.k, self.p = region, k, p

    def params(self, n):
        return np.concatenate([[float(self.region)], _vec(self.k, n), _vec(self.p, n)])


class CatalysisBoundaryReaction(RegisteredPhysics):
    """Example115 `breaction!` :125-135 on boundary region `region`: surface species C exchanges with the bulk species A and B,
    R_XC = kp_XC u_X (1 - u_C) - km_XC u_C;  f_A = S R_AC, f_B = S R_BC, f_C = -R_BC - R_AC"""

    slot, id, min_species = SLOT_BREACTION, BREACTION_CATALYSIS, 3

    def __init__(self, region, S=0.01, kp_AC=100.0, km_AC=1.0, kp_BC=0.1, km_BC=1.0, iA=1, iB=2, iC=3):
        self.p = [region, S, kp_AC, km_AC, kp_BC, km_BC, iA - 1, iB - 1, iC - 1]

    def params(self, n):
        return np.array(self.p, dtype=np.float64)


# ------------------------------------------------------------------------------------------- edge reaction
class DiamondEdgeReaction(RegisteredPhysics):
    """f_i = c_i h^2 / (2 dim), h = meas(edge): a constant volume density given per edge (DevEx002_EdgeReaction.jl:83-87)"""

    slot, id = SLOT_EDGEREACTION, EDGEREACTION_DIAMOND

    def __init__(self, c=-1.0):
        self.c = c

    def params(self, n):
        return _vec(self.c, n)


class JouleHeatEdgeReaction(RegisteredPhysics):
    """f_iT = -kappa (u_iphi,K - u_iphi,L)^2   (Example206_JouleHeat.jl:83-86)"""

    slot, id, min_species = SLOT_EDGEREACTION, EDGEREACTION_JOULE, 2

    def __init__(self, kappa=1.0, iphi=1, iT=2):
        self.p = [kappa, iphi - 1, iT - 1]

    def params(self, n):
        return np.array(self.p, dtype=np.float64)


# ------------------------------------------------------------------------------------------- boundary storage
class LinearBoundaryStorage(RegisteredPhysics):
    """if bnode.region == region: f_i = c_i u_i   (Example115:138-143, Example311:78-83)"""

    slot, id = SLOT_BSTORAGE, BSTORAGE_LINEAR

    def __init__(self, region, c):
        self.region, self.c = region, c

    def params(self, n):
        return np.concatenate([[float(self.region)], _vec(self.c, n)])


class BCondition(RegisteredPhysics):
    """A `bcondition` callback made of boundary_dirichlet!/neumann!/robin! calls (src/vfvm_physics.jl:487-564),
    optionally after a registered boundary reaction.  `region=None` means "all boundary regions"
    (the keyword default `region = bnode.region`)."""

    slot = SLOT_BREACTION

    def __init__(self, reaction: RegisteredPhysics | None = None):
        self.reaction = reaction
        self.entries = []

    @property
    def id(self):
        return 0 if self.reaction is None else self.reaction.id

    def params(self, n):
        return np.zeros(0) if self.reaction is None else self.reaction.params(n)

    def _add(self, kind, species, region, value, factor=0.0, ramp=None):
        e = dict(kind=kind, species=species - 1, region=0 if region is None else int(region), value=float(value), factor=float(factor), has_ramp=0, t0=0.0, t1=0.0, v0=0.0, v1=0.0)
        if ramp is not None:  # ramp(bnode.time; dt=(t0,t1), du=(v0,v1))
            (t0, t1), (v0, v1) = ramp
            e.update(has_ramp=1, t0=float(t0), t1=float(t1), v0=float(v0), v1=float(v1))
        self.entries.append(e)
        return self

    def dirichlet(self, species=1, region=None, value=0.0, ramp=None):
        return self._add(BC_DIRICHLET, species, region, value, ramp=ramp)

    def neumann(self, species=1, region=None, value=0.0, ramp=None):
        return self._add(BC_NEUMANN, species, region, value, ramp=ramp)

    def robin(self, species=1, region=None, factor=0.0, value=0.0, ramp=None):
        return self._add(BC_ROBIN, species, region, value, factor=factor, ramp=ramp)


_UNSUPPORTED = ("bflux", "bsource", "boutflow", "generic_operator", "generic_operator_sparsity")


class Physics:
    """Mirror of `VoronoiFVM.Physics(; flux, reaction, storage, source, breaction/bcondition, data, ...)`
    (src/vfvm_physics.jl:184-238) restricted to registered device callbacks."""

    def __init__(self, flux=None, reaction=None, storage=None, source=None, breaction=None, bcondition=None, edgereaction=None, bstorage=None, data=None, **other):
        for k, v in other.items():
            if k in _UNSUPPORTED and v is not None:
                raise NotImplementedError(f"physics callback `{k}` is outside the B200 hot-path scope (SURVEY.md section 8f)")
            if k not in _UNSUPPORTED:
                raise TypeError(f"unknown Physics keyword {k!r}")
        if breaction is not None and bcondition is not None:
            raise ValueError("specify either breaction or bcondition")
        breaction = bcondition if bcondition is not None else breaction
        if breaction is not None and not isinstance(breaction, BCondition):
            if isinstance(breaction, RegisteredPhysics):
                breaction = BCondition(reaction=breaction)
        self.data = data
        self.slots = [flux, reaction, storage, source, breaction, edgereaction, bstorage]
        for name, cb, slot in zip(SLOT_NAMES, self.slots, range(7)):
            if cb is None:
                continue
            if not isinstance(cb, RegisteredPhysics):
                raise UnregisteredPhysicsError(
                    f"{name}={cb!r} is not a registered device callback; the B200 backend does not fall back to the CPU. "
                    "Use one of the classes in vfvm_b200.physics (or register a new device function)."
                )
            if cb.slot != slot:
                raise UnregisteredPhysicsError(f"{type(cb).__name__} cannot be used as `{name}`")

    flux = property(lambda self: self.slots[0])
    reaction = property(lambda self: self.slots[1])
    storage = property(lambda self: self.slots[2])
    source = property(lambda self: self.slots[3])
    breaction = property(lambda self: self.slots[4])
    edgereaction = property(lambda self: self.slots[5])
    bstorage = property(lambda self: self.slots[6])
