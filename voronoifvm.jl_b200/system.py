"""Host mirror of `VoronoiFVM.System` (src/vfvm_system.jl:16-121, :215-261): a backend-agnostic description
of grid + species + registered physics + legacy boundary tables.  The device twin lives in `state.py`.

Labels (species, regions) are 1-based as in the reference; array indices are 0-based.
Solution arrays are numpy arrays of shape (nspecies, nnodes) in Fortran order, i.e. exactly the memory layout
of the reference's dense solution (dof = K*nspecies + ispec, src/vfvm_densesolution.jl:49).
"""
from __future__ import annotations

import numpy as np

from .grid import Grid
from .physics import BCondition, NodalSource, Physics

DIRICHLET = 1.0e30  # Dirichlet(Tv), src/vfvm_system.jl:329


class System:
    def __init__(self, grid: Grid, physics: Physics | None = None, *, species=None, assembly="edgewise", unknown_storage="dense",
                 matrixtype="sparse", is_linear=False, nparams=0, **physics_kwargs):
        if not isinstance(grid, Grid):
            raise TypeError("grid must be a vfvm_b200 Grid")
        if matrixtype != "sparse":
            raise NotImplementedError("matrixtype != :sparse (tridiagonal/banded) is a 1D CPU convenience outside the hot path")
        if nparams != 0:
            raise NotImplementedError("parameter derivatives (nparams > 0) are a 'next' row (SURVEY.md section 8f)")
        if assembly not in ("edgewise", "cellwise"):
            raise ValueError("assembly must be 'edgewise' or 'cellwise'")
        if unknown_storage not in ("dense", "sparse"):
            raise ValueError("specify either unknown_storage='dense' or unknown_storage='sparse'")
        self.grid = grid
        # the device path always uses the edgewise data (src/vfvm_system.jl:690-755); :cellwise gives the same
        # operator up to summation order, so it is accepted and mapped onto the same kernels
        self.assembly_type = assembly
        self.unknown_storage = unknown_storage
        self.is_linear = bool(is_linear)
        self.num_species = 0
        self.region_species = np.zeros((0, grid.num_cellregions), dtype=np.uint8, order="F")
        self.bregion_species = np.zeros((0, grid.num_bfaceregions), dtype=np.uint8, order="F")  # enable_boundary_species!
        self.boundary_factors = np.zeros((0, grid.num_bfaceregions), order="F")
        self.boundary_values = np.zeros((0, grid.num_bfaceregions), order="F")
        self.physics = physics if physics is not None else Physics(**physics_kwargs)
        if physics is not None and physics_kwargs:
            raise TypeError("pass either a Physics object or physics keyword arguments")
        self._version = 0  # bumped on every mutation so a SystemState can tell it is stale
        if species is not None:
            enable_species(self, species=species)

    # -- src/vfvm_system.jl:378-395 increase_num_species!
    def _increase_num_species(self, nspec: int):
        if nspec <= self.num_species:
            return
        def grow(a, dtype):
            b = np.zeros((nspec, a.shape[1]), dtype=dtype, order="F")
            b[: a.shape[0], :] = a
            return b
        self.region_species = grow(self.region_species, np.uint8)
        self.bregion_species = grow(self.bregion_species, np.uint8)
        self.boundary_factors = grow(self.boundary_factors, np.float64)
        self.boundary_values = grow(self.boundary_values, np.float64)
        self.num_species = nspec
        self._version += 1

    @property
    def num_nodes(self):
        return self.grid.num_nodes

    def node_dof(self) -> np.ndarray:
        """(n, N) mask: species defined in node (enable_species!, src/vfvm_system.jl:445-456)"""
        g = self.grid
        mask = np.zeros((self.num_species, g.num_nodes), dtype=bool, order="F")
        for ireg in range(g.num_cellregions):
            nodes = np.unique(g.cellnodes[:, g.cellregions == ireg + 1])
            for i in range(self.num_species):
                if self.region_species[i, ireg]:
                    mask[i, nodes] = True
        for ibreg in range(g.num_bfaceregions):  # boundary species, src/vfvm_system.jl:502-513
            if self.bregion_species[:, ibreg].any():
                nodes = np.unique(g.bfacenodes[:, g.bfaceregions == ibreg + 1])
                for i in range(self.num_species):
                    if self.bregion_species[i, ibreg]:
                        mask[i, nodes] = True
        return mask

    def has_legacy_bc(self) -> bool:
        return bool(np.any(self.boundary_factors != 0) or np.any(self.boundary_values != 0))

    def physics_slots(self):
        """[(slot, id, params)] for the C ABI, + bc entries + optional nodal source table"""
        n = self.num_species
        out = []
        for slot, cb in enumerate(self.physics.slots):
            if cb is None:
                out.append((slot, 0, np.zeros(0)))
                continue
            if n < cb.min_species:
                raise ValueError(f"{type(cb).__name__} needs at least {cb.min_species} species")
            out.append((slot, int(cb.id), np.ascontiguousarray(cb.params(n), dtype=np.float64)))
        return out

    def bc_entries(self):
        b = self.physics.breaction
        return list(b.entries) if isinstance(b, BCondition) else []

    def nodal_source(self):
        s = self.physics.source
        return s.tabulate(self.grid, self.num_species) if isinstance(s, NodalSource) else None


def physics(system: System, phys: Physics | None = None, **kw):
    """`physics!(system, physics)` src/vfvm_system.jl:335-356"""
    system.physics = phys if phys is not None else Physics(**kw)
    system._version += 1
    return system


def enable_species(system: System, ispec=None, regions=None, *, species=None):
    """`enable_species!(system, ispec, regions)` src/vfvm_system.jl:433-480"""
    if species is None:
        species = ispec
    if species is None:
        raise ValueError("no species given")
    if np.isscalar(species):
        species = [species]
    if regions is None:
        regions = range(1, system.grid.num_cellregions + 1)
    for isp in species:
        system._increase_num_species(int(isp))
        for ireg in regions:
            system.region_species[int(isp) - 1, int(ireg) - 1] = 1
    system._version += 1
    return system


def enable_boundary_species(system: System, ispec: int, bregions):
    """`enable_boundary_species!(system, ispec, bregions)` src/vfvm_system.jl:492-515: a species that lives on boundary regions only"""
    system._increase_num_species(int(ispec))
    if system.region_species[int(ispec) - 1].any():
        raise ValueError(f"Species {ispec} is already bulk species")
    for ireg in bregions:
        system.bregion_species[int(ispec) - 1, int(ireg) - 1] = 1
    system._version += 1
    return system


def boundary_dirichlet(system: System, ispec=None, ibc=None, v=None, *, species=1, region=1, value=0.0):
    """`boundary_dirichlet!(system, ispec, ibc, v)` src/vfvm_system.jl:854-873"""
    ispec = species if ispec is None else ispec
    ibc = region if ibc is None else ibc
    v = value if v is None else v
    system._increase_num_species(ispec)
    system.boundary_factors[ispec - 1, ibc - 1] = DIRICHLET
    system.boundary_values[ispec - 1, ibc - 1] = v
    system._version += 1


def boundary_neumann(system: System, ispec, ibc, v):
    """`boundary_neumann!(system, ispec, ibc, v)` src/vfvm_system.jl:886-890"""
    system._increase_num_species(ispec)
    system.boundary_factors[ispec - 1, ibc - 1] = 0.0
    system.boundary_values[ispec - 1, ibc - 1] = v
    system._version += 1


def boundary_robin(system: System, ispec, ibc, alpha, v):
    """`boundary_robin!(system, ispec, ibc, alpha, v)` src/vfvm_system.jl:915-919"""
    system._increase_num_species(ispec)
    system.boundary_factors[ispec - 1, ibc - 1] = alpha
    system.boundary_values[ispec - 1, ibc - 1] = v
    system._version += 1


def num_dof(system: System) -> int:
    """`num_dof(system)`: n N for dense storage (src/vfvm_system.jl:790), the number of defined dofs for sparse storage (:806)"""
    if system.unknown_storage == "sparse":
        return int(system.node_dof().sum())
    return system.num_species * system.grid.num_nodes


def unknowns(system: System, inival=None):
    """`unknowns(system; inival)` src/vfvm_system.jl:1086-1191: a dense (n, N) array, or a `SparseSolutionArray` that stores only the
    dofs of species defined at a node when the system was created with `unknown_storage="sparse"`"""
    if system.unknown_storage == "sparse":
        from .sparsesolution import SparseSolutionArray

        u = SparseSolutionArray(system.node_dof())
        if inival is not None:
            u.nzval[:] = inival
        return u
    u = np.zeros((system.num_species, system.grid.num_nodes), order="F")
    if inival is not None:
        u[...] = inival
    return u
