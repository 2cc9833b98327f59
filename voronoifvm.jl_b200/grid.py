"""Host-side grid container + tensor-product simplex grid generator.

This is the subset of ExtendableGrids (third party, not under /root/reference; Project.toml:58) the hot
path reads: `grid[Coordinates]`, `grid[CellNodes]`, `grid[CellRegions]`, `grid[BFaceNodes]`,
`grid[BFaceRegions]`, `grid[CoordinateSystem]` (src/vfvm_system.jl:691-700).  Arrays keep the Julia memory
layout (column-major, one column per node / cell / bface) but node indices are 0-based.

`simplexgrid(X[,Y[,Z]])` follows the conventions the reference's examples rely on:
nodes numbered x-fastest; every rectangle is cut into 2 triangles along its (+,+) diagonal and every cube into
the 6 Kuhn tetrahedra around its (+,+,+) diagonal; one cell region; boundary regions 1=south (y min; in 1D:
left), 2=east (x max; in 1D: right), 3=north (y max), 4=west (x min), 5=bottom (z min), 6=top (z max).
"""
from __future__ import annotations

import dataclasses
import itertools

import numpy as np

CARTESIAN = 0
CYLINDRICAL = 1
SPHERICAL = 2


@dataclasses.dataclass
class Grid:
    dim: int
    coord: np.ndarray  # (dim, N) float64, Fortran order
    cellnodes: np.ndarray  # (dim+1, C) int32, Fortran order, 0-based
    cellregions: np.ndarray  # (C,) int32, labels >= 1
    bfacenodes: np.ndarray  # (dim, NB) int32, Fortran order, 0-based
    bfaceregions: np.ndarray  # (NB,) int32, labels >= 1
    coordsys: int = CARTESIAN

    def __post_init__(self):
        self.coord = np.asfortranarray(self.coord, dtype=np.float64)
        self.cellnodes = np.asfortranarray(self.cellnodes, dtype=np.int32)
        self.cellregions = np.ascontiguousarray(self.cellregions, dtype=np.int32)
        self.bfacenodes = np.asfortranarray(self.bfacenodes, dtype=np.int32)
        self.bfaceregions = np.ascontiguousarray(self.bfaceregions, dtype=np.int32)
        assert self.coord.shape[0] == self.dim
        assert self.cellnodes.shape[0] == self.dim + 1
        assert self.bfacenodes.shape[0] == self.dim

    @property
    def num_nodes(self) -> int:
        return self.coord.shape[1]

    @property
    def num_cells(self) -> int:
        return self.cellnodes.shape[1]

    @property
    def num_bfaces(self) -> int:
        return self.bfacenodes.shape[1]

    @property
    def num_cellregions(self) -> int:
        if getattr(self, "_num_cellregions", None):
            return self._num_cellregions  # a rank-local piece keeps the region count of the whole grid
        return int(self.cellregions.max()) if self.cellregions.size else 0

    @property
    def num_bfaceregions(self) -> int:
        if getattr(self, "_num_bfaceregions", None):
            return self._num_bfaceregions
        return int(self.bfaceregions.max()) if self.bfaceregions.size else 0


def cartesian(grid: Grid) -> Grid:
    """src/vfvm_xgrid.jl:6-16 `cartesian!`"""
    grid.coordsys = CARTESIAN
    return grid


def circular_symmetric(grid: Grid) -> Grid:
    """src/vfvm_xgrid.jl:23-32 `circular_symmetric!`"""
    if grid.dim == 3:
        raise ValueError("Unable to handle circular symmetry for 3D grid")
    grid.coordsys = CYLINDRICAL
    return grid


def spherical_symmetric(grid: Grid) -> Grid:
    """src/vfvm_xgrid.jl:39-47 `spherical_symmetric!`"""
    if grid.dim != 1:
        raise ValueError(f"Unable to handle spherical symmetry for {grid.dim}D grid")
    grid.coordsys = SPHERICAL
    return grid


def simplexgrid(X, Y=None, Z=None) -> Grid:
    X = np.asarray(X, dtype=np.float64)
    if Y is None:
        return _grid1d(X)
    Y = np.asarray(Y, dtype=np.float64)
    if Z is None:
        return _grid2d(X, Y)
    return _grid3d(X, Y, np.asarray(Z, dtype=np.float64))


def _grid1d(X):
    n = X.size
    coord = X.reshape(1, n)
    i = np.arange(n - 1, dtype=np.int32)
    cellnodes = np.stack([i, i + 1])
    bfacenodes = np.array([[0, n - 1]], dtype=np.int32)
    return Grid(1, coord, cellnodes, np.ones(n - 1, np.int32), bfacenodes, np.array([1, 2], np.int32))


def _grid2d(X, Y):
    nx, ny = X.size, Y.size
    xx, yy = np.meshgrid(X, Y, indexing="xy")  # shape (ny, nx): x fastest when raveled
    coord = np.stack([xx.ravel(), yy.ravel()])
    ix, iy = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    p00 = (ix + iy * nx).ravel().astype(np.int32)
    p10, p01, p11 = p00 + 1, p00 + nx, p00 + nx + 1
    cells = np.empty((3, 2 * p00.size), np.int32)
    cells[:, 0::2] = np.stack([p00, p10, p11])
    cells[:, 1::2] = np.stack([p11, p01, p00])
    ax = np.arange(nx - 1, dtype=np.int32)
    ay = np.arange(ny - 1, dtype=np.int32)
    south = np.stack([ax, ax + 1])
    east = np.stack([ay * nx + nx - 1, (ay + 1) * nx + nx - 1])
    north = np.stack([(ny - 1) * nx + ax, (ny - 1) * nx + ax + 1])
    west = np.stack([ay * nx, (ay + 1) * nx])
    bfacenodes = np.concatenate([south, east, north, west], axis=1)
    bfaceregions = np.concatenate([np.full(nx - 1, 1), np.full(ny - 1, 2), np.full(nx - 1, 3), np.full(ny - 1, 4)])
    return Grid(2, coord, cells, np.ones(cells.shape[1], np.int32), bfacenodes, bfaceregions.astype(np.int32))


def _grid3d(X, Y, Z):
    nx, ny, nz = X.size, Y.size, Z.size
    zz, yy, xx = np.meshgrid(Z, Y, X, indexing="ij")  # shape (nz, ny, nx): x fastest
    coord = np.stack([xx.ravel(), yy.ravel(), zz.ravel()])
    iz, iy, ix = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    p0 = (ix + nx * (iy + ny * iz)).ravel().astype(np.int64)
    step = (1, nx, nx * ny)
    ncube = p0.size
    cells = np.empty((4, 6 * ncube), np.int32)
    for t, perm in enumerate(itertools.permutations(range(3))):  # Kuhn: one tetrahedron per axis order
        a = p0
        b = a + step[perm[0]]
        c = b + step[perm[1]]
        d = c + step[perm[2]]
        cells[:, t::6] = np.stack([a, b, c, d]).astype(np.int32)

    def face(fixed_axis, fixed_index, u_axis, v_axis, nu, nv):
        """2 triangles per boundary rectangle, cut along the (+,+) in-plane diagonal (matches the Kuhn split)."""
        iu, iv = np.meshgrid(np.arange(nu - 1), np.arange(nv - 1), indexing="xy")
        q0 = (fixed_index * step[fixed_axis] + iu * step[u_axis] + iv * step[v_axis]).ravel().astype(np.int64)
        q10, q01, q11 = q0 + step[u_axis], q0 + step[v_axis], q0 + step[u_axis] + step[v_axis]
        tri = np.empty((3, 2 * q0.size), np.int32)
        tri[:, 0::2] = np.stack([q0, q10, q11])
        tri[:, 1::2] = np.stack([q11, q01, q0])
        return tri

    faces = [
        (face(1, 0, 0, 2, nx, nz), 1),  # south  y = ymin
        (face(0, nx - 1, 1, 2, ny, nz), 2),  # east   x = xmax
        (face(1, ny - 1, 0, 2, nx, nz), 3),  # north  y = ymax
        (face(0, 0, 1, 2, ny, nz), 4),  # west   x = xmin
        (face(2, 0, 0, 1, nx, ny), 5),  # bottom z = zmin
        (face(2, nz - 1, 0, 1, nx, ny), 6),  # top    z = zmax
    ]
    bfacenodes = np.concatenate([f for f, _ in faces], axis=1)
    bfaceregions = np.concatenate([np.full(f.shape[1], r, np.int32) for f, r in faces])
    return Grid(3, coord, cells, np.ones(cells.shape[1], np.int32), bfacenodes, bfaceregions)


def cellmask(grid: Grid, lo, hi, region: int, tol: float = 1e-10) -> Grid:
    """ExtendableGrids `cellmask!`: cells whose nodes all lie in the box [lo, hi] get `region`."""
    lo = np.asarray(lo, dtype=np.float64).reshape(-1, 1) - tol
    hi = np.asarray(hi, dtype=np.float64).reshape(-1, 1) + tol
    inside = np.all((grid.coord >= lo) & (grid.coord <= hi), axis=0)
    sel = np.all(inside[grid.cellnodes], axis=0)
    grid.cellregions[sel] = region
    return grid
