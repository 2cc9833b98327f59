"""Node-owner partitioning of a grid across ranks (one process per GPU) + halo maps.

The reference only partitions for shared-memory threads (ExtendableGrids `PartitionNodes/PartitionCells`,
src/vfvm_system.jl:673-682,741-748; coloured loops src/vfvm_assembly.jl:571-612; Metis via `PlainMetisPartitioning`,
examples/Example201_Laplace2D.jl:27).  Here rank p owns a contiguous range of node numbers (z-slabs on x-fastest tensor grids:
method "ranges") or -- for numberings without locality (unstructured generators, scrambled input) -- the nodes of one box of a
recursive coordinate bisection (method "rcb": the stand-in for Metis, which is not in this image); the grid is then renumbered
part by part, so that every rank again owns a contiguous range of the renumbered nodes and all maps below apply unchanged.  Its local grid holds every cell that touches an owned node, so that the
form factors of all owned nodes and of all edges with an owned end are complete; the local node numbering is
[owned nodes (ascending global)] + [halo nodes grouped by owner rank, ascending global].  Each rank assembles only its own
rows -- assembly needs no collective -- and NCCL carries the halo refresh of U / Krylov vectors and the dot-product
all-reduces (csrc/comm.cu).  Every rank derives all maps from the replicated host grid, without communication.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

from .grid import Grid
from .system import System


@dataclasses.dataclass
class PartitionInfo:
    rank: int
    nparts: int
    node_ranges: np.ndarray  # (nparts+1,) global node offsets
    local_nodes: np.ndarray  # (Nloc,) global index of each local node; first n_owned are owned
    n_owned: int
    neighbor_ranks: np.ndarray  # (nn,) int32
    send_ptr: np.ndarray  # (nn+1,) int64
    send_idx: np.ndarray  # local (owned) node indices to send, grouped by neighbour
    recv_ptr: np.ndarray  # (nn+1,) int64 offsets into the halo range
    num_bfaces: int
    grid: Grid
    method: str = "ranges"
    owned_global: np.ndarray | None = None  # global (original) ids of the owned nodes; == arange(node_ranges[rank], node_ranges[rank + 1]) for "ranges"

    @property
    def num_halo(self):
        return self.local_nodes.size - self.n_owned


def node_ranges(num_nodes: int, nparts: int) -> np.ndarray:
    return (np.arange(nparts + 1, dtype=np.int64) * num_nodes) // nparts


def rcb_parts(coord: np.ndarray, nparts: int) -> np.ndarray:
    """Recursive coordinate bisection: part id (0..nparts-1) of every node.  A box is cut perpendicular to its longest side at the
    node-count quantile that matches the split of its part count (floor/ceil halves, so any nparts works); ties are broken by node
    number, which makes the result deterministic and the part sizes equal up to one node."""
    N = coord.shape[1]
    part = np.zeros(N, dtype=np.int32)

    def split(ids, p0, np_):
        if np_ == 1:
            part[ids] = p0
            return
        nl = np_ // 2
        xyz = coord[:, ids]
        axis = int(np.argmax(xyz.max(axis=1) - xyz.min(axis=1)))
        order = np.lexsort((ids, xyz[axis]))  # by coordinate, then by node number
        k = (ids.size * nl) // np_
        split(ids[order[:k]], p0, nl)
        split(ids[order[k:]], p0 + nl, np_ - nl)

    split(np.arange(N, dtype=np.int64), 0, int(nparts))
    return part


def cut_cells(grid: Grid, part: np.ndarray) -> int:
    """number of cells whose nodes belong to more than one part (what the halo volume grows with)"""
    pc = part[grid.cellnodes]
    return int((pc != pc[0]).any(axis=0).sum())


def choose_parts(grid: Grid, nparts: int, method: str | None = None):
    """(method, part ids or None).  "ranges" (default, or VFVM_PARTITION): contiguous node ranges -- z-slabs with two neighbours per
    rank on x-fastest tensor grids, which is what the latency-bound exchanges of the Krylov iteration want; "rcb": recursive coordinate
    bisection; "auto": whichever cuts fewer cells (rcb when the numbering has no locality)."""
    method = method or os.environ.get("VFVM_PARTITION", "ranges")
    if method not in ("ranges", "rcb", "auto"):
        raise ValueError("partition method must be 'ranges', 'rcb' or 'auto'")
    if nparts <= 1 or method == "ranges":
        return "ranges", None
    part = rcb_parts(grid.coord, nparts)
    if method == "auto":
        rng = node_ranges(grid.num_nodes, nparts)
        part_r = (np.searchsorted(rng, np.arange(grid.num_nodes), side="right") - 1).astype(np.int32)
        if cut_cells(grid, part_r) <= cut_cells(grid, part):
            return "ranges", None
    return "rcb", part


def subgrid_for_node_range(grid: Grid, lo: int, hi: int, rng: np.ndarray | None = None):
    """Grid piece of the node range [lo, hi): every cell that touches one of these nodes, local numbering = [the range, ascending]
    + [the other nodes of those cells, grouped by owner (if `rng` gives owner ranges) and ascending].  The form factors of the
    range's nodes and of every edge with an end in the range are complete on the piece, so its operator rows equal the global rows
    (tests/test_partition_gloo.py).  Returns (local grid, local_nodes, selected cells, selected bfaces, halo owners)."""
    N = grid.num_nodes
    cn = grid.cellnodes
    csel = ((cn >= lo) & (cn < hi)).any(axis=0)
    lcells = cn[:, csel]
    nodes = np.unique(lcells)
    is_owned = (nodes >= lo) & (nodes < hi)
    owned = np.arange(lo, hi, dtype=np.int64)  # every node of the range is kept, even if isolated
    halo = nodes[~is_owned].astype(np.int64)
    halo_owner = np.searchsorted(rng, halo, side="right") - 1 if rng is not None else np.zeros(halo.size, np.int64)
    order = np.lexsort((halo, halo_owner))
    halo, halo_owner = halo[order], halo_owner[order]
    local_nodes = np.concatenate([owned, halo])
    g2l = np.full(N, -1, dtype=np.int64)
    g2l[local_nodes] = np.arange(local_nodes.size)
    lcellnodes = g2l[lcells].astype(np.int32)
    bsel = ((grid.bfacenodes >= lo) & (grid.bfacenodes < hi)).any(axis=0)
    lbf = g2l[grid.bfacenodes[:, bsel]]
    assert (lbf >= 0).all(), "a boundary face with an owned node must belong to a local cell"
    lgrid = Grid(grid.dim, grid.coord[:, local_nodes], lcellnodes, grid.cellregions[csel], lbf.astype(np.int32), grid.bfaceregions[bsel], grid.coordsys)
    # the grid-wide region counts must survive on every piece (physics tables are indexed by region label)
    lgrid._num_cellregions = grid.num_cellregions
    lgrid._num_bfaceregions = grid.num_bfaceregions
    return lgrid, local_nodes, lcells, g2l, halo_owner, int(bsel.sum())


def partition_grid(grid: Grid, rank: int, nparts: int, method: str | None = None) -> PartitionInfo:
    method, part = choose_parts(grid, nparts, method)
    perm = None
    if part is not None:  # renumber part by part: new id -> old id (stable, so the old order survives inside a part)
        perm = np.argsort(part, kind="stable").astype(np.int64)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        pgrid = Grid(grid.dim, grid.coord[:, perm], inv[grid.cellnodes].astype(np.int32), grid.cellregions, inv[grid.bfacenodes].astype(np.int32), grid.bfaceregions, grid.coordsys)
        pgrid._num_cellregions = grid.num_cellregions
        pgrid._num_bfaceregions = grid.num_bfaceregions
        rng = np.concatenate([[0], np.cumsum(np.bincount(part, minlength=nparts))]).astype(np.int64)
        grid = pgrid
    else:
        rng = node_ranges(grid.num_nodes, nparts)
    lo, hi = int(rng[rank]), int(rng[rank + 1])
    lgrid, local_nodes, lcells, g2l, halo_owner, nbf = subgrid_for_node_range(grid, lo, hi, rng)
    # neighbours + exchange lists
    nbr = np.unique(halo_owner).astype(np.int32)
    recv_ptr = np.concatenate([[0], np.cumsum([np.count_nonzero(halo_owner == q) for q in nbr])]).astype(np.int64)
    send_lists = []
    cell_owner = np.searchsorted(rng, lcells, side="right") - 1  # (nn, Cloc)
    mine = cell_owner == rank
    for q in nbr:
        touches_q = (cell_owner == q).any(axis=0)
        snodes = np.unique(lcells[:, touches_q][mine[:, touches_q]])
        send_lists.append(g2l[snodes].astype(np.int32))
    send_ptr = np.concatenate([[0], np.cumsum([s.size for s in send_lists])]).astype(np.int64)
    send_idx = np.concatenate(send_lists).astype(np.int32) if send_lists else np.zeros(0, np.int32)
    if perm is not None:  # back to the caller's node numbers
        local_nodes = perm[local_nodes]
    return PartitionInfo(rank, nparts, rng, local_nodes, hi - lo, nbr, send_ptr, send_idx, recv_ptr, nbf, lgrid, method, np.asarray(local_nodes[: hi - lo]))


def local_system(system: System, info: PartitionInfo) -> System:
    """the same species / physics / legacy boundary tables on the rank's local grid"""
    ls = System(info.grid, system.physics, is_linear=system.is_linear, assembly=system.assembly_type, unknown_storage=system.unknown_storage)
    ls._increase_num_species(system.num_species)
    ls.region_species[:, :] = system.region_species
    ls.bregion_species[:, :] = system.bregion_species
    ls.boundary_factors[:, :] = system.boundary_factors
    ls.boundary_values[:, :] = system.boundary_values
    ls._version += 1
    return ls


def partitioned_state(system: System, rank: int, world: int, device: int, method: str | None = None):
    """device twin of rank `rank`: local grid, owned rows, NCCL communicator (id shared through torch.distributed)"""
    import ctypes as C

    import torch.distributed as dist

    from . import _lib
    from .state import SystemState

    info = partition_grid(system.grid, rank, world, method)
    ls = local_system(system, info)
    st = SystemState(ls, device=device, owned_nodes=info.n_owned)
    L = st.L
    uid = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(st.h, L.vfvm_comm_unique_id(uid))
    box = [uid.raw]
    dist.broadcast_object_list(box, src=0)
    _lib.check(st.h, L.vfvm_comm_init(st.h, rank, world, box[0]))
    nb = np.ascontiguousarray(info.neighbor_ranks, dtype=np.int32)
    _lib.check(st.h, L.vfvm_set_halo(st.h, nb.size, _lib.i32ptr(nb), _lib.i64ptr(info.send_ptr), _lib.i32ptr(np.ascontiguousarray(info.send_idx, dtype=np.int32)),
                                     _lib.i64ptr(info.recv_ptr)))
    # peer mailboxes: halo exchange fused into the SpMV kernel, reductions into their finalize kernel (csrc/peer.cuh)
    st.peer = False
    if world <= 8 and not os.environ.get("VFVM_NO_PEER"):
        hbuf = C.create_string_buffer(64)
        _lib.check(st.h, L.vfvm_peer_export(st.h, hbuf))
        handles = [None] * world
        dist.all_gather_object(handles, hbuf.raw)
        rc = L.vfvm_peer_connect(st.h, b"".join(handles))
        oks = [None] * world
        dist.all_gather_object(oks, rc == 0)  # all ranks or none: the transports must match
        if not all(oks):
            os.environ["VFVM_NO_PEER"] = "1"
            if rc == 0:
                L.vfvm_peer_connect(st.h, b"".join(handles))  # with VFVM_NO_PEER set this switches the transport back to NCCL
        st.peer = bool(L.vfvm_peer_active(st.h))
    st.partition = info
    return st, info
