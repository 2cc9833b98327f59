"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: node-owner partition, local grids, halo maps.

The data path itself has no collective in assembly; what must hold is that (1) the rank-local grids reproduce the rows of
the global operator for the owned nodes -- checked with the CPU oracle on each rank's local grid -- and (2) the halo
send/recv lists move exactly the owners' values into the halo slots (exchanged over gloo here, over NCCL on the GPUs)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_system(dim):
    import vfvm_b200 as v
    from vfvm_b200 import physics as ph

    X = np.linspace(0, 1, 9 if dim == 3 else 17)
    g = v.simplexgrid(*([X] * dim))
    if dim == 3:
        v.cellmask(g, [0, 0, 0.4], [1, 1, 0.6], 2)
    s = v.System(g, flux=ph.PowerDiffusion([1.0, 0.5], 2), reaction=ph.AffineReaction([[1.0, -1.0], [-1.0, 1.0]]), storage=ph.LinearStorage(1.0), species=[1, 2])
    v.boundary_dirichlet(s, 1, 2 * dim - 1, 0.0)
    v.boundary_dirichlet(s, 2, 2 * dim, 1.0)
    return s


def _worker(rank, world, port, dim, q, method="ranges"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from vfvm_b200 import partition as P
    from oracle import oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        s = _make_system(dim)
        n, N = s.num_species, s.grid.num_nodes
        info = P.partition_grid(s.grid, rank, world, method)
        assert info.method == method
        ls = P.local_system(s, info)
        rng = np.random.default_rng(3)
        Ug = np.asfortranarray(rng.uniform(0.1, 1.0, (n, N)))
        # ---- halo exchange over gloo with the send/recv lists the C library would use over NCCL
        Ul = np.asfortranarray(Ug[:, info.local_nodes])
        Ul[:, info.n_owned:] = -1.0  # poison the halo
        reqs, recv_bufs = [], []
        for i, qrank in enumerate(info.neighbor_ranks):
            sidx = info.send_idx[info.send_ptr[i]:info.send_ptr[i + 1]]
            sb = torch.from_numpy(np.ascontiguousarray(Ul[:, sidx].T))
            rb = torch.empty((int(info.recv_ptr[i + 1] - info.recv_ptr[i]), n), dtype=torch.float64)
            recv_bufs.append(rb)
            reqs.append(dist.isend(sb, int(qrank)))
            reqs.append(dist.irecv(rb, int(qrank)))
        for r in reqs:
            r.wait()
        for i, rb in enumerate(recv_bufs):
            Ul[:, info.n_owned + info.recv_ptr[i]:info.n_owned + info.recv_ptr[i + 1]] = rb.numpy().T
        assert np.array_equal(Ul, Ug[:, info.local_nodes]), "halo exchange did not reproduce the owners' values"
        # ---- owned rows of the local operator == rows of the global operator (oracle on both grids)
        Fg, Ag = O.OracleSystem(s).assemble(Ug, Ug, tstep=0.1)
        Fl, Al = O.OracleSystem(ls).assemble(Ul, Ul, tstep=0.1)
        Ag, Al = Ag.tocsr(), Al.tocsr()
        gl = np.asarray(info.local_nodes)
        own = np.asarray(info.owned_global)  # == lo + arange(n_owned) for contiguous ranges, the box's nodes for rcb
        assert np.array_equal(own, gl[:info.n_owned])
        for K in range(info.n_owned):
            for i in range(n):
                rl, rg = Al.getrow(K * n + i), Ag.getrow(own[K] * n + i)
                cols_g = gl[rl.indices // n] * n + rl.indices % n
                o = np.argsort(cols_g)
                assert np.array_equal(cols_g[o], rg.indices), "row pattern differs"
                np.testing.assert_allclose(rl.data[o], rg.data, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(Fl[:, :info.n_owned], Fg[:, own], rtol=1e-12, atol=1e-14)
        tot = torch.tensor([info.n_owned], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot) == N
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,method", [(2, "ranges"), (3, "ranges"), (3, "rcb")])
def test_partition_two_ranks_gloo(dim, method):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + dim + (7 if method == "rcb" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dim, q, method)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_partition_covers_all_nodes_four_ranks():
    sys.path.insert(0, ROOT)
    from vfvm_b200 import partition as P

    s = _make_system(3)
    N = s.grid.num_nodes
    seen = np.zeros(N, int)
    for r in range(4):
        info = P.partition_grid(s.grid, r, 4)
        seen[info.local_nodes[:info.n_owned]] += 1
        # send lists of r towards q must equal the halo q receives from r
        for i, qrank in enumerate(info.neighbor_ranks):
            other = P.partition_grid(s.grid, int(qrank), 4)
            j = list(other.neighbor_ranks).index(r)
            recv_global = other.local_nodes[other.n_owned + other.recv_ptr[j]:other.n_owned + other.recv_ptr[j + 1]]
            send_global = info.local_nodes[info.send_idx[info.send_ptr[i]:info.send_ptr[i + 1]]]
            assert np.array_equal(recv_global, send_global)
    assert np.all(seen == 1)


def _scrambled_system(seed=5):
    """a 2D tensor grid whose nodes are numbered at random: contiguous node ranges have no locality at all"""
    import vfvm_b200 as v
    from vfvm_b200 import physics as ph
    from vfvm_b200.grid import Grid

    X = np.linspace(0, 1, 21)
    g = v.simplexgrid(X, X)
    perm = np.random.default_rng(seed).permutation(g.num_nodes)  # new id -> old id
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    g2 = Grid(g.dim, g.coord[:, perm], inv[g.cellnodes].astype(np.int32), g.cellregions, inv[g.bfacenodes].astype(np.int32), g.bfaceregions, g.coordsys)
    s = v.System(g2, flux=ph.PowerDiffusion([1.0, 0.5], 2), reaction=ph.AffineReaction([[1.0, -1.0], [-1.0, 1.0]]), storage=ph.LinearStorage(1.0), species=[1, 2])
    v.boundary_dirichlet(s, 1, 3, 0.0)
    v.boundary_dirichlet(s, 2, 4, 1.0)
    return s


def test_rcb_partition_of_a_scrambled_numbering():
    """recursive coordinate bisection (the Metis stand-in): balanced parts, a halo of boundary-layer size where contiguous ranges
    would make almost every node a halo node, symmetric exchange lists, and owned rows of the local operator == global rows"""
    sys.path.insert(0, ROOT)
    from vfvm_b200 import partition as P
    from oracle import oracle as O

    s = _scrambled_system()
    g, n, N = s.grid, s.num_species, s.grid.num_nodes
    part = P.rcb_parts(g.coord, 3)
    assert np.bincount(part).max() - np.bincount(part).min() <= 1
    assert P.choose_parts(g, 4, "auto")[0] == "rcb"
    X = np.linspace(0, 1, 9)
    import vfvm_b200 as v
    assert P.choose_parts(v.simplexgrid(X, X, X), 2, "auto")[0] == "ranges"  # x-fastest tensor grid: slabs are the better cut
    infos = [P.partition_grid(g, r, 4, "rcb") for r in range(4)]
    ranged = [P.partition_grid(g, r, 4, "ranges") for r in range(4)]
    assert sum(i.num_halo for i in infos) * 3 < sum(i.num_halo for i in ranged)
    seen = np.zeros(N, int)
    for info in infos:
        seen[info.owned_global] += 1
        assert np.array_equal(info.owned_global, info.local_nodes[:info.n_owned])
        for i, q in enumerate(info.neighbor_ranks):
            other = infos[int(q)]
            j = list(other.neighbor_ranks).index(info.rank)
            recv_global = other.local_nodes[other.n_owned + other.recv_ptr[j]:other.n_owned + other.recv_ptr[j + 1]]
            send_global = info.local_nodes[info.send_idx[info.send_ptr[i]:info.send_ptr[i + 1]]]
            assert np.array_equal(recv_global, send_global)
    assert np.all(seen == 1)
    rng = np.random.default_rng(3)
    Ug = np.asfortranarray(rng.uniform(0.1, 1.0, (n, N)))
    Fg, Ag = O.OracleSystem(s).assemble(Ug, Ug, tstep=0.1)
    Ag = Ag.tocsr()
    info = infos[2]
    ls = P.local_system(s, info)
    Ul = np.asfortranarray(Ug[:, info.local_nodes])
    Fl, Al = O.OracleSystem(ls).assemble(Ul, Ul, tstep=0.1)
    Al = Al.tocsr()
    gl = np.asarray(info.local_nodes)
    for K in range(info.n_owned):
        for i in range(n):
            rl, rg = Al.getrow(K * n + i), Ag.getrow(gl[K] * n + i)
            cols_g = gl[rl.indices // n] * n + rl.indices % n
            o = np.argsort(cols_g)
            assert np.array_equal(cols_g[o], rg.indices), "row pattern differs"
            np.testing.assert_allclose(rl.data[o], rg.data, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(Fl[:, :info.n_owned], Fg[:, gl[:info.n_owned]], rtol=1e-11, atol=1e-13)
