"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/vfvm_b200.h declares; the host
mirror rejects unregistered callbacks; grid generator invariants; the product package never touches the oracle."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import vfvm_b200 as v
from vfvm_b200 import physics as ph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vfvm_b200.h")).read()
    return sorted(set(re.findall(r"\b(vfvm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(v._lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/vfvm_b200.h but not exported"
    # and the ctypes binding covers exactly the same set
    bound = set(v._lib.SIGNATURES) | set(v._lib.OTHER)
    assert bound == set(names), (bound ^ set(names))


def test_no_cpu_fallback_without_gpu():
    """creating a device state must fail loudly when there is no CUDA device (this container has none)"""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import ctypes as C; import vfvm_b200 as v; L = v._lib.lib(); h = C.c_void_p(); "
            "rc = L.vfvm_create(0, C.byref(h)); print(rc)" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.stdout.strip() == str(v._lib.ERR_CUDA), out
    X = np.linspace(0, 1, 4)
    sys_ = v.System(v.simplexgrid(X, X), flux=ph.LinearDiffusion(), species=[1])
    code2 = ("import sys; sys.path.insert(0, %r); import numpy as np; import vfvm_b200 as v; from vfvm_b200 import physics as ph; X=np.linspace(0,1,4); "
             "s=v.System(v.simplexgrid(X,X), flux=ph.LinearDiffusion(), species=[1]); v.SystemState(s)" % ROOT)
    out = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, env=env)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr


def test_unregistered_callback_raises():
    X = np.linspace(0, 1, 4)
    g = v.simplexgrid(X, X)

    def flux(f, u, edge, data):  # an arbitrary host closure, as a reference user would write it
        f[0] = u[0, 0] - u[0, 1]

    with pytest.raises(v.UnregisteredPhysicsError):
        v.System(g, flux=flux)
    with pytest.raises(v.UnregisteredPhysicsError):
        v.System(g, flux=ph.LinearStorage())  # registered, but not a flux
    with pytest.raises(NotImplementedError):
        v.System(g, flux=ph.LinearDiffusion(), bflux=ph.LinearDiffusion())


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "voronoifvm.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.lower().replace("no cpu oracle", ""), f"{fn} mentions the oracle"


@pytest.mark.parametrize("dim,nx", [(1, 7), (2, 5), (3, 4)])
def test_simplexgrid_counts(dim, nx):
    X = np.linspace(0, 1, nx)
    g = v.simplexgrid(*([X] * dim))
    m = nx - 1
    assert g.num_nodes == nx**dim
    assert g.num_cells == {1: m, 2: 2 * m * m, 3: 6 * m**3}[dim]
    assert g.num_bfaces == {1: 2, 2: 4 * m, 3: 12 * m * m}[dim]
    assert g.num_bfaceregions == 2 * dim
    # every cell has positive volume and the volumes add up to 1
    P = g.coord[:, g.cellnodes]  # (dim, dim+1, C)
    E = P[:, 1:, :] - P[:, :1, :]
    vol = np.abs(np.linalg.det(np.moveaxis(E, 2, 0))) / {1: 1, 2: 2, 3: 6}[dim]
    assert vol.min() > 0 and vol.sum() == pytest.approx(1.0)
    # E = 3m nx^2 + 3 m^2 nx + m^3 in 3D
    if dim == 3:
        pairs = set()
        for a in range(4):
            for b in range(a + 1, 4):
                lo = np.minimum(g.cellnodes[a], g.cellnodes[b]).astype(np.int64)
                hi = np.maximum(g.cellnodes[a], g.cellnodes[b]).astype(np.int64)
                pairs.update((lo * g.num_nodes + hi).tolist())
        assert len(pairs) == 3 * m * nx * nx + 3 * m * m * nx + m**3


def test_system_mirror_bookkeeping():
    X = np.linspace(0, 1, 4)
    s = v.System(v.simplexgrid(X, X), flux=ph.LinearDiffusion())
    v.enable_species(s, 2, [1])
    assert s.num_species == 2 and s.region_species[1, 0] == 1 and s.region_species[0, 0] == 0
    v.boundary_dirichlet(s, 1, 3, 1.0)
    assert s.boundary_factors[0, 2] == v.DIRICHLET and s.boundary_values[0, 2] == 1.0
    v.boundary_robin(s, 2, 1, 0.5, 2.0)
    assert s.boundary_factors[1, 0] == 0.5 and s.has_legacy_bc()
    u = v.unknowns(s, 0.5)
    assert u.shape == (2, 16) and u.flags.f_contiguous and v.num_dof(s) == 32


def test_golden_file_is_consistent():
    """tests/golden/reference_known_answers.json: every value is asserted by the oracle tests, and (build container only) the literal
    stands at the cited line of the reference"""
    import json

    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "reference_known_answers.json")))
    tests_text = open(os.path.join(here, "test_oracle_golden.py")).read()
    for key, g in gold.items():
        assert g["literal"] in tests_text, key
        assert float(g["literal"]) == g["value"]
        path, line = g["reference"].rsplit(":", 1)
        full = os.path.join("/root/reference", path)
        if os.path.exists(full):
            assert g["literal"] in open(full).read().splitlines()[int(line) - 1], key


def test_postprocess_rejects_unregistered_functions_before_touching_the_device():
    """integrate / edgeintegrate take registered physics objects only (no silent CPU evaluation of arbitrary callbacks)"""
    X = np.linspace(0, 1, 5)
    s = v.System(v.simplexgrid(X, X), species=[1])
    U = np.zeros((1, s.grid.num_nodes))
    with pytest.raises(v.UnregisteredPhysicsError):
        v.integrate(s, lambda y, u, node, data=None: None, U)
    with pytest.raises(v.UnregisteredPhysicsError):
        v.edgeintegrate(s, lambda y, u, edge, data=None: None, U)
    with pytest.raises(v.UnregisteredPhysicsError):
        v.integrate(s, ph.LinearDiffusion(), U)  # a flux is not a node function


def test_region_affine_reaction_parameter_block():
    r = ph.RegionAffineReaction([np.eye(2), 2 * np.eye(2)], [None, [1.0, 2.0]])
    p = r.params(2)
    assert p[0] == 2 and p.size == 1 + 2 * (4 + 2)
    assert np.array_equal(p[1:5], [1, 0, 0, 1]) and np.array_equal(p[5:7], [0, 0]) and np.array_equal(p[11:13], [1, 2])


def test_amg_builder_options_vector():
    b = v.AMGPreconBuilder(wdepth=2, alpha=2.0)
    assert b.precon == v._lib.PRECON_AMG and len(b.options) == 6
    assert b.options[1] == 2.0 and b.options[5] == 2.0 and all(o != o for o in (b.options[0], b.options[2], b.options[3], b.options[4]))
    assert v.SmoothedAggregationPreconBuilder is v.AMGPreconBuilder


def test_sparse_unknown_storage_host_mirror():
    """unknown_storage = :sparse (src/vfvm_system.jl:813-825, src/vfvm_sparsesolution.jl): only dofs of species defined at a node are
    stored, column by column; undefined dofs read as NaN and ignore assignments; dense() / from_dense() are the device conversions"""
    import vfvm_b200 as v
    from vfvm_b200 import physics as ph
    from vfvm_b200.sparsesolution import SparseSolutionArray

    X = np.linspace(0, 1, 5)
    g = v.simplexgrid(X, X)
    v.cellmask(g, [0.0, 0.0], [0.5, 1.0], 2)
    s = v.System(g, flux=ph.LinearDiffusion([1.0, 1.0, 1.0]), unknown_storage="sparse")
    v.enable_species(s, 1, [1, 2])
    v.enable_species(s, 2, [2])
    v.enable_boundary_species(s, 3, [2])
    mask = s.node_dof()
    u = v.unknowns(s, inival=0.5)
    assert isinstance(u, SparseSolutionArray) and u.shape == (3, g.num_nodes)
    assert len(u) == v.num_dof(s) == int(mask.sum()) < 3 * g.num_nodes
    K_in = int(np.nonzero(mask[1])[0][0])
    K_out = int(np.nonzero(~mask[1])[0][0])
    assert u[1, K_in] == 0.5 and math.isnan(u[1, K_out]) and u.dof(1, K_out) == -1
    u[1, K_out] = 7.0  # ignored
    u[1, K_in] = 2.0
    d = u.dense()
    assert d.flags.f_contiguous and d[1, K_in] == 2.0 and d[1, K_out] == 0.0 and np.all(d[~mask] == 0.0)
    # CSC layout of node_dof: species ascending inside a node
    for K in (K_in, K_out):
        sp = u.rowval[u.colptr[K]:u.colptr[K + 1]]
        assert np.array_equal(sp, np.nonzero(mask[:, K])[0])
    rng = np.random.default_rng(0)
    D = np.asfortranarray(rng.uniform(size=mask.shape)) * mask
    w = SparseSolutionArray.from_dense(mask, D)
    assert np.array_equal(w.dense(), D)
    assert np.array_equal((w + w).dense(), 2 * D) and np.array_equal((w - w).dofs(), np.zeros(len(w)))
    c = w.copy()
    c[0, 0] = -1.0
    assert w[0, 0] != -1.0 and w.similar().shape == w.shape
    # dense systems keep the dense array
    s2 = v.System(g, flux=ph.LinearDiffusion(), species=[1])
    assert isinstance(v.unknowns(s2), np.ndarray) and v.num_dof(s2) == g.num_nodes


def test_bench_byte_formulas_match_the_survey():
    """bench.py's algorithmic bytes are SURVEY.md section 8d's formulas on the stored coupling planes: 3D tensor grid, one species -> 52.6-52.9 B
    per edge; a CG iteration with an AMG V-cycle = 3 level-0 SpMVs + the coarser levels' share + vector streams"""
    import bench

    nx = 193
    m = nx - 1
    N, E, NB = nx**3, 3 * m * nx * nx + 3 * m * m * nx + m**3, 12 * m * m
    assert E == 49877568
    b = bench.algorithmic_bytes(1, N, E, NB, 3, 1, 1, False)
    assert 52.5 < b / E < 53.0  # SURVEY: 52.6 B/edge for n = 1 (+ the boundary items)
    b3 = bench.algorithmic_bytes(3, N, E, NB, 3, 9, 9, True)
    assert 190.0 < b3 / E < 200.0  # SURVEY: 194 B/edge for n = 3, fully coupled
    nnz = 2 * E
    spmv = nnz * 12 + N * 24
    it_v = bench.iteration_bytes(1, N, nnz, 1, 1, "cg", 1.1)
    assert it_v == pytest.approx(spmv + 1.1 * (2 * spmv + 7 * 8 * N) + 6 * 8 * N)
    assert bench.iteration_bytes(1, N, nnz, 1, 1, "cg", 1.22) > it_v > bench.iteration_bytes(1, N, nnz, 1, 1, "cg", 0)
    assert bench.iteration_bytes(1, N, nnz, 1, 1, "bicgstab", 1.1) == pytest.approx(2 * spmv + 2 * 1.1 * (2 * spmv + 7 * 8 * N) + 10 * 8 * N)
