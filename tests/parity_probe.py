"""TEST INFRASTRUCTURE: parity probe of an assembled device state at full problem size (used by bench.py's `parity` object,
tests/test_gpu_large.py and tests/mgpu_check.py).

A rank cannot afford the CPU oracle on a 193^3 grid inside a bench run, but it can on a few node planes: the operator rows of a node
range only depend on the cells touching that range (tests/test_partition_gloo.py: "local rows == global rows"), so the oracle is run
on that grid piece and every residual entry and every Jacobian entry of the range's rows is compared with what the device holds --
pattern bit-exact, values to the same bound as tests/test_gpu_parity.py (1e-12 relative + a few ulps of the summed row terms).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

import vfvm_b200 as v
from vfvm_b200 import partition as P
from oracle import oracle as O

RTOL_ASM = 1.0e-12
EPS = float(np.finfo(float).eps)


def _piece_system(system, lgrid):
    ls = v.System(lgrid, system.physics, is_linear=system.is_linear, assembly=system.assembly_type, unknown_storage=system.unknown_storage)
    ls._increase_num_species(system.num_species)
    ls.region_species[:, :] = system.region_species
    ls.bregion_species[:, :] = system.bregion_species
    ls.boundary_factors[:, :] = system.boundary_factors
    ls.boundary_values[:, :] = system.boundary_values
    ls._version += 1
    return ls


def probe_ranges(system, lo, hi, nprobes=3):
    """node ranges inside the owned range [lo, hi): its first, middle and last two grid planes (tensor grids number x fastest)"""
    g = system.grid
    N = g.num_nodes
    plane = max(1, int(round(N ** ((g.dim - 1) / g.dim)))) if g.dim > 1 else 1
    m = min(2 * plane, max(1, (hi - lo) // 3))
    starts = [lo, lo + (hi - lo - m) // 2, hi - m][:nprobes] if hi - lo >= 3 * m else [lo]
    return [(int(a), int(min(a + m, hi))) for a in starts]


def probe_rows(system, st, pinfo, Uglob, UOldglob=None, time=0.0, tstep=math.inf, embed=0.0, ranges=None):
    """Compares the device's residual + Jacobian rows (already assembled at Uglob restricted to the rank) with the oracle on the
    probe ranges.  Returns {"rows", "entries", "pattern_equal", "max_rel_err_entry", "max_err_over_bound", "residual_max_err_over_bound"}."""
    from vfvm_b200 import _lib

    g = system.grid
    n = system.num_species
    lo, hi = (0, g.num_nodes) if pinfo is None else (int(pinfo.node_ranges[pinfo.rank]), int(pinfo.node_ranges[pinfo.rank + 1]))
    local_nodes = np.arange(g.num_nodes, dtype=np.int64) if pinfo is None else pinfo.local_nodes
    Fdev = st.get_vector(_lib.VEC_RESIDUAL)
    out = {"rows": 0, "entries": 0, "pattern_equal": True, "max_rel_err_entry": 0.0, "max_err_over_bound": 0.0, "residual_max_err_over_bound": 0.0}
    for a, b in (ranges or probe_ranges(system, lo, hi)):
        lgrid, lnodes, _, _, _, _ = P.subgrid_for_node_range(g, a, b)
        o = O.OracleSystem(_piece_system(system, lgrid))
        Ul = np.asfortranarray(Uglob[:, lnodes])
        Uol = Ul if UOldglob is None else np.asfortranarray(UOldglob[:, lnodes])
        Fo, Ao = o.assemble(Ul, Uol, time=time, tstep=tstep, embed=embed)
        m = b - a
        Ao = Ao.tocsr()[: m * n]
        Ao.sort_indices()
        # device rows, columns in global dof numbers
        Ad = st.rows_csr(a - lo, b - lo)
        cols_dev = local_nodes[Ad.indices // n] * n + Ad.indices % n
        cols_or = lnodes[Ao.indices // n] * n + Ao.indices % n
        Adg = sp.csr_matrix((Ad.data, cols_dev, Ad.indptr), shape=(m * n, n * g.num_nodes))
        Aog = sp.csr_matrix((Ao.data, cols_or, Ao.indptr), shape=(m * n, n * g.num_nodes))
        Adg.sort_indices()
        Aog.sort_indices()
        same = np.array_equal(Adg.indptr, Aog.indptr) and np.array_equal(Adg.indices, Aog.indices)
        out["pattern_equal"] = bool(out["pattern_equal"] and same)
        out["rows"] += m * n
        out["entries"] += int(Aog.nnz)
        if not same:
            # The reference's pattern is value dependent (_addnz skips exact zeros, src/vfvm_assembly.jl:21-28), the device pattern
            # is not: device entries the oracle does not hold are tolerated only if they are exactly zero; nothing may be missing.
            ncol = n * g.num_nodes
            kd = np.repeat(np.arange(m * n, dtype=np.int64), np.diff(Adg.indptr)) * ncol + Adg.indices
            ko = np.repeat(np.arange(m * n, dtype=np.int64), np.diff(Aog.indptr)) * ncol + Aog.indices
            pos = np.searchsorted(kd, ko)
            found = (pos < kd.size) & (kd[np.minimum(pos, kd.size - 1)] == ko)
            extra = np.ones(kd.size, bool)
            extra[pos[found]] = False
            # an entry the reference skipped because its value was EXACTLY zero may come out of the device's (FMA-contracted) arithmetic
            # as a rounding-level number: tolerated up to the same absolute allowance as a cancelling sum, 8 eps sum|row terms|
            rows_d = np.repeat(np.arange(m * n, dtype=np.int64), np.diff(Adg.indptr))
            rows_o = np.repeat(np.arange(m * n, dtype=np.int64), np.diff(Aog.indptr))
            tscale = np.bincount(rows_o, weights=np.where(np.abs(Aog.data) < 1e29, np.abs(Aog.data), 0.0), minlength=m * n)
            if not found.all() or np.any(np.abs(Adg.data[extra]) > 8 * EPS * tscale[rows_d[extra]]):
                out["pattern_superset_with_zero_extras"] = False
                out["max_rel_err_entry"] = out["max_err_over_bound"] = float("inf")
                continue
            out["explicit_zero_extras"] = out.get("explicit_zero_extras", 0) + int(extra.sum())
            out["max_abs_extra"] = max(out.get("max_abs_extra", 0.0), float(np.abs(Adg.data[extra]).max()) if extra.any() else 0.0)
            keep = ~extra
            counts = np.bincount(np.repeat(np.arange(m * n), np.diff(Adg.indptr))[keep], minlength=m * n)
            Adg = sp.csr_matrix((Adg.data[keep], Adg.indices[keep], np.concatenate([[0], np.cumsum(counts)])), shape=Adg.shape)
        rows = np.repeat(np.arange(m * n), np.diff(Aog.indptr))
        mag = np.where(np.abs(Aog.data) < 1e29, np.abs(Aog.data), 0.0)
        termscale = np.bincount(rows, weights=mag, minlength=m * n)
        err = np.abs(Adg.data - Aog.data)
        bound = RTOL_ASM * np.abs(Aog.data) + 8 * EPS * termscale[rows]
        out["max_rel_err_entry"] = max(out["max_rel_err_entry"], float((err / np.maximum(np.abs(Aog.data), 1e-300)).max()))
        out["max_err_over_bound"] = max(out["max_err_over_bound"], float((err / np.maximum(bound, 1e-300)).max()))
        f = Fdev[:, a - lo : b - lo].ravel(order="F")
        fo = Fo[:, :m].ravel(order="F")
        fbound = RTOL_ASM * np.abs(fo) + 8 * EPS * (termscale * max(1.0, float(np.abs(Ul).max())) + np.abs(fo))
        out["residual_max_err_over_bound"] = max(out["residual_max_err_over_bound"], float((np.abs(f - fo) / np.maximum(fbound, 1e-300)).max()))
    out["ok"] = bool((out["pattern_equal"] or out.get("pattern_superset_with_zero_extras", True)) and out["max_err_over_bound"] <= 1.0 and out["residual_max_err_over_bound"] <= 1.0)
    return out
