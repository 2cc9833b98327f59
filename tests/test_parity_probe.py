"""CPU test of the full-size parity probe (tests/parity_probe.py): with the global oracle standing in for the device state, the rows
of the probe ranges assembled on their grid pieces must equal the global rows -- pattern bit-exact, values to rounding."""
import numpy as np
import scipy.sparse as sp

import vfvm_b200 as v
from vfvm_b200 import _lib
from vfvm_b200 import physics as ph
from oracle import oracle as O

from parity_probe import probe_rows, probe_ranges


class _OracleBackedState:
    """what probe_rows needs from a SystemState, answered by the oracle on the whole grid"""

    def __init__(self, system, U, tstep):
        self.n = system.num_species
        self.F, A = O.OracleSystem(system).assemble(U, U, tstep=tstep)
        self.A = A.tocsr()

    def get_vector(self, which):
        assert which == _lib.VEC_RESIDUAL
        return self.F

    def rows_csr(self, a, b):
        return self.A[a * self.n : b * self.n]


def _system(nx=9):
    X = np.linspace(0, 1, nx)
    g = v.simplexgrid(X, X, X)
    v.cellmask(g, [0, 0, 0.5], [1, 1, 1.0], 2)
    s = v.System(g, flux=ph.PowerDiffusion([1.0e-1, 2.0e-1], 2), reaction=ph.PowerReaction([1.0, 0.5], [2.0, 1.0]), storage=ph.LinearStorage([1.0, 2.0]), species=[1, 2])
    v.boundary_dirichlet(s, 1, 5, 0.1)
    v.boundary_dirichlet(s, 2, 6, 0.2)
    return s


def test_probe_ranges_cover_first_middle_last_planes():
    s = _system()
    r = probe_ranges(s, 0, s.grid.num_nodes)
    assert r[0][0] == 0 and r[-1][1] == s.grid.num_nodes and len(r) == 3
    assert all(b - a == 2 * 81 for a, b in r)


def test_probe_accepts_identical_rows_and_detects_a_wrong_entry():
    s = _system()
    U = np.asfortranarray(np.random.default_rng(3).uniform(0.1, 1.0, (2, s.grid.num_nodes)))
    st = _OracleBackedState(s, U, 0.1)
    res = probe_rows(s, st, None, U, tstep=0.1)
    assert res["ok"] and res["pattern_equal"] and res["entries"] > 1000 and res["max_rel_err_entry"] < 1e-13
    # a single perturbed Jacobian entry in a probed row must be flagged
    A = st.A.copy().tolil()
    r0 = probe_ranges(s, 0, s.grid.num_nodes)[1][0] * 2 + 1
    c0 = A.rows[r0][0]
    A[r0, c0] = A[r0, c0] * (1 + 1e-9) + 1e-9
    st.A = sp.csr_matrix(A)
    bad = probe_rows(s, st, None, U, tstep=0.1)
    assert not bad["ok"] and bad["max_err_over_bound"] > 1.0
    # an explicit zero the oracle does not hold is tolerated (value-independent device pattern), a non-zero extra entry is not
    good = _OracleBackedState(s, U, 0.1).A
    A = good.copy().tolil()
    free = [c for c in range(A.shape[1]) if c not in A.rows[r0]][0]
    A[r0, free] = 1.0
    A = sp.csr_matrix(A)
    A.data[A.indptr[r0] + list(A.indices[A.indptr[r0]:A.indptr[r0 + 1]]).index(free)] = 0.0
    st.A = A
    res0 = probe_rows(s, st, None, U, tstep=0.1)
    assert res0["ok"] and not res0["pattern_equal"] and res0["explicit_zero_extras"] == 1
    A.data[A.indptr[r0] + list(A.indices[A.indptr[r0]:A.indptr[r0 + 1]]).index(free)] = 1e-20  # rounding-level: tolerated
    assert probe_rows(s, st, None, U, tstep=0.1)["ok"]
    A.data[A.indptr[r0] + list(A.indices[A.indptr[r0]:A.indptr[r0 + 1]]).index(free)] = 1e-9
    assert not probe_rows(s, st, None, U, tstep=0.1)["ok"]
    st.A = good
    st.F = st.F.copy()
    st.F[1, 5] += 1e-6
    assert not probe_rows(s, st, None, U, tstep=0.1)["ok"]
