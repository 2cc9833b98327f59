"""Device twin of `VoronoiFVM.SystemState` (src/vfvm_state.jl:16-157): owns a libvfvmb200 handle holding the grid,
form factors, DBSR Jacobian and the solution / residual / update vectors in HBM."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import BCEntry, check, dptr, i32ptr, i64ptr, lib
from .system import System


class SystemState:
    def __init__(self, system: System, data=None, device: int = 0, owned_nodes: int | None = None):
        if system.num_species == 0:
            raise ValueError("No species enabled.\n Call enable_species(system,species_number, list_of_regions) at least once.")  # src/vfvm_system.jl:551-553
        self.system = system
        self.data = data
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.vfvm_create(device, C.byref(h))
        if rc != _lib.OK:
            raise _lib.VfvmError(rc, "vfvm_create failed: no CUDA device / driver; the B200 path has no CPU fallback")
        self.h = h
        g = system.grid
        self.n = system.num_species
        self.N = g.num_nodes
        check(h, self.L.vfvm_set_grid(h, g.dim, g.coordsys, g.num_nodes, g.num_cells, g.num_bfaces, dptr(g.coord.ravel(order="F")),
                                      i32ptr(g.cellnodes.ravel(order="F")), i32ptr(g.cellregions), i32ptr(g.bfacenodes.ravel(order="F")),
                                      i32ptr(g.bfaceregions)))
        self.Nown = self.N
        if owned_nodes is not None:
            check(h, self.L.vfvm_set_owned_nodes(h, owned_nodes))
            self.Nown = owned_nodes
        check(h, self.L.vfvm_build_geometry(h))  # update_grid! (src/vfvm_system.jl:607-631)
        self._push_species()
        self._version = -1
        self._push_physics()
        check(h, self.L.vfvm_build_pattern(h))
        self.linear_cache = None  # (krylov, precon) currently set up
        self.history = None

    # ------------------------------------------------------------------------------------------------
    def _push_species(self):
        """species enabled per cell / boundary region (enable_species!, enable_boundary_species!): resets the device twin's physics and pattern"""
        sysm, h = self.system, self.h
        rs = np.ascontiguousarray(sysm.region_species.ravel(order="F"), dtype=np.uint8)
        check(h, self.L.vfvm_set_system(h, self.n, rs.ctypes.data_as(C.POINTER(C.c_uint8))))
        if sysm.bregion_species.any():
            bs = np.ascontiguousarray(sysm.bregion_species.ravel(order="F"), dtype=np.uint8)
            check(h, self.L.vfvm_set_boundary_species(h, sysm.bregion_species.shape[1], bs.ctypes.data_as(C.POINTER(C.c_uint8))))
        self._species_pushed = (sysm.region_species.copy(), sysm.bregion_species.copy())

    def _push_physics(self):
        sysm, h, L = self.system, self.h, self.L
        if sysm._version == self._version:
            return
        if sysm.num_species != self.n:
            raise RuntimeError("the number of species changed after the SystemState was created")
        if not (np.array_equal(sysm.region_species, self._species_pushed[0]) and np.array_equal(sysm.bregion_species, self._species_pushed[1])):
            self._push_species()  # enable_species! after the state was created: new masks, then all physics again, then a new pattern (sync)
        for slot, pid, params in sysm.physics_slots():
            params = np.ascontiguousarray(params, dtype=np.float64)
            check(h, L.vfvm_set_physics(h, slot, pid, dptr(params) if params.size else None, params.size))
        tab = sysm.nodal_source()
        if tab is not None:
            check(h, L.vfvm_set_nodal_source(h, dptr(np.ascontiguousarray(tab.ravel(order="F")))))
        ents = sysm.bc_entries()
        arr = (BCEntry * max(1, len(ents)))()
        for i, e in enumerate(ents):
            for k, v in e.items():
                setattr(arr[i], k, v)
        check(h, L.vfvm_set_bc_entries(h, len(ents), arr))
        bf = np.ascontiguousarray(sysm.boundary_factors.ravel(order="F"))
        bv = np.ascontiguousarray(sysm.boundary_values.ravel(order="F"))
        check(h, L.vfvm_set_legacy_bc(h, sysm.grid.num_bfaceregions, dptr(bf) if bf.size else None, dptr(bv) if bv.size else None))
        self._version = sysm._version

    def sync(self):
        """push changed physics parameters; rebuild the pattern if a coupling mask changed"""
        before = self._version
        self._push_physics()
        if before != self._version and before != -1:
            nr, nnz = C.c_int64(), C.c_int64()
            rc = self.L.vfvm_pattern_size(self.h, C.byref(nr), C.byref(nnz))
            if rc == _lib.ERR_STATE:
                check(self.h, self.L.vfvm_build_pattern(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.vfvm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- geometry getters (parity probes) ---------------------------------------------------------------
    @property
    def num_edges(self):
        e = C.c_int64()
        check(self.h, self.L.vfvm_num_edges(self.h, C.byref(e)))
        return e.value

    def edgenodes(self):
        out = np.zeros(2 * self.num_edges, np.int32)
        check(self.h, self.L.vfvm_get_edgenodes(self.h, i32ptr(out)))
        return out.reshape(-1, 2).T

    def celledges(self):
        g = self.system.grid
        ne = g.dim * (g.dim + 1) // 2
        out = np.zeros(ne * g.num_cells, np.int32)
        check(self.h, self.L.vfvm_get_celledges(self.h, i32ptr(out)))
        return out.reshape(-1, ne).T

    def _factors(self, which):
        nn, ne = C.c_int64(), C.c_int64()
        check(self.h, self.L.vfvm_num_factors(self.h, C.byref(nn), C.byref(ne)))
        nitems, nf = (self.N, nn.value) if which == "node" else (self.num_edges, ne.value)
        colptr, reg, fac = np.zeros(nitems + 1, np.int64), np.zeros(nf, np.int32), np.zeros(nf)
        fn = self.L.vfvm_get_nodefactors if which == "node" else self.L.vfvm_get_edgefactors
        check(self.h, fn(self.h, i64ptr(colptr), i32ptr(reg), dptr(fac)))
        return colptr, reg, fac

    def nodefactors(self):
        return self._factors("node")

    def edgefactors(self):
        return self._factors("edge")

    def bfacefactors(self):
        g = self.system.grid
        out = np.zeros(g.dim * g.num_bfaces)
        check(self.h, self.L.vfvm_get_bfacefactors(self.h, dptr(out)))
        return out.reshape(-1, g.dim).T

    # ---- vectors --------------------------------------------------------------------------------------------
    def set_vector(self, which, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))
        assert a.size == self.n * self.N
        check(self.h, self.L.vfvm_set_vector(self.h, which, a.ctypes.data, _lib.HOST))

    def get_vector(self, which):
        out = np.zeros(self.n * self.N)
        check(self.h, self.L.vfvm_get_vector(self.h, which, out.ctypes.data, _lib.HOST))
        return out.reshape((self.n, self.N), order="F")

    # ---- assembly -----------------------------------------------------------------------------------------
    def assemble(self, time=0.0, tstep=np.inf, embed=0.0):
        self.sync()
        rc = self.L.vfvm_assemble(self.h, time, tstep, embed)
        return rc

    def eval_res_jac(self, U, UOld=None, time=0.0, tstep=np.inf, embed=0.0):
        """host U -> host F through the C ABI (the Jacobian stays in HBM)"""
        self.sync()
        u = np.ascontiguousarray(np.asarray(U, dtype=np.float64).ravel(order="F"))
        uo = None if UOld is None else np.ascontiguousarray(np.asarray(UOld, dtype=np.float64).ravel(order="F"))
        F = np.zeros(self.n * self.N)
        rc = self.L.vfvm_eval_res_jac(self.h, u.ctypes.data, None if uo is None else uo.ctypes.data, F.ctypes.data, _lib.HOST, time, tstep, embed)
        check(self.h, rc)
        return F.reshape((self.n, self.N), order="F")

    def matrix(self, fmt="csc"):
        """the assembled Jacobian in the scalar pattern the reference would hold (SparseMatrixCSC(flush!(matrix)))"""
        nr, nnz = C.c_int64(), C.c_int64()
        check(self.h, self.L.vfvm_pattern_size(self.h, C.byref(nr), C.byref(nnz)))
        ptr = np.zeros((nr.value if fmt == "csr" else self.n * self.N) + 1, np.int64)
        idx = np.zeros(nnz.value, np.int64)
        val = np.zeros(nnz.value)
        if fmt == "csr":
            check(self.h, self.L.vfvm_get_pattern_csr(self.h, i64ptr(ptr), i64ptr(idx)))
            check(self.h, self.L.vfvm_get_nzval_csr(self.h, val.ctypes.data, _lib.HOST))
            return sp.csr_matrix((val, idx, ptr), shape=(nr.value, self.n * self.N))
        check(self.h, self.L.vfvm_get_pattern_csc(self.h, i64ptr(ptr), i64ptr(idx)))
        check(self.h, self.L.vfvm_get_nzval_csc(self.h, val.ctypes.data, _lib.HOST))
        return sp.csc_matrix((val, idx, ptr), shape=(nr.value, self.n * self.N))

    def rows_csr(self, node0, node1):
        """scalar CSR rows of the owned nodes [node0, node1) (local dof column numbers) without downloading the whole Jacobian"""
        nnz = C.c_int64()
        check(self.h, self.L.vfvm_get_rows_csr(self.h, node0, node1, C.byref(nnz), None, None, None))
        ptr = np.zeros((node1 - node0) * self.n + 1, np.int64)
        idx = np.zeros(nnz.value, np.int64)
        val = np.zeros(nnz.value)
        check(self.h, self.L.vfvm_get_rows_csr(self.h, node0, node1, C.byref(nnz), i64ptr(ptr), i64ptr(idx), dptr(val)))
        return sp.csr_matrix((val, idx, ptr), shape=((node1 - node0) * self.n, self.n * self.N))

    def spmv(self, x):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel(order="F"))
        y = np.zeros(self.n * self.Nown)
        check(self.h, self.L.vfvm_spmv(self.h, x.ctypes.data, y.ctypes.data, _lib.HOST))
        return y

    # ---- instrumentation --------------------------------------------------------------------------------
    def timings(self):
        t = np.zeros(_lib.NUM_TIMES)
        check(self.h, self.L.vfvm_timings(self.h, dptr(t)))
        return t

    def launch_count(self):
        n = C.c_int64()
        check(self.h, self.L.vfvm_launch_count(self.h, C.byref(n)))
        return n.value

    def device_bytes(self):
        n = C.c_int64()
        check(self.h, self.L.vfvm_device_bytes(self.h, C.byref(n)))
        return n.value

    def matrix_plane_counts(self):
        a, b = C.c_int(), C.c_int()
        check(self.h, self.L.vfvm_plane_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def block_counts(self):
        a, b = C.c_int64(), C.c_int64()
        check(self.h, self.L.vfvm_block_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value
