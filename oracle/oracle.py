"""TEST INFRASTRUCTURE (oracle) -- never imported by the product package.

ctypes front end of oracle/libvfvm_oracle.so plus restatements of the reference's Newton loop
(`solve_step!`, src/vfvm_solver.jl:13-222) and transient/embedding loop (`solve_transient!`, :267-537).
The linear solve uses SciPy SuperLU as stand-in for the reference default UMFPACK (src/vfvm_solver.jl:34-41);
both are sparse direct LU, so Newton iterates agree to rounding.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class BCEntryC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("species", C.c_int32), ("region", C.c_int32), ("has_ramp", C.c_int32), ("value", C.c_double),
                ("factor", C.c_double), ("t0", C.c_double), ("t1", C.c_double), ("v0", C.c_double), ("v1", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libvfvm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("vfvm_oracle.cpp", "physics.hpp", "dual.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.vo_create.restype = C.c_void_p
        L.vo_num_nodefactors.restype = C.c_int64
        L.vo_num_edgefactors.restype = C.c_int64
        L.vo_matrix_nnz.restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class ConvergenceError(Exception):
    pass


class AssemblyError(Exception):
    pass


class OracleSystem:
    """Built from a host description with attributes grid, num_species, region_species, boundary_factors,
    boundary_values, physics_slots(), bc_entries(), nodal_source(), has_legacy_bc(), is_linear."""

    def __init__(self, desc):
        L = lib()
        self.L = L
        self.desc = desc
        g = desc.grid
        self.g = g
        self.n = int(desc.num_species)
        self.h = C.c_void_p(L.vo_create())
        self._keep = [np.ascontiguousarray(g.coord.T).ravel(), np.ascontiguousarray(g.cellnodes.T).ravel(), g.cellregions,
                      np.ascontiguousarray(g.bfacenodes.T).ravel(), g.bfaceregions]
        rc = L.vo_set_grid(self.h, g.dim, g.coordsys, g.num_nodes, g.num_cells, g.num_bfaces, _p(self._keep[0], C.c_double), _p(self._keep[1], C.c_int),
                           _p(self._keep[2], C.c_int), _p(self._keep[3], C.c_int), _p(self._keep[4], C.c_int))
        assert rc == 0
        L.vo_update_grid(self.h)
        rs = np.ascontiguousarray(desc.region_species.T, dtype=np.uint8).ravel()  # column-major n x nreg
        rc = L.vo_set_system(self.h, self.n, _p(rs, C.c_uint8))
        assert rc == 0, rc
        bs = getattr(desc, "bregion_species", None)
        if bs is not None and np.any(bs):
            b = np.ascontiguousarray(np.asarray(bs).T, dtype=np.uint8).ravel()  # column-major n x nbreg
            assert L.vo_set_boundary_species(self.h, int(np.asarray(bs).shape[1]), _p(b, C.c_uint8)) == 0
        self.push_physics()

    def push_physics(self):
        L, desc = self.L, self.desc
        for slot, pid, params in desc.physics_slots():
            params = np.ascontiguousarray(params, dtype=np.float64)
            assert L.vo_set_physics(self.h, slot, pid, _p(params, C.c_double), params.size) == 0
        tab = desc.nodal_source()
        if tab is not None:
            t = np.ascontiguousarray(tab.T).ravel()
            L.vo_set_nodal_source(self.h, _p(t, C.c_double))
        ents = desc.bc_entries()
        arr = (BCEntryC * max(1, len(ents)))()
        for i, e in enumerate(ents):
            for k, v in e.items():
                setattr(arr[i], k, v)
        L.vo_set_bc_entries(self.h, len(ents), arr)
        bf = np.ascontiguousarray(desc.boundary_factors.T).ravel()
        bv = np.ascontiguousarray(desc.boundary_values.T).ravel()
        assert L.vo_set_legacy_bc(self.h, self.g.num_bfaceregions, _p(bf, C.c_double), _p(bv, C.c_double)) == 0

    def __del__(self):
        try:
            self.L.vo_destroy(self.h)
        except Exception:
            pass

    # ---- geometry -------------------------------------------------------------------------------
    @property
    def num_edges(self):
        return self.L.vo_num_edges(self.h)

    def edgenodes(self):
        out = np.zeros(2 * self.num_edges, np.int32)
        self.L.vo_get_edgenodes(self.h, _p(out, C.c_int))
        return out.reshape(-1, 2).T  # (2, E)

    def celledges(self):
        ne = self.g.dim * (self.g.dim + 1) // 2
        out = np.zeros(ne * self.g.num_cells, np.int32)
        self.L.vo_get_celledges(self.h, _p(out, C.c_int))
        return out.reshape(-1, ne).T

    def _factors(self, which, nitems):
        nf = getattr(self.L, f"vo_num_{which}factors")(self.h)
        colptr = np.zeros(nitems + 1, np.int64)
        reg = np.zeros(nf, np.int32)
        fac = np.zeros(nf, np.float64)
        getattr(self.L, f"vo_get_{which}factors")(self.h, _p(colptr, C.c_int64), _p(reg, C.c_int), _p(fac, C.c_double))
        return colptr, reg, fac

    def nodefactors(self):
        return self._factors("node", self.g.num_nodes)

    def edgefactors(self):
        return self._factors("edge", self.num_edges)

    def bfacefactors(self):
        out = np.zeros(self.g.dim * self.g.num_bfaces)
        self.L.vo_get_bfacefactors(self.h, _p(out, C.c_double))
        return out.reshape(-1, self.g.dim).T

    # ---- assembly -------------------------------------------------------------------------------
    def assemble(self, U, UOld=None, time=0.0, tstep=math.inf, embed=0.0, nthreads=1, want_matrix=True):
        """eval_and_assemble + flush!; returns (F (n,N) F-order, scipy CSC matrix)"""
        n, N = self.n, self.g.num_nodes
        u = np.asfortranarray(U, dtype=np.float64).ravel(order="F")
        uo = u if UOld is None else np.asfortranarray(UOld, dtype=np.float64).ravel(order="F")
        F = np.zeros(n * N)
        rc = self.L.vo_assemble(self.h, _p(u, C.c_double), _p(uo, C.c_double), _p(F, C.c_double), C.c_double(time), C.c_double(tstep), C.c_double(embed), nthreads)
        if rc == -4:
            raise AssemblyError("trying to assemble NaN")
        assert rc == 0, rc
        F = F.reshape((n, N), order="F")
        if not want_matrix:
            return F, None
        return F, self.matrix()

    def matrix(self):
        nnz = self.L.vo_matrix_nnz(self.h)
        ndof = self.n * self.g.num_nodes
        colptr = np.zeros(ndof + 1, np.int64)
        rowval = np.zeros(nnz, np.int64)
        nzval = np.zeros(nnz)
        self.L.vo_get_matrix_csc(self.h, _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval, C.c_double))
        return sp.csc_matrix((nzval, rowval, colptr), shape=(ndof, ndof))

    def krylov_solve(self, b, method="bicgstab", precon="blockjacobi", reltol=1e-10, maxiters=5000, nthreads=1):
        """multithreaded CPU Krylov solve with the matrix of the last `assemble` (the stand-in for KrylovJL_CG / KrylovJL_BICGSTAB with
        Jacobi / node-block preconditioning where a sparse LU is out of reach) -> (x, iterations, true relative residual, seconds)"""
        bb = np.ascontiguousarray(np.asarray(b, dtype=np.float64).ravel(order="F"))
        x = np.zeros_like(bb)
        it, rel, sec = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
        rc = self.L.vo_krylov_solve(self.h, 1 if method == "cg" else 0, 1 if precon == "jacobi" else 2, _p(bb, C.c_double), _p(x, C.c_double), C.c_double(reltol), int(maxiters),
                                    int(nthreads), C.byref(it), C.byref(rel), C.byref(sec))
        assert rc == 0, rc
        return x, it.value, rel.value, sec.value

    def integrate(self, U, slot=1, pid=0, params=()):
        """integrate(system, F, U) src/vfvm_postprocess.jl:18-67 for a registered node function (pid 0: the identity) -> n x ncellregions"""
        u = np.ascontiguousarray(np.asarray(U, dtype=np.float64).T).ravel()
        prm = np.ascontiguousarray(params, dtype=np.float64)
        out = np.zeros(self.n * self.g.num_cellregions)
        assert self.L.vo_integrate(self.h, slot, pid, _p(prm, C.c_double), prm.size, _p(u, C.c_double), _p(out, C.c_double)) == 0
        return out.reshape((self.n, self.g.num_cellregions), order="F")

    def integrate_boundary(self, U, slot=4, pid=0, params=()):
        """integrate(system, F, U; boundary=true) src/vfvm_postprocess.jl:29-46 -> n x nbfaceregions"""
        u = np.ascontiguousarray(np.asarray(U, dtype=np.float64).T).ravel()
        prm = np.ascontiguousarray(params, dtype=np.float64)
        out = np.zeros(self.n * self.g.num_bfaceregions)
        assert self.L.vo_integrate_boundary(self.h, slot, pid, _p(prm, C.c_double), prm.size, _p(u, C.c_double), _p(out, C.c_double)) == 0
        return out.reshape((self.n, self.g.num_bfaceregions), order="F")

    def edgeintegrate(self, U, pid, params=()):
        """edgeintegrate(system, F, U) src/vfvm_postprocess.jl:109-146 for a registered flux (pid -1: the W^{1,p} integrand)"""
        u = np.ascontiguousarray(np.asarray(U, dtype=np.float64).T).ravel()
        prm = np.ascontiguousarray(params, dtype=np.float64)
        out = np.zeros(self.n * self.g.num_cellregions)
        assert self.L.vo_edgeintegrate(self.h, pid, _p(prm, C.c_double), prm.size, _p(u, C.c_double), _p(out, C.c_double)) == 0
        return out.reshape((self.n, self.g.num_cellregions), order="F")

    def edgeflux(self, U, pid, params=()):
        """flux callback of every edge (no form factor) -> (n, E); the edge loop of nodeflux, src/vfvm_postprocess.jl:191-207"""
        u = np.ascontiguousarray(np.asarray(U, dtype=np.float64).T).ravel()
        prm = np.ascontiguousarray(params, dtype=np.float64)
        out = np.zeros(self.n * self.num_edges)
        assert self.L.vo_edgeflux(self.h, pid, _p(prm, C.c_double), prm.size, _p(u, C.c_double), _p(out, C.c_double)) == 0
        return out.reshape((self.n, self.num_edges), order="F")

    def mass_matrix(self):
        """mass_matrix(state) src/vfvm_diffeq_interface.jl:60-101 -> (N, n, n) node blocks"""
        out = np.zeros(self.n * self.n * self.g.num_nodes)
        assert self.L.vo_mass_matrix(self.h, _p(out, C.c_double)) == 0
        return out.reshape((self.g.num_nodes, self.n, self.n))

    def initialize(self, U, time=0.0, embed=0.0):
        u = np.asfortranarray(U, dtype=np.float64).ravel(order="F").copy()
        self.L.vo_initialize(self.h, _p(u, C.c_double), C.c_double(time), C.c_double(embed))
        return u.reshape((self.n, self.g.num_nodes), order="F")

    # ---- solve_step! (src/vfvm_solver.jl:13-222) with a direct solver ---------------------------
    def solve_step(self, inival, oldsol=None, time=0.0, tstep=math.inf, embed=0.0, abstol=1e-10, reltol=1e-10, maxiters=100, tol_round=1e-10,
                   tol_mono=1e-3, damp_initial=1.0, damp_growth=1.2, max_round=1000, log=None):
        oldsol = np.asfortranarray(inival if oldsol is None else oldsol, dtype=np.float64)
        solution = self.initialize(oldsol.copy(), time, embed)  # :28,:31
        is_linear = bool(getattr(self.desc, "is_linear", False))
        oldnorm, converged, damp = 1.0, False, 1.0
        rnorm = 0.0
        if not is_linear:
            damp = damp_initial
            rnorm = np.abs(solution).sum()
        nround, tolx, niter = 0, 0.0, 1
        while niter <= maxiters:
            F, A = self.assemble(solution, oldsol, time, tstep, embed)
            update = spla.splu(A).solve(F.ravel(order="F"))
            solution = (solution.ravel(order="F") - damp * update).reshape(solution.shape, order="F")
            if is_linear:
                converged = True
                break
            damp = min(damp * damp_growth, 1.0)
            norm = np.abs(update).max()
            if tolx == 0.0:
                tolx = norm * reltol
            dnorm = 1.0
            rnorm_new = np.abs(solution).sum()
            if rnorm > 1.0e-50:
                dnorm = abs((rnorm - rnorm_new) / rnorm)
            nround = nround + 1 if dnorm < tol_round else 0
            if log is not None:
                log.append(norm)
            if niter > 1 and norm / oldnorm > 1.0 / tol_mono:
                converged = False
                break
            if norm < abstol or norm < tolx:
                converged = True
                break
            oldnorm, rnorm = norm, rnorm_new
            if nround > max_round:
                converged = True
                break
            niter += 1
        if not converged:
            raise ConvergenceError()
        return solution

    # ---- solve_transient! (src/vfvm_solver.jl:267-537), exceptions not handled (handle_exceptions=false) ----
    def solve_transient(self, inival, lambdas, transient=True, time=0.0, dt=0.1, dt_min=1e-3, dt_max=1.0, dt_grow=1.2, dt_decrease=0.5, du_opt=0.1,
                        du_max_factor=1.2, force_first_step=False, num_final_steps=5, **newton):
        solution = np.asfortranarray(inival, dtype=np.float64).copy()
        oldsolution = solution.copy()
        times, sols = [float(lambdas[0])], []
        dl = dt
        if transient:
            sols.append(solution.copy())
            solution = self.initialize(solution, time, float(lambdas[0]))
        else:
            solution = self.initialize(solution, time, float(lambdas[0]))
            solution = self.solve_step(solution, oldsolution, time, math.inf, float(lambdas[0]), **newton)
            oldsolution = solution.copy()
            sols.append(solution.copy())
        istep = 0
        for i in range(len(lambdas) - 1):
            dl = max(dl, dt_min)
            lam, lend = float(lambdas[i]), float(lambdas[i + 1])
            while lam < lend:
                solved, lam0, du = False, lam, 0.0
                while not solved:
                    solved, forced, errored = True, False, False
                    try:
                        lam = lam0 + dl
                        if transient:
                            sol_new = self.solve_step(solution, oldsolution, lam, dl, 0.0, **newton)
                        else:
                            sol_new = self.solve_step(solution, oldsolution, time, math.inf, lam, **newton)
                        solution = sol_new
                    except (ConvergenceError, AssemblyError):
                        raise
                    if solved:
                        du = np.abs(solution - oldsolution).max()
                        if du > du_max_factor * du_opt:
                            solved = False
                    if not solved:
                        if math.isclose(dl, dt_min, rel_tol=1.4901161193847656e-8):
                            if not (force_first_step and istep == 0):
                                raise RuntimeError("dt_min reached")
                            forced, solved = True, True
                        else:
                            dl = max(dt_min, dl * dt_decrease)
                if solved:
                    istep += 1
                    times.append(lam)
                    sols.append(solution.copy())
                    oldsolution = solution.copy()
                    steps_to_go = math.ceil((lend - lam) / dl)
                    lpredict = lend - lam
                    if steps_to_go < num_final_steps and steps_to_go > 0:
                        lpredict = (lend - lam) / steps_to_go
                    if math.isclose(dt_max, dt_min, rel_tol=1.4901161193847656e-8):
                        lpredict = dt_max
                    if lam < lend:
                        dl = min(dt_max, dl * dt_grow, dl * du_opt / (du + 1.0e-14), lpredict, lend - lam)
                        if abs(lam + dl - lend) <= max(1.0e-15, 1.0e-15 * max(abs(lam + dl), abs(lend))):
                            dl = lend - lam
                else:
                    break
        return times, sols


def fbernoulli_pm(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    bp, bm, b = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    lib().vo_fbernoulli_pm(x.size, _p(x, C.c_double), _p(bp, C.c_double), _p(bm, C.c_double), _p(b, C.c_double))
    return bp, bm, b


def fbernoulli_dual(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = [np.zeros_like(x) for _ in range(4)]
    lib().vo_fbernoulli_dual(x.size, _p(x, C.c_double), *[_p(o, C.c_double) for o in out])
    return out  # bp, dbp, bm, dbm


def cellfactors(dim, coord, coordsys=0):
    """coord: (dim, dim+1) node coordinates of one simplex -> (npar, epar)"""
    c = np.ascontiguousarray(np.asarray(coord, dtype=np.float64).T).ravel()
    nodes = np.arange(dim + 1, dtype=np.int32)
    npar, epar = np.zeros(dim + 1), np.zeros(max(1, dim * (dim + 1) // 2))
    lib().vo_cellfactors(dim, coordsys, _p(c, C.c_double), _p(nodes, C.c_int), _p(npar, C.c_double), _p(epar, C.c_double))
    return npar, epar


def bfacefactors(dim, coord, coordsys=0):
    """coord: (dim, dim) node coordinates of one boundary face of a dim-dimensional grid"""
    c = np.ascontiguousarray(np.asarray(coord, dtype=np.float64).T).ravel()
    nodes = np.arange(dim, dtype=np.int32)
    npar, epar = np.zeros(3), np.zeros(3)
    lib().vo_bfacefactors(dim, coordsys, _p(c, C.c_double), _p(nodes, C.c_int), _p(npar, C.c_double), _p(epar, C.c_double))
    return npar[:dim], epar
