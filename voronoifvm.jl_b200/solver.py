"""Host mirror of the Newton / transient drivers of the reference, running on device-resident vectors.

  SolverControl                 src/vfvm_solvercontrol.jl:13-270
  NewtonSolverHistory           src/vfvm_history.jl:10-37
  solve_step!                   src/vfvm_solver.jl:13-222     (loop stays on the host, as in north_star item 4)
  solve_transient!              src/vfvm_solver.jl:267-537
  solve! / solve                src/vfvm_solver.jl:551-668
  evaluate_residual_and_jacobian src/vfvm_solver.jl:224-260
  exceptions                    src/vfvm_logging_exceptions.jl:5-29
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import time as _time
from typing import Callable

import numpy as np

from . import _lib
from ._lib import check
from .state import SystemState
from .system import System, unknowns


class ConvergenceError(Exception):
    """src/vfvm_logging_exceptions.jl:5-9"""


class AssemblyError(Exception):
    """src/vfvm_logging_exceptions.jl:11-15"""


class LinearSolverError(Exception):
    """src/vfvm_logging_exceptions.jl:17-21"""


class EmbeddingError(Exception):
    """src/vfvm_logging_exceptions.jl:23-29"""


# ---- LinearSolve.jl algorithm stand-ins accepted in SolverControl.method_linear --------------------------------
class JacobiPreconBuilder:
    precon = _lib.PRECON_JACOBI


class BlockPreconBuilder:
    """node-block Jacobi (the `BlockPreconBuilder` use with per-node blocks)"""
    precon = _lib.PRECON_BLOCKJACOBI


class AMGPreconBuilder:
    """aggregation AMG V-cycle on the node-block matrix (csrc/amg.cu); the reference reaches AMG through
    `precs = AMGPreconBuilder()` (AMGCLWrap) or `SmoothedAggregationPreconBuilder()` (AlgebraicMultigrid),
    examples/DevEx003_Solvers.jl:149-169, examples/DevEx004_EquationBlock3D.jl:225-251"""
    precon = _lib.PRECON_AMG

    def __init__(self, omega=None, alpha=None, theta=None, sweeps=None, coarse_sweeps=None, wdepth=None):
        nan = float("nan")
        self.options = [nan if o is None else float(o) for o in (omega, alpha, theta, sweeps, coarse_sweeps, wdepth)]


SmoothedAggregationPreconBuilder = AMGPreconBuilder  # same entry point; the device hierarchy uses plain aggregation with an over-weighted correction


class ILUZeroPreconBuilder:
    """ILU(0) on the node-block pattern; `multicolor=True` eliminates in multicolour order (few, wide levels)"""

    def __init__(self, multicolor: bool = False):
        self.precon = _lib.PRECON_ILU0_MC if multicolor else _lib.PRECON_ILU0


class _Krylov:
    krylov = _lib.KRYLOV_BICGSTAB

    def __init__(self, precs=None, restart=30):
        self.precs = precs
        self.restart = restart

    @property
    def precon(self):
        return _lib.PRECON_NONE if self.precs is None else self.precs.precon


class KrylovJL_BICGSTAB(_Krylov):
    krylov = _lib.KRYLOV_BICGSTAB


class KrylovJL_CG(_Krylov):
    krylov = _lib.KRYLOV_CG


class KrylovJL_GMRES(_Krylov):
    krylov = _lib.KRYLOV_GMRES


class DeviceDirectLike(_Krylov):
    """Stand-in for the reference default `UMFPACKFactorization()` (src/vfvm_solver.jl:34-41): the device has no sparse LU,
    so the default is BiCGStab + node-block Jacobi driven to (near) machine precision, which makes the Newton iterates
    agree with a direct solve to ~1e-12."""
    krylov = _lib.KRYLOV_BICGSTAB

    def __init__(self, precs=None, reltol=1.0e-13, abstol=0.0, maxiters=20000):
        super().__init__(precs if precs is not None else BlockPreconBuilder())
        self.reltol, self.abstol, self.maxiters = reltol, abstol, maxiters


def _no_pre(sol, t):
    return None


def _no_post(sol, oldsol, t, dt):
    return None


def _no_sample(sol, t):
    return None


@dataclasses.dataclass
class SolverControl:
    verbose: bool | str = False
    abstol: float = 1.0e-10
    reltol: float = 1.0e-10
    maxiters: int = 100
    tol_round: float = 1.0e-10
    tol_mono: float = 1.0e-3
    damp_initial: float = 1.0
    damp_growth: float = 1.2
    max_round: int = 1000
    updatecontrol: bool = True
    method_linear: object = None
    reltol_linear: float = 1.0e-4
    abstol_linear: float = 1.0e-8
    maxiters_linear: int = 100
    factorize_every_newtonstep: bool = False
    factorize_every_timestep: int = 1
    Δp: float = 1.0
    Δp_max: float = 1.0
    Δp_min: float = 1.0e-3
    Δp_grow: float = 1.0
    Δp_decrease: float = 0.5
    Δt: float = 0.1
    Δt_max: float = 1.0
    Δt_min: float = 1.0e-3
    Δt_grow: float = 1.2
    Δt_decrease: float = 0.5
    Δu_opt: float = 0.1
    Δu_max_factor: float = 1.2
    force_first_step: bool = False
    num_final_steps: int = 5
    handle_exceptions: bool = False
    store_all: bool = True
    log: bool = False
    pre: Callable = _no_pre
    post: Callable = _no_post
    sample: Callable = _no_sample


def fixed_timesteps(control: SolverControl, dt: float, grow: float = 1.0) -> SolverControl:
    """`fixed_timesteps!` src/vfvm_solvercontrol.jl:320-327"""
    control.Δt = control.Δt_max = control.Δt_min = dt
    control.Δt_grow = grow
    control.Δu_opt = float(np.finfo(np.float64).max)
    return control


@dataclasses.dataclass
class NewtonSolverHistory:
    nlu: int = 0
    nlin: int = 0
    time: float = 0.0
    tasm: float = 0.0
    tlinsolve: float = 0.0
    tlinsolve_setup: float = 0.0
    tlinsolve_solve: float = 0.0
    linres: float = 0.0  # residual norm of the last linear solve
    updatenorm: list = dataclasses.field(default_factory=list)
    l1normdiff: list = dataclasses.field(default_factory=list)

    def __len__(self):
        return len(self.updatenorm)


class TransientSolution:
    """minimal `TransientSolution` (src/vfvm_transientsolution.jl:31-89): lists of times and solutions"""

    def __init__(self, t0, u0):
        self.t = [float(t0)]
        self.u = [u0]
        self.history = []

    def append(self, t, u):
        self.t.append(float(t))
        self.u.append(u)

    def __len__(self):
        return len(self.t)


def _linear_setup(state: SystemState, control: SolverControl):
    m = control.method_linear
    if m is None:
        m = DeviceDirectLike()
    if not isinstance(m, _Krylov):
        raise TypeError("method_linear must be one of the device Krylov stand-ins (KrylovJL_BICGSTAB/CG/GMRES) or None")
    key = (m.krylov, m.precon, getattr(m, "restart", 30), repr(getattr(getattr(m, "precs", None), "options", None)))
    if state.linear_cache != key:
        check(state.h, state.L.vfvm_linsolve_setup(state.h, m.krylov, m.precon, getattr(m, "restart", 30)))
        opts = getattr(getattr(m, "precs", None), "options", None)
        if opts is not None:
            arr = (C.c_double * len(opts))(*opts)
            check(state.h, state.L.vfvm_amg_set_options(state.h, arr, len(opts)))
        state.linear_cache = key
        fresh = True
    else:
        fresh = False
    if isinstance(m, DeviceDirectLike):
        return fresh, m.abstol, m.reltol, m.maxiters
    return fresh, control.abstol_linear, control.reltol_linear, control.maxiters_linear


def _solve_linear(state: SystemState, hist: NewtonSolverHistory, control: SolverControl, reuse_precs: bool):
    """`_solve_linear!` src/vfvm_linsolve.jl:6-61: A * update = residual on the device"""
    fresh, abstol, reltol, maxit = _linear_setup(state, control)
    reuse = bool(reuse_precs and not fresh)
    # nlu counts the preconditioner set-ups that really happen: the device keeps only an ILU factorisation across solves, everything else
    # (Jacobi, block-Jacobi, the AMG numeric phase) is rebuilt from the current Jacobian whatever `reuse` says (csrc/linsolve.cu)
    keeps = getattr(control.method_linear, "precon", None) in (_lib.PRECON_ILU0, _lib.PRECON_ILU0_MC)
    if not (reuse and keeps):
        hist.nlu += 1
    iters, resn = C.c_int(0), C.c_double(0.0)
    rc = state.L.vfvm_linsolve(state.h, abstol, reltol, maxit, 1 if reuse else 0, C.byref(iters), C.byref(resn))
    t = state.timings()
    hist.tlinsolve_setup += t[_lib.TIME_LINSOLVE_SETUP] * 1e-3
    hist.tlinsolve_solve += t[_lib.TIME_LINSOLVE_SOLVE] * 1e-3
    hist.nlin = iters.value
    if rc == _lib.ERR_LINSOLVE:
        raise LinearSolverError(state.L.vfvm_last_error(state.h).decode())
    check(state.h, rc)
    hist.linres = resn.value
    m = control.method_linear
    if m is None or isinstance(m, DeviceDirectLike):
        # the stand-in for a direct solve must not hand Newton an unconverged update silently
        conv, bnorm = C.c_int(0), C.c_double(0.0)
        check(state.h, state.L.vfvm_linsolve_status(state.h, C.byref(conv), C.byref(bnorm)))
        if not conv.value:
            raise LinearSolverError(f"direct-like default solver (BiCGStab + block-Jacobi) stopped at maxiters={maxit} with |r| = {resn.value:.3e}, |b| = {bnorm.value:.3e}; "
                                    "choose method_linear=KrylovJL_BICGSTAB/CG(precs=AMGPreconBuilder()) for this problem")
    return iters.value, resn.value


def solve_step(state: SystemState, oldsol_on_device: bool, control: SolverControl, time: float, tstep: float, embedparam: float, istep_factorize: int):
    """`solve_step!` (src/vfvm_solver.jl:13-222).  Precondition: VEC_OLDSOL holds the old time step / initial value.
    On return VEC_SOLUTION holds the new solution (device resident); returns the NewtonSolverHistory."""
    L, h = state.L, state.h
    sysm = state.system
    hist = NewtonSolverHistory()
    t0 = _time.perf_counter()
    state.sync()
    check(h, L.vfvm_copy_vector(h, _lib.VEC_SOLUTION, _lib.VEC_OLDSOL))  # solution .= oldsol  (:28)
    check(h, L.vfvm_init_dirichlet(h, time, embedparam))  # _initialize! (:31)
    oldnorm, converged, damp = 1.0, False, 1.0
    ninf, n1 = C.c_double(), C.c_double()
    rnorm = 0.0
    if not sysm.is_linear:
        damp = control.damp_initial
        check(h, L.vfvm_vector_norms(h, _lib.VEC_SOLUTION, C.byref(ninf), C.byref(n1)))
        rnorm = n1.value  # control.rnorm(solution) = ||.||_1 (:55)
    nround, tolx, niter = 0, 0.0, 1
    while niter <= control.maxiters:
        rc = L.vfvm_assemble(h, time, tstep, embedparam)  # eval_and_assemble (:69)
        hist.tasm += state.timings()[_lib.TIME_ASSEMBLE] * 1e-3
        if rc == _lib.ERR_NAN:
            raise AssemblyError("trying to assemble NaN")  # src/vfvm_assembly.jl:10-12 -> AssemblyError (:87-96)
        check(h, rc)
        reuse_precs = (not control.factorize_every_newtonstep and niter > 1) or (istep_factorize % control.factorize_every_timestep != 0)  # :99
        norm = None
        if not control.updatecontrol:
            check(h, L.vfvm_vector_norms(h, _lib.VEC_RESIDUAL, C.byref(ninf), C.byref(n1)))
            norm = ninf.value
        tl0 = _time.perf_counter()
        _solve_linear(state, hist, control, reuse_precs)  # :105
        hist.tlinsolve += _time.perf_counter() - tl0
        check(h, L.vfvm_newton_update(h, damp, C.byref(ninf), C.byref(n1)))  # :116 + norms (:126,:132)
        if sysm.is_linear:
            converged = True
            break
        damp = min(damp * control.damp_growth, 1.0)
        if control.updatecontrol:
            norm = ninf.value
        if tolx == 0.0:
            tolx = norm * control.reltol
        dnorm = 1.0
        rnorm_new = n1.value
        if rnorm > 1.0e-50:
            dnorm = abs((rnorm - rnorm_new) / rnorm)
        nround = nround + 1 if dnorm < control.tol_round else 0
        if control.log:
            hist.l1normdiff.append(dnorm)
            hist.updatenorm.append(norm)
        if control.verbose is True or (isinstance(control.verbose, str) and "n" in control.verbose):
            print(f"  [n]ewton: {niter:3d}({hist.nlin:3d}) {norm:.3e} {'' if niter == 1 else f'{norm / oldnorm:.3e}'} {dnorm:.3e} {nround:2d}")
        if niter > 1 and norm / oldnorm > 1.0 / control.tol_mono:
            converged = False
            break
        if norm < control.abstol or norm < tolx:
            converged = True
            break
        oldnorm, rnorm = norm, rnorm_new
        if nround > control.max_round:
            converged = True
            break
        niter += 1
    if not converged:
        raise ConvergenceError()
    hist.time = _time.perf_counter() - t0
    state.history = hist
    return hist


def _as_inival(system: System, inival):
    """dense (n, N) initial value for the device twin; a SparseSolutionArray is scattered (undefined dofs are zero on the device)"""
    if np.isscalar(inival):
        u = unknowns(system, inival)
        return u.dense() if hasattr(u, "dense") else u
    if hasattr(inival, "dense"):
        return inival.dense()
    a = np.asfortranarray(inival, dtype=np.float64)
    if a.shape != (system.num_species, system.grid.num_nodes):
        raise ValueError(f"wrong shape of inival: {a.shape}")
    return a


def solve_transient(state: SystemState, inival, lambdas, control: SolverControl, transient=True, time=0.0):
    """`solve_transient!` (src/vfvm_solver.jl:267-537): implicit Euler / embedding with step size control; the state
    vectors stay on the device, one solution download per stored step."""
    L, h = state.L, state.h
    if transient:
        dl, dl_min, dl_max, dl_grow, dl_decrease = control.Δt, control.Δt_min, control.Δt_max, control.Δt_grow, control.Δt_decrease
    else:
        dl, dl_min, dl_max, dl_grow, dl_decrease = control.Δp, control.Δp_min, control.Δp_max, control.Δp_grow, control.Δp_decrease
    du_opt, du_max_factor = control.Δu_opt, control.Δu_max_factor
    state.set_vector(_lib.VEC_OLDSOL, inival)
    istep_factorize = 0
    if transient:
        tsol = TransientSolution(lambdas[0], np.array(inival, order="F", copy=True))
    else:
        if control.pre is not _no_pre:
            control.pre(np.array(inival, order="F", copy=True), float(lambdas[0]))
        hist = solve_step(state, True, control, time, math.inf, float(lambdas[0]), istep_factorize)
        sol = state.get_vector(_lib.VEC_SOLUTION)
        control.post(sol, inival, lambdas[0], 0)
        check(h, L.vfvm_copy_vector(h, _lib.VEC_OLDSOL, _lib.VEC_SOLUTION))
        tsol = TransientSolution(lambdas[0], sol)
        tsol.history.append(hist)
    istep, solved = 0, False
    dnorm = C.c_double()
    lam0 = float(lambdas[0])
    for i in range(len(lambdas) - 1):
        dl = max(dl, dl_min)
        lam, lend = float(lambdas[i]), float(lambdas[i + 1])
        while lam < lend:
            solved, lam0, du = False, lam, 0.0
            while not solved:
                solved, forced, errored = True, False, False
                try:
                    lam = lam0 + dl
                    if control.pre is not _no_pre:  # the callbacks see the arrays the reference hands them (src/vfvm_solver.jl:380, :478, :519);
                        control.pre(state.get_vector(_lib.VEC_OLDSOL), lam)  # they are downloaded only when a callback is set
                    if transient:
                        hist = solve_step(state, True, control, lam, dl, 0.0, istep)
                    else:
                        hist = solve_step(state, True, control, time, math.inf, lam, istep)
                except (ConvergenceError, AssemblyError, LinearSolverError) as err:
                    if not control.handle_exceptions:
                        raise RuntimeError(f"Solver problem at {lam:.5g}, step {dl:.5g}: {err!r}") from err
                    solved, errored = False, True
                if solved:
                    check(h, L.vfvm_vector_diffnorm(h, _lib.VEC_SOLUTION, _lib.VEC_OLDSOL, C.byref(dnorm)))  # control.delta (:407)
                    du = dnorm.value
                    if du > du_max_factor * du_opt:
                        solved = False
                    istep_factorize += 1
                if not solved:
                    if math.isclose(dl, dl_min, rel_tol=1.4901161193847656e-8):
                        if not (control.force_first_step and istep == 0):
                            msg = f"Δ_min={dl_min:.5g} reached while Δu/Δu_opt={du / du_opt:.5g}"
                            if control.handle_exceptions:
                                print("warning:", msg)
                                break
                            raise RuntimeError(msg)
                        elif not errored:
                            forced, solved = True, True
                        else:
                            if control.handle_exceptions:
                                break
                            raise RuntimeError("Convergence problem in first timestep")
                    else:
                        dl = max(dl_min, dl * dl_decrease)
                        istep_factorize = 0
            if solved:
                istep += 1
                want_post = control.post is not _no_post
                sol = state.get_vector(_lib.VEC_SOLUTION) if (control.store_all or want_post) else None
                if control.log:
                    tsol.history.append(hist)
                if control.store_all:
                    tsol.append(lam, sol)
                if want_post:
                    control.post(sol, state.get_vector(_lib.VEC_OLDSOL), lam, dl)
                check(h, L.vfvm_copy_vector(h, _lib.VEC_OLDSOL, _lib.VEC_SOLUTION))  # oldsolution .= solution (:480)
                steps_to_go = math.ceil((lend - lam) / dl)
                lpredict = lend - lam
                if 0 < steps_to_go < control.num_final_steps:
                    lpredict = (lend - lam) / steps_to_go
                if math.isclose(dl_max, dl_min, rel_tol=1.4901161193847656e-8):
                    lpredict = dl_max
                if lam < lend:
                    dl = min(dl_max, dl * dl_grow, dl * du_opt / (du + 1.0e-14), lpredict, lend - lam)
                    if abs(lam + dl - lend) <= max(1.0e-15, 1.0e-15 * max(abs(lam + dl), abs(lend))):
                        dl = lend - lam
            else:
                break
        last = state.get_vector(_lib.VEC_SOLUTION) if (not control.store_all or control.sample is not _no_sample) else None
        if not control.store_all:
            tsol.append(lam0, last)
        if control.sample is not _no_sample:
            control.sample(last, lam0)
        if not solved:
            break
    return tsol


def solve_state(state: SystemState, inival=0, control: SolverControl | None = None, time=0.0, tstep=math.inf, times=None, embed=None, **kwargs):
    """`solve!(state; ...)` src/vfvm_solver.jl:551-617"""
    control = dataclasses.replace(control) if control is not None else SolverControl()
    for k, v in kwargs.items():  # any SolverControl field may be given as keyword (:570-576)
        if k in ("data", "params"):
            continue
        if not hasattr(control, k):
            raise TypeError(f"unknown keyword {k!r}")
        setattr(control, k, v)
    inival = _as_inival(state.system, inival)
    if times is not None:
        return solve_transient(state, inival, list(times), control, transient=True, time=times[0])
    if embed is not None:
        return solve_transient(state, inival, list(embed), control, transient=False, time=time)
    state.set_vector(_lib.VEC_OLDSOL, inival)
    hist = solve_step(state, True, control, time, tstep, 0.0, 0)
    sol = state.get_vector(_lib.VEC_SOLUTION)
    if state.system.unknown_storage == "sparse":  # the reference hands back the storage type of the system (src/vfvm_solver.jl:607-616)
        from .sparsesolution import SparseSolutionArray

        return SparseSolutionArray.from_dense(state.system.node_dof(), sol, history=hist)
    return sol


def solve(system: System, state: SystemState | None = None, **kwargs):
    """`solve(system; kwargs...)` src/vfvm_solver.jl:665-668.  Pass `state=` to reuse a device twin across calls."""
    own = state is None
    if own:
        state = SystemState(system)
    try:
        return solve_state(state, **kwargs)
    finally:
        if own:
            state.close()


def evaluate_residual_and_jacobian(system: System, u, t=0.0, tstep=math.inf, embed=0.0, state: SystemState | None = None):
    """`evaluate_residual_and_jacobian(sys, u; ...)` src/vfvm_solver.jl:256-260 -> (residual (n,N), scipy CSC matrix)"""
    own = state is None
    if own:
        state = SystemState(system)
    try:
        F = state.eval_res_jac(u, u, time=t, tstep=tstep, embed=embed)
        return F, state.matrix("csc")
    finally:
        if own:
            state.close()
