#!/usr/bin/env python
"""Benchmark of the Newton hot path (BASELINE.json metric: fp64 residual+Jacobian assembly Medges/s; Newton step time at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--nx NX] [--impl reference]

One "step" = one eval_and_assemble-equivalent pass (row kernel + boundary-node kernel) over the whole grid at a generic state resident
in HBM.  Default workload = cfg3 (Example301 physics on the 193^3 tensor grid, 49.9 M edges); its line also carries
  * `newton_step`: assembly + Krylov solve (1e-10) + update, median of 5, with its own roofline,
  * `parity`: the residual and Jacobian rows of six node planes per rank against the CPU oracle, entry by entry, and the solution
    of the Newton step (solved to 1e-13) at 4096 sample nodes against the committed reference (tests/golden/newton_samples_*.json),
  * `north_star`: the same figures for the north_star target itself -- cfg4, the three-species bipolar drift-diffusion system of
    Example161 on the same 193^3 grid (49.9 M edges).
Prints ONE JSON line.  `--impl reference` times the CPU oracle (the C++/OpenMP restatement of the reference's own edge loop; the Julia
reference cannot run in this image) on all host cores at the SAME grid, plus its Newton step with the oracle's CPU Krylov solver.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vfvm_b200 as v  # noqa: E402
from vfvm_b200 import physics as ph  # noqa: E402

DEFAULT_NX = {"cfg1": 578, "cfg2": 2583, "cfg3": 193, "cfg4": 193, "cfg5": 97}
AMG_WDEPTH = {"cfg2": 3, "cfg3": 2}  # levels 1..wdepth are visited twice per visit of their parent; elsewhere the V-cycle is the faster one
W_CYCLE_MIN_NODES_PER_RANK = 2500000


# ------------------------------------------------------------------------------------------------ workloads
def make_system(workload: str, nx: int | None):
    nx = nx or DEFAULT_NX[workload]
    X = np.linspace(0, 1, nx)
    if workload == "cfg1":  # Example201_Laplace2D
        s = v.System(v.simplexgrid(X, X), flux=ph.LinearDiffusion(), is_linear=True)
        v.enable_species(s, 1, [1])
        v.boundary_dirichlet(s, 1, 1, 0.0)
        v.boundary_dirichlet(s, 1, 3, 1.0)
        return s, dict(tstep=math.inf), f"cfg1 Example201_Laplace2D {nx}^2"
    if workload == "cfg2":  # Example207_NonlinearPoisson2D
        s = v.System(v.simplexgrid(X, X), flux=ph.PowerDiffusion(1.0e-2, 2), reaction=ph.PowerReaction(1.0, 2.0), source=ph.GaussSource(1, 20.0, (0.5, 0.5)),
                     storage=ph.LinearStorage(1.0))
        v.enable_species(s, 1, [1])
        v.boundary_dirichlet(s, 1, 2, 0.1)
        v.boundary_dirichlet(s, 1, 4, 0.1)
        return s, dict(tstep=0.01), f"cfg2 Example207_NonlinearPoisson2D {nx}^2"
    if workload == "cfg3":  # Example301_Laplace3D
        s = v.System(v.simplexgrid(X, X, X), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0))
        v.enable_species(s, 1, [1])
        v.boundary_dirichlet(s, 1, 5, 0.0)
        v.boundary_dirichlet(s, 1, 6, 0.0)
        return s, dict(tstep=math.inf), f"cfg3 Example301_Laplace3D {nx}^3"
    if workload == "cfg4":  # Example161 bipolar drift-diffusion on a 3D grid, 3 species, 3 z-slab regions
        g = v.simplexgrid(X, X, X)
        v.cellmask(g, [0, 0, 1 / 3], [1, 1, 2 / 3], 2)
        v.cellmask(g, [0, 0, 2 / 3], [1, 1, 1.0], 3)
        bc = ph.BCondition()
        e = math.sqrt(math.exp(-1.0))
        for sp, val in ((1, 0.0), (2, 0.0), (3, 0.5 + math.asinh(10.0 / (2 * e)))):
            bc.dirichlet(species=sp, region=5, value=val)
        for sp, val in ((1, 0.0), (2, 0.0), (3, 0.5 + math.asinh(-10.0 / (2 * e)))):
            bc.dirichlet(species=sp, region=6, value=val)
        s = v.System(g, flux=ph.BipolarSGFlux(), reaction=ph.BipolarReaction([10.0, 0.0, -10.0]), storage=ph.BipolarStorage(), bcondition=bc, species=[1, 2, 3])
        return s, dict(tstep=1.0e-2), f"cfg4 Example161 bipolar drift-diffusion {nx}^3"
    if workload == "cfg5":  # Example410 scaled to 3D, 10 species, implicit Euler
        s = v.System(v.simplexgrid(X, X, X), flux=ph.LinearDiffusion(1.0), storage=ph.LinearStorage(1.0))
        for i in range(1, 11):
            v.enable_species(s, i, [1])
            v.boundary_dirichlet(s, i, 5, 0)
            v.boundary_dirichlet(s, i, 6, 1)
        return s, dict(tstep=0.1), f"cfg5 Example410_ManySpecies(10) {nx}^3"
    raise SystemExit(f"unknown workload {workload}")


def generic_state(system, seed=20261017):
    """smooth deterministic field without exact zeros"""
    g = system.grid
    U = np.empty((system.num_species, g.num_nodes), order="F")
    for i in range(system.num_species):
        U[i, :] = 0.5 + 0.25 * np.sin((3.0 + i) * g.coord[0] + 0.7 * i) * np.cos(2.0 * g.coord[-1] + 0.3)
    return U


def is_spd(system):
    """symmetric positive definite Jacobians (pure linear diffusion) take CG, everything else BiCGStab"""
    fid = system.physics.flux.id if system.physics.flux is not None else 0
    return fid == ph.FLUX_DIFFUSION and system.physics.reaction is None


def algorithmic_bytes(n, N, E, NB, dim, cF, cD, transient):
    """B_asm of SURVEY.md section 8d with c = stored coupling entries: every input read once, every output written once"""
    b = E * (8 + 8 + 16) + N * (8 + 8 * n + 8 * n) + 8 * (cD * N + cF * 2 * E) + NB * dim * (4 + 8)
    if transient:
        b += 8 * n * N
    return b


def iteration_bytes(n, N, nnz_stored, cF, cD, krylov, amg):
    """B_iter of SURVEY.md section 8d on the stored planes: SpMV = nnz (8 cF + 4) + N (8 cD + 16 n); a Krylov iteration = its SpMVs +
    preconditioner applications + vector streams of 8 n N bytes.  AMG cycle = 2 level-0 SpMVs + 7 streams, x (1 + the coarser levels' share:
    0.1 for a V-cycle at 11x coarsening, 0.2 / 0.22 with levels 1 / 1-2 visited twice).  fp64 throughout: the fp32 copy of the finest off-diagonal
    planes the cycle actually reads makes the real traffic smaller than this figure."""
    spmv = nnz_stored * (8 * cF + 4) + N * (8 * cD + 16 * n)
    stream = 8 * n * N
    pre = float(amg) * (2 * spmv + 7 * stream) if amg else 2 * stream  # amg = 1 + share of the coarser levels (1.1 V-cycle, 1.2 / 1.22 W on one / more levels)
    if krylov == "cg":
        return spmv + pre + 6 * stream
    return 2 * spmv + 2 * pre + 10 * stream


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self):
        self.samples, self.reasons, self.maxmhz = [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", os.environ.get("LOCAL_RANK", "0"), f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split("\n")[0].split(",")]
                self.samples.append(float(parts[0]))
                self.maxmhz = float(parts[1])
                for nm, val in zip(names, parts[2:]):
                    if val.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle = port of the reference)
def host_threads():
    """all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port(workload, nx, steps=3, warmup=1, newton=True, newton_budget_s=150.0):
    """The oracle (C++/OpenMP restatement of the reference's coloured edge loop) at the given grid with all host threads: steady-state
    assembly throughput and -- with the oracle's CPU Krylov solver, the stand-in for KrylovJL_CG / KrylovJL_BICGSTAB with Jacobi / node-
    block preconditioning; a sparse LU (the reference default) is out of reach at these sizes -- the Newton step."""
    from oracle import oracle as O

    system, kw, name = make_system(workload, nx)
    t0 = time.perf_counter()
    o = O.OracleSystem(system)
    setup_s = time.perf_counter() - t0
    U = generic_state(system)
    nth = host_threads()
    o.assemble(U, U, tstep=kw["tstep"], nthreads=nth, want_matrix=False)  # first assembly builds the pattern (allocating)
    for _ in range(max(0, warmup - 1)):
        o.assemble(U, U, tstep=kw["tstep"], nthreads=nth, want_matrix=False)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        o.assemble(U, U, tstep=kw["tstep"], nthreads=nth, want_matrix=False)
        ts.append(time.perf_counter() - t0)
    E = o.num_edges
    res = {"name": name, "edges": int(E), "threads": nth, "ms_per_step": float(np.mean(ts)) * 1e3, "medges_per_s": E / float(np.mean(ts)) / 1e6, "setup_s": setup_s}
    if newton:
        try:
            U0 = o.initialize(U)
            t0 = time.perf_counter()
            F, _ = o.assemble(U0, U, tstep=kw["tstep"], nthreads=nth, want_matrix=False)
            t_asm = time.perf_counter() - t0
            spd = is_spd(system)
            method, precon = ("cg" if spd else "bicgstab"), ("jacobi" if system.num_species == 1 else "blockjacobi")
            # bounded: the iteration cap keeps the arm inside its time budget (calibrated on a ten-iteration run of the same solve on this
            # host); an unconverged solve is reported as such
            _, it0, _, sec0 = o.krylov_solve(F, method, precon, reltol=1e-30, maxiters=10, nthreads=nth)
            per_it = sec0 / max(1, it0)
            cap = int(max(50, min(20000, newton_budget_s / max(per_it, 1e-5))))
            x, it, rel, sec = o.krylov_solve(F, method, precon, reltol=1e-10, maxiters=cap, nthreads=nth)
            res["newton_step"] = {"ms": (t_asm + sec) * 1e3, "assemble_ms": t_asm * 1e3, "linsolve_ms": sec * 1e3, "iters": it, "relres": rel, "converged": bool(rel <= 2e-10),
                                  "krylov": ("CG" if spd else "BiCGStab") + ("+Jacobi" if precon == "jacobi" else "+node-block-Jacobi") + f" (oracle CPU Krylov, {nth} threads)",
                                  "unknowns": int(system.num_species * system.grid.num_nodes), "reltol": 1e-10, "iteration_cap": cap}
        except Exception as exc:  # pragma: no cover
            res["newton_step"] = {"error": repr(exc)}
    return res


def run_reference(args):
    """the reference arm: the CPU port at the same grid as the b200 arm, all host cores; rank 0 only"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    nx = args.nx or DEFAULT_NX[args.workload]
    r = cpu_port(args.workload, nx, steps=max(1, args.steps), warmup=max(1, args.warmup), newton=not args.no_newton)
    val = r["medges_per_s"]
    line = {"impl": "reference", "metric": "fp64 residual+Jacobian assembly throughput", "value": val, "unit": "Medges/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": r["name"], "edges": r["edges"], "sample": "the whole workload (same grid as the b200 arm)",
                       "note": "CPU oracle = C++/OpenMP restatement of the reference's coloured edgewise loop (the Julia reference is not runnable here), all host threads"},
            "cpu_baseline": {"value": val, "unit": "Medges/s", "cores": r["threads"], "kind": "port", "sample": f"{r['name']}, {r['edges']} edges per step (full size)"},
            "e2e": {"value": val, "unit": "Medges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "newton_step": r.get("newton_step"), "setup_s": r["setup_s"]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def _stream_ptr(st):
    p = C.c_void_p()
    st.L.vfvm_stream(st.h, C.byref(p))
    return p.value


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def _traffic(workload, nx, world):
    """DRAM bytes per launch of the row kernel from the committed `ncu --set full` capture of the same workload (profiles/traffic.json);
    only meaningful for the single-GPU launch at the captured size"""
    if world != 1 or nx != DEFAULT_NX.get(workload):
        return None
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
    except Exception:
        return None


def linear_setup(st, system, world, workload=None):
    """Krylov method + AMG options of the Newton step: CG for SPD Jacobians, BiCGStab otherwise, aggregation AMG (csrc/amg.cu);
    VFVM_BENCH_PRECON=jacobi gives the one-level baseline"""
    L, h = st.L, st.h
    spd = is_spd(system)
    krylov = v._lib.KRYLOV_CG if spd else v._lib.KRYLOV_BICGSTAB
    precon = v._lib.PRECON_JACOBI if system.num_species == 1 or spd else v._lib.PRECON_BLOCKJACOBI
    if os.environ.get("VFVM_BENCH_PRECON", "amg") == "amg":
        precon = v._lib.PRECON_AMG
    v._lib.check(h, L.vfvm_linsolve_setup(h, krylov, precon, 0))
    wdepth = 0
    if precon == v._lib.PRECON_AMG:
        # W-cycle on the top levels where it pays (profiles/r2_amg_sweeps.txt); the cycle in use is named in `krylov`.
        # VFVM_BENCH_AMG_OPTS = "omega,alpha,theta,sweeps,coarse_sweeps,wdepth" overrides (empty = keep)
        # The second visits are pure latency (small coarse-level kernels): they pay while an iteration is bandwidth bound, i.e. while a rank holds
        # enough of the finest level -- measured on cfg3: 1 GPU (7.2 M nodes) 78 -> 53 ms, 2 GPUs (3.6 M per rank) 59 -> 49 ms, but 8 GPUs (0.9 M per rank) 45.5 -> 57 ms
        wdepth = AMG_WDEPTH.get(workload, 0) if st.Nown >= W_CYCLE_MIN_NODES_PER_RANK else 0
        vals = [float("nan")] * 5 + [float(wdepth)]
        if os.environ.get("VFVM_BENCH_AMG_OPTS"):
            vals = [float(x) if x.strip() else float("nan") for x in os.environ["VFVM_BENCH_AMG_OPTS"].split(",")]
            if len(vals) > 5 and vals[5] == vals[5]:
                wdepth = int(vals[5])
        v._lib.check(h, L.vfvm_amg_set_options(h, (C.c_double * len(vals))(*vals), len(vals)))
    label = ("CG" if spd else "BiCGStab") + {v._lib.PRECON_JACOBI: "+Jacobi", v._lib.PRECON_BLOCKJACOBI: "+block-Jacobi",
                                           v._lib.PRECON_AMG: "+aggregation-AMG (" + (f"W-cycle on levels 1-{wdepth}" if wdepth else "V-cycle") + ")"}[precon]
    return ("cg" if spd else "bicgstab"), ((1.1, 1.2, 1.22)[min(wdepth, 2)] if precon == v._lib.PRECON_AMG else 0), label


def newton_once(st, U, tstep, reltol, maxiters=5000):
    """one Newton step from the generic state (solve_step!, src/vfvm_solver.jl:13-222, first iteration): SOLUTION = OLDSOL with Dirichlet
    values, assemble, solve, update.  U = None: OLDSOL is already resident (timed repetitions do not re-upload it)."""
    L, h = st.L, st.h
    if U is not None:
        st.set_vector(v._lib.VEC_OLDSOL, U)
    v._lib.check(h, L.vfvm_copy_vector(h, v._lib.VEC_SOLUTION, v._lib.VEC_OLDSOL))
    v._lib.check(h, L.vfvm_init_dirichlet(h, 0.0, 0.0))
    iters, resn = C.c_int(), C.c_double()
    rc = L.vfvm_assemble(h, 0.0, tstep, 0.0)
    assert rc == 0, rc
    rc = L.vfvm_linsolve(h, 0.0, reltol, maxiters, 0, C.byref(iters), C.byref(resn))
    ninf, n1 = C.c_double(), C.c_double()
    v._lib.check(h, L.vfvm_newton_update(h, 1.0, C.byref(ninf), C.byref(n1)))
    return rc, iters.value, resn.value, ninf.value


def newton_solution(st, system, U, tstep, reltol=1.0e-13, workload=None):
    """solution after one Newton step solved to `reltol` (tools/make_newton_golden.py --source device)"""
    _, _, label = linear_setup(st, system, 1, workload)
    rc, it, resn, ninf = newton_once(st, U, tstep, reltol)
    return st.get_vector(v._lib.VEC_SOLUTION), f"{label}, reltol {reltol:g}, {it} iterations, |r| = {resn:.3e}, rc {rc}"


def run_workload(args, workload, nx, ctx, steps, warmup, newton_reps):
    """all device figures of one workload: assembly throughput, roofline of the row kernel, e2e, Newton step, parity"""
    import torch

    rank, world, local, dist = ctx["rank"], ctx["world"], ctx["local"], ctx["dist"]
    system, kw, name = make_system(workload, nx)
    nx = nx or DEFAULT_NX[workload]
    tstep = kw["tstep"]
    n = system.num_species
    t_setup0 = time.perf_counter()
    if world > 1:
        from vfvm_b200 import partition as part_mod

        st, pinfo = part_mod.partitioned_state(system, rank, world, local)
    else:
        st, pinfo = v.SystemState(system, device=local), None
    setup_s = time.perf_counter() - t_setup0
    Uglob = generic_state(system)
    U = Uglob if pinfo is None else np.asfortranarray(Uglob[:, pinfo.local_nodes])
    st.set_vector(v._lib.VEC_SOLUTION, U)
    st.set_vector(v._lib.VEC_OLDSOL, U)
    my_edges = st.block_counts()[0] / 2.0  # every edge appears in exactly two owned rows over all ranks
    stream = torch.cuda.ExternalStream(_stream_ptr(st), device=torch.device("cuda", local))
    L, h = st.L, st.h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def one_step():  # enqueue one eval_and_assemble pass on the handle's stream (status collected by vfvm_sync after the loop)
        rc = L.vfvm_assemble_async(h, 0.0, tstep, 0.0)
        assert rc == 0, rc

    for _ in range(warmup):
        one_step()
    assert L.vfvm_sync(h) == 0
    ktimes = []
    for _ in range(min(10, steps)):  # kernel-only time of the row kernel (CUDA events around it inside the library)
        assert st.assemble(time=0.0, tstep=tstep, embed=0.0) == 0
        ktimes.append(st.timings()[v._lib.TIME_EDGE_KERNEL])
    clocks = ClockSampler()
    if rank == 0 and not args.no_clocks:
        clocks.start()
    launches0 = st.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        one_step()
    e1.record(stream)
    assert L.vfvm_sync(h) == 0  # also reports a NaN raised by any of the K passes
    barrier()
    ms_total = allmax(e0.elapsed_time(e1))
    launches = st.launch_count() - launches0
    clk = clocks.stop() if (rank == 0 and not args.no_clocks) else None
    E_total = allsum(my_edges)
    ms_step = ms_total / steps
    value = E_total / (ms_step * 1e-3) / 1e6

    # ---- parity of the assembled state (rank-local probe planes against the CPU oracle, entry by entry)
    parity = {}
    if not args.no_parity:
        from parity_probe import probe_rows

        pr = probe_rows(system, st, pinfo, Uglob, Uglob, tstep=tstep)
        if world > 1:
            allp = [None] * world
            dist.all_gather_object(allp, pr)
        else:
            allp = [pr]
        parity["assembly"] = {"ok": all(p["ok"] for p in allp), "pattern_equal": all(p["pattern_equal"] for p in allp), "rows_checked": sum(p["rows"] for p in allp),
                              "entries_checked": sum(p["entries"] for p in allp), "max_rel_err_entry": max(p["max_rel_err_entry"] for p in allp),
                              "max_err_over_bound": max(p["max_err_over_bound"] for p in allp), "residual_max_err_over_bound": max(p["residual_max_err_over_bound"] for p in allp),
                              "bound": "1e-12 |a| + 8 eps sum|row terms| (tests/test_gpu_parity.py); first, middle and last two node planes of every rank's rows vs the CPU oracle"}

    # ---- roofline of the dominant kernel (row assembly), measured live with CUDA events on the handle's stream
    g = system.grid
    cF, cD = st.matrix_plane_counts()
    bytes_alg = algorithmic_bytes(n, st.Nown, my_edges, g.num_bfaces if pinfo is None else pinfo.num_bfaces, g.dim, cF, cD, math.isfinite(tstep))
    peak, peak_src = _peaks()
    kms = float(np.mean(ktimes))
    achieved = bytes_alg / (kms * 1e-3) / 1e9
    traffic = _traffic(workload, nx, world)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture of this kernel" if traffic else None,
                "peak_source": peak_src, "kernel": "k_node_transform + k_assemble_rows_bipolar" if workload == "cfg4" else "k_assemble_rows", "kernel_ms": kms,
                "algorithmic_bytes": bytes_alg, "per_rank": world > 1}

    # ---- e2e: the same pass through the C ABI with HOST buffers: pinned U -> device, assemble, residual -> pinned host
    nd = n * st.N
    hU = torch.empty(nd, dtype=torch.float64).pin_memory()
    hF = torch.empty(nd, dtype=torch.float64).pin_memory()
    hU.numpy()[:] = U.ravel(order="F")

    def e2e_step():
        rc = L.vfvm_eval_res_jac(h, hU.data_ptr(), None, hF.data_ptr(), v._lib.HOST, 0.0, tstep, 0.0)
        assert rc == 0, rc

    for _ in range(3):
        e2e_step()
    barrier()
    k2 = max(3, steps // 2)
    t0 = time.perf_counter()
    for _ in range(k2):
        e2e_step()
    barrier()
    e2e_ms = allmax((time.perf_counter() - t0) / k2 * 1e3)
    e2e = {"value": E_total / (e2e_ms * 1e-3) / 1e6, "unit": "Medges/s", "h2d_bytes_per_step": 8 * nd, "d2h_bytes_per_step": 8 * nd, "ms_per_step": e2e_ms,
           "call": "vfvm_eval_res_jac(host U -> host F), Jacobian stays in HBM for the linear solve"}

    # ---- Newton step (assembly + Krylov solve to 1e-10 + update): median of `newton_reps`, CUDA events on the handle's stream, max over ranks
    newton = None
    if not args.no_newton:
        krylov, amg, label = linear_setup(st, system, world, workload)
        newton_once(st, U, tstep, 1.0e-10, maxiters=3)  # warm-up: work vectors, hierarchy, NCCL channels
        reps = []
        for _ in range(newton_reps):
            barrier()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = st.launch_count()
            n0.record(stream)
            rc, it, resn, ninf = newton_once(st, None, tstep, 1.0e-10)
            n1.record(stream)
            torch.cuda.synchronize()
            t = st.timings()
            reps.append((allmax(n0.elapsed_time(n1)), it, resn, ninf, rc, float(t[0]), float(t[1]), float(t[2]), st.launch_count() - l0))
        reps.sort(key=lambda r: r[0])
        med = reps[len(reps) // 2]
        nnz_stored = st.block_counts()[1]
        b_iter = iteration_bytes(n, st.Nown, nnz_stored, cF, cD, krylov, amg)
        b_newton = bytes_alg + med[1] * b_iter + 3 * 8 * n * st.Nown
        ach = b_newton / (med[0] * 1e-3) / 1e9
        newton = {"ms": med[0], "ms_all": [r[0] for r in reps], "assemble_ms": med[5], "linsolve_setup_ms": med[6], "linsolve_solve_ms": med[7], "krylov": label, "reltol": 1e-10,
                  "iters": med[1], "resnorm": med[2], "update_norm_inf": med[3], "rc": med[4], "gpu_launches": int(med[8]),
                  "ms_per_iteration": med[7] / max(1, med[1]),
                  "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes": b_newton, "per_rank": world > 1,
                               "formula": "B_asm + iters x B_iter + 3 vector streams (SURVEY 8d; B_iter: bench.py iteration_bytes)"}}
        if not args.no_parity:
            # the same step solved to 1e-13: its solution at the golden sample nodes against the committed reference
            rc, it, resn, ninf = newton_once(st, None, tstep, 1.0e-13)
            parity["newton"] = _newton_parity(st, system, pinfo, workload, nx, ctx, it, resn)

    out = {"name": name, "value": value, "ms_per_step": ms_step, "E_total": E_total, "setup_s": setup_s, "roofline": roofline, "e2e": e2e, "newton_step": newton,
           "parity": parity or None, "gpu_launches": int(launches), "clocks": clk, "device_bytes": st.device_bytes(), "species": n, "nodes": int(g.num_nodes),
           "transport": ("peer mailboxes (CUDA IPC over NVLink)" if getattr(st, "peer", False) else "NCCL") if world > 1 else None}
    st.close()
    return out


def _newton_parity(st, system, pinfo, workload, nx, ctx, iters, resn):
    path = os.path.join(ROOT, "tests", "golden", f"newton_samples_{workload}_{nx}.json")
    if not os.path.exists(path):
        return {"golden": None, "note": f"no committed reference for {workload} at nx={nx} (tools/make_newton_golden.py)"}
    gold = json.load(open(path))
    nodes = np.asarray(gold["nodes"], dtype=np.int64)
    ref = np.asarray(gold["solution"], dtype=np.float64)
    sol = st.get_vector(v._lib.VEC_SOLUTION)
    if pinfo is None:
        diff = float(np.max(np.abs(sol[:, nodes] - ref)))
        cnt = int(nodes.size)
    else:
        g2l = np.full(system.grid.num_nodes, -1, dtype=np.int64)
        g2l[pinfo.owned_global] = np.arange(pinfo.n_owned)
        mine = g2l[nodes] >= 0
        diff = float(np.max(np.abs(sol[:, g2l[nodes[mine]]] - ref[:, mine]))) if mine.any() else 0.0
        cnt = int(mine.sum())
    if ctx["world"] > 1:
        alld = [None] * ctx["world"]
        ctx["dist"].all_gather_object(alld, (diff, cnt))
        diff, cnt = max(d for d, _ in alld), sum(c for _, c in alld)
    return {"max_abs_diff": diff, "samples": cnt, "tolerance": 1e-10, "ok": bool(diff <= 1e-10 and cnt == nodes.size), "solve": f"reltol 1e-13, {iters} iterations, |r| = {resn:.2e}",
            "golden": os.path.relpath(path, ROOT), "golden_source": gold.get("source"), "golden_how": gold.get("how")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-north-star", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    default_run = args.workload is None
    args.workload = args.workload or "cfg3"
    if args.impl == "reference":
        return run_reference(args)

    # the JSON line is the only thing on stdout: native libraries write their banners to fd 1 (NCCL prints its version there at
    # NCCL_DEBUG=VERSION/WARN), so fd 1 is pointed at stderr for the run and the line goes to a duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = {"rank": rank, "world": world, "local": local, "dist": dist}
    r = run_workload(args, args.workload, args.nx, ctx, args.steps, args.warmup, newton_reps=5)
    ns = None
    if default_run and not args.no_north_star:
        # the north_star target itself: a Newton step of the three-species bipolar drift-diffusion system on the 49.9 M-edge grid
        q = run_workload(args, "cfg4", None, ctx, max(5, args.steps // 2), args.warmup, newton_reps=3)
        ns = {"workload": q["name"], "edges": int(q["E_total"]), "species": q["species"], "value": q["value"], "unit": "Medges/s", "ms_per_step": q["ms_per_step"],
              "roofline": q["roofline"], "e2e": q["e2e"], "newton_step": q["newton_step"], "parity": q["parity"], "gpu_launches": q["gpu_launches"], "device_bytes": q["device_bytes"],
              "setup_s": q["setup_s"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        c = cpu_port(args.workload, args.nx or DEFAULT_NX[args.workload], steps=3, warmup=1, newton=not args.no_newton, newton_budget_s=60.0)
        cpu = {"value": c["medges_per_s"], "unit": "Medges/s", "cores": c["threads"], "kind": "port", "sample": f"{c['name']}, {c['edges']} edges per step (the whole workload), mean of 3",
               "newton_step": c.get("newton_step")}

    if rank == 0:
        line = {"metric": "fp64 residual+Jacobian assembly throughput", "value": r["value"], "unit": "Medges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": r["name"], "species": r["species"], "nodes": r["nodes"], "edges": int(r["E_total"]),
                           "l2": "inputs larger than L2 (tile stream >> 126 MB)" if r["E_total"] * 24 / world > 2.0e8 else "inputs fit L2",
                           "parallelism": f"node-owner partitions x{world}" if world > 1 else "single GPU", "transport": r["transport"], "setup_s": r["setup_s"]},
                "roofline": r["roofline"], "cpu_baseline": cpu, "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "newton_step": r["newton_step"],
                "parity": r["parity"], "north_star": ns, "device_bytes": r["device_bytes"]}
        if cpu and cpu.get("newton_step") and r["newton_step"] and "ms" in cpu["newton_step"]:
            line["newton_vs_cpu_port"] = {"ratio": cpu["newton_step"]["ms"] / r["newton_step"]["ms"], "same_config": True,
                                          "note": "CPU port Newton step (oracle assembly + oracle Krylov, all host threads) / device Newton step, same grid, same tolerance"
                                                  + ("" if cpu["newton_step"].get("converged") else "; the CPU solve hit its iteration cap before 1e-10 (lower bound of its time)")}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
