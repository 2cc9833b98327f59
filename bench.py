#!/usr/bin/env python
"""Benchmark of the Newton hot path (BASELINE.json metric: fp64 residual+Jacobian assembly Medges/s; Newton step time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--nx NX] [--impl reference]

One "step" = one eval_and_assemble-equivalent pass (row-tile kernel + boundary-node kernel) over the whole grid at a
generic state resident in HBM.  Default workload = cfg3 (Example301 physics on the 193^3 tensor grid, 49.9 M edges).
Prints ONE JSON line (see the contract in the task description).  `--impl reference` times the CPU oracle (the
restatement of the reference's own edge loop; the Julia reference cannot run in this image) on the host cores.
"""
from __future__ import annotations

import argparse
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

import vfvm_b200 as v  # noqa: E402
from vfvm_b200 import physics as ph  # noqa: E402


# ------------------------------------------------------------------------------------------------ workloads
def make_system(workload: str, nx: int | None):
    if workload == "cfg1":  # Example201_Laplace2D
        nx = nx or 578
        X = np.linspace(0, 1, nx)
        s = v.System(v.simplexgrid(X, X), flux=ph.LinearDiffusion(), is_linear=True)
        v.enable_species(s, 1, [1])
        v.boundary_dirichlet(s, 1, 1, 0.0)
        v.boundary_dirichlet(s, 1, 3, 1.0)
        return s, dict(tstep=math.inf), f"cfg1 Example201_Laplace2D {nx}^2"
    if workload == "cfg2":  # Example207_NonlinearPoisson2D
        nx = nx or 2583
        X = np.linspace(0, 1, nx)
        s = v.System(v.simplexgrid(X, X), flux=ph.PowerDiffusion(1.0e-2, 2), reaction=ph.PowerReaction(1.0, 2.0), source=ph.GaussSource(1, 20.0, (0.5, 0.5)),
                     storage=ph.LinearStorage(1.0))
        v.enable_species(s, 1, [1])
        v.boundary_dirichlet(s, 1, 2, 0.1)
        v.boundary_dirichlet(s, 1, 4, 0.1)
        return s, dict(tstep=0.01), f"cfg2 Example207_NonlinearPoisson2D {nx}^2"
    if workload == "cfg3":  # Example301_Laplace3D
        nx = nx or 193
        X = np.linspace(0, 1, nx)
        s = v.System(v.simplexgrid(X, X, X), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0))
        v.enable_species(s, 1, [1])
        v.boundary_dirichlet(s, 1, 5, 0.0)
        v.boundary_dirichlet(s, 1, 6, 0.0)
        return s, dict(tstep=math.inf), f"cfg3 Example301_Laplace3D {nx}^3"
    if workload == "cfg4":  # Example161 bipolar drift-diffusion on a 3D grid, 3 species, 3 z-slab regions
        nx = nx or 129
        X = np.linspace(0, 1, nx)
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
        nx = nx or 97
        X = np.linspace(0, 1, nx)
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


def algorithmic_bytes(n, N, E, NB, dim, cF, cD, transient):
    """B_asm of SURVEY.md section 8d with c = stored coupling entries: every input read once, every output written once"""
    b = E * (8 + 8 + 16) + N * (8 + 8 * n + 8 * n) + 8 * (cD * N + cF * 2 * E) + NB * dim * (4 + 8)
    if transient:
        b += 8 * n * N
    return b


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


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_time(workload, nx_sample, repeats=3):
    """the oracle (C++/OpenMP restatement of the reference's coloured edge loop) on a bounded sample of the workload"""
    from oracle import oracle as O

    system, kw, name = make_system(workload, nx_sample)
    o = O.OracleSystem(system)
    U = generic_state(system)
    nthreads = O.lib().vo_max_threads()
    o.assemble(U, U, tstep=kw["tstep"], nthreads=nthreads, want_matrix=False)  # first assembly builds the pattern (allocating)
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.assemble(U, U, tstep=kw["tstep"], nthreads=nthreads, want_matrix=False)
        ts.append(time.perf_counter() - t0)
    E = o.num_edges
    # Newton step on the same sample: oracle assembly + SciPy Krylov with Jacobi scaling to the tolerance the GPU arm uses
    newton = None
    try:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        U0 = o.initialize(U)
        t0 = time.perf_counter()
        F, A = o.assemble(U0, U0, tstep=kw["tstep"], nthreads=nthreads)
        t_asm = time.perf_counter() - t0
        A = A.tocsr()
        d = A.diagonal()
        M = spla.LinearOperator(A.shape, matvec=lambda x: x / d)
        its = [0]
        t0 = time.perf_counter()
        spd = system.physics.flux is not None and system.physics.flux.id == ph.FLUX_DIFFUSION and system.physics.reaction is None
        solver = spla.cg if spd else spla.bicgstab
        x, info = solver(A, F.ravel(order="F"), rtol=1e-10, atol=0.0, maxiter=5000, M=M, callback=lambda xk: its.__setitem__(0, its[0] + 1))
        t_sol = time.perf_counter() - t0
        newton = {"assemble_ms": t_asm * 1e3, "linsolve_ms": t_sol * 1e3, "iters": its[0], "krylov": ("CG" if spd else "BiCGStab") + "+Jacobi (SciPy, 1 thread)", "unknowns": int(A.shape[0])}
    except Exception as exc:  # pragma: no cover
        newton = {"error": repr(exc)}
    return E / min(ts) / 1e6, nthreads, f"{name}, {E} edges, best of {repeats} steady-state assemblies", E, newton


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_nx = {"cfg1": 578, "cfg2": 900, "cfg3": 65, "cfg4": 33, "cfg5": 33}[args.workload]
    from oracle import oracle as O

    system, kw, name = make_system(args.workload, sample_nx)
    full_nx = {"cfg1": 578, "cfg2": 2583, "cfg3": 193, "cfg4": 129, "cfg5": 97}[args.workload]
    full_name = name.replace(f" {sample_nx}^", f" {full_nx}^")  # the b200 arm's workload; this arm times a bounded sample of it
    o = O.OracleSystem(system)
    U = generic_state(system)
    nthreads = O.lib().vo_max_threads()
    for _ in range(max(1, args.warmup)):
        o.assemble(U, U, tstep=kw["tstep"], nthreads=nthreads, want_matrix=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.assemble(U, U, tstep=kw["tstep"], nthreads=nthreads, want_matrix=False)
    dt = (time.perf_counter() - t0) / args.steps
    val = o.num_edges / dt / 1e6
    line = {"impl": "reference", "metric": "fp64 residual+Jacobian assembly throughput", "value": val, "unit": "Medges/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": full_name, "sample": name,
                       "note": "CPU oracle = C++/OpenMP restatement of the reference's coloured edgewise loop (Julia reference not runnable here), timed on a bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": "Medges/s", "cores": nthreads, "kind": "port", "sample": f"{name}, {o.num_edges} edges per step"},
            "e2e": {"value": val, "unit": "Medges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        from vfvm_b200 import partition as part_mod
    system, kw, name = make_system(args.workload, args.nx)
    tstep = kw["tstep"]
    n = system.num_species
    t_setup0 = time.perf_counter()
    if world > 1:
        st, pinfo = part_mod.partitioned_state(system, rank, world, local)
    else:
        st, pinfo = v.SystemState(system, device=local), None
    setup_s = time.perf_counter() - t_setup0
    Uglob = generic_state(system)
    U = Uglob if pinfo is None else np.asfortranarray(Uglob[:, pinfo.local_nodes])
    st.set_vector(v._lib.VEC_SOLUTION, U)
    st.set_vector(v._lib.VEC_OLDSOL, U)
    E_total = None
    my_edges = st.block_counts()[0] / 2.0  # every edge appears in exactly two owned rows over all ranks
    stream = torch.cuda.ExternalStream(_stream_ptr(st), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L, h = st.L, st.h

    def one_step():  # enqueue one eval_and_assemble pass on the handle's stream (status collected by vfvm_sync after the loop)
        rc = L.vfvm_assemble_async(h, 0.0, tstep, 0.0)
        assert rc == 0, rc

    for _ in range(args.warmup):
        one_step()
    assert L.vfvm_sync(h) == 0
    ktimes = []
    for _ in range(min(10, args.steps)):  # kernel-only time of the row kernel (CUDA events around it inside the library)
        assert st.assemble(time=0.0, tstep=tstep, embed=0.0) == 0
        ktimes.append(st.timings()[v._lib.TIME_EDGE_KERNEL])
    clocks = ClockSampler()
    if rank == 0 and not args.no_clocks:
        clocks.start()
    launches0 = st.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    e1.record(stream)
    assert L.vfvm_sync(h) == 0  # also reports a NaN raised by any of the K passes
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = st.launch_count() - launches0
    clk = clocks.stop() if (rank == 0 and not args.no_clocks) else None
    tt = torch.tensor([ms_total, float(my_edges)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, E_total = float(tmax[0]), float(tsum[1])
    else:
        E_total = float(my_edges)
    ms_step = ms_total / args.steps
    value = E_total / (ms_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (row-tile assembly), measured live with CUDA events on the handle's stream
    g = system.grid
    Nloc = st.Nown
    Eloc = my_edges
    cF, cD = _planes(st)
    bytes_alg = algorithmic_bytes(n, Nloc, Eloc, g.num_bfaces if pinfo is None else pinfo.num_bfaces, g.dim, cF, cD, math.isfinite(tstep))
    peaks = _peaks()
    kms = float(np.mean(ktimes))
    achieved = bytes_alg / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks[0], "unit": "GB/s", "frac": achieved / peaks[0], "traffic": _traffic(args.workload),
                "peak_source": peaks[1], "kernel": "k_node_transform + k_assemble_rows_bipolar" if args.workload == "cfg4" else "k_assemble_rows", "kernel_ms": kms, "algorithmic_bytes": bytes_alg}

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
    k2 = max(3, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(k2):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) / k2 * 1e3
    tt = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = E_total / (float(tt[0]) * 1e-3) / 1e6
    e2e = {"value": e2e_val, "unit": "Medges/s", "h2d_bytes_per_step": 8 * nd, "d2h_bytes_per_step": 8 * nd, "ms_per_step": float(tt[0]),
           "call": "vfvm_eval_res_jac(host U -> host F), Jacobian stays in HBM for the linear solve"}

    # ---- Newton step (assembly + Krylov solve + update), reported beside the headline
    newton = None
    if not args.no_newton:
        newton = _newton_step(st, system, U, tstep, world)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_nx = {"cfg1": 578, "cfg2": 900, "cfg3": 65, "cfg4": 33, "cfg5": 33}[args.workload]
        val, cores, sample, _, cpu_newton = cpu_port_time(args.workload, sample_nx)
        cpu = {"value": val, "unit": "Medges/s", "cores": cores, "kind": "port", "sample": sample, "newton_step": cpu_newton}

    if rank == 0:
        line = {"metric": "fp64 residual+Jacobian assembly throughput", "value": value, "unit": "Medges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": name, "species": n, "nodes": int(g.num_nodes), "edges": int(E_total), "l2": "inputs larger than L2 (tile stream >> 126 MB)" if E_total * 24 > 2.0e8 else "inputs fit L2",
                           "parallelism": f"node-owner partitions x{world}" if world > 1 else "single GPU", "setup_s": setup_s},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "newton_step": newton,
                "device_bytes": st.device_bytes()}
        print(json.dumps(line))
    st.close()
    if world > 1:
        dist.destroy_process_group()


def _stream_ptr(st):
    import ctypes as C

    p = C.c_void_p()
    st.L.vfvm_stream(st.h, C.byref(p))
    return p.value


def _planes(st):
    m = st.matrix_plane_counts() if hasattr(st, "matrix_plane_counts") else None
    if m:
        return m
    n = st.n
    return n * n, n * n


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def _traffic(workload):
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(workload)
    except Exception:
        return None


def _newton_step(st, system, U, tstep, world):
    import ctypes as C

    L, h = st.L, st.h
    st.set_vector(v._lib.VEC_OLDSOL, U)
    L.vfvm_copy_vector(h, v._lib.VEC_SOLUTION, v._lib.VEC_OLDSOL)
    L.vfvm_init_dirichlet(h, 0.0, 0.0)
    # symmetric positive definite Jacobians (pure linear diffusion) take CG, everything else BiCGStab; node-block Jacobi for systems
    fid = system.physics.flux.id if system.physics.flux is not None else 0
    spd = fid == ph.FLUX_DIFFUSION and system.physics.reaction is None
    krylov = v._lib.KRYLOV_CG if spd else v._lib.KRYLOV_BICGSTAB
    precon = v._lib.PRECON_JACOBI if system.num_species == 1 or spd else v._lib.PRECON_BLOCKJACOBI
    if os.environ.get("VFVM_BENCH_PRECON", "amg") == "amg":
        precon = v._lib.PRECON_AMG  # aggregation AMG (csrc/amg.cu); VFVM_BENCH_PRECON=jacobi gives the one-level baseline
    v._lib.check(h, L.vfvm_linsolve_setup(h, krylov, precon, 0))
    # W-cycle on the top coarse levels where it pays (measured): the single-GPU scalar problems cfg3 (3D Laplace,
    # two levels) and cfg2 (2D nonlinear Poisson, three levels); V-cycle everywhere else
    wdepth = 0
    if precon == v._lib.PRECON_AMG and world == 1 and system.num_species == 1:
        wdepth = 2 if (spd and system.grid.dim == 3) else (3 if (not spd and system.grid.dim == 2) else 0)
    wcycle = wdepth > 0
    if wcycle:
        opts = (C.c_double * 6)(*([float("nan")] * 5 + [float(wdepth)]))
        v._lib.check(h, L.vfvm_amg_set_options(h, opts, 6))
    iters, resn = C.c_int(), C.c_double()
    assert L.vfvm_assemble(h, 0.0, tstep, 0.0) == 0
    L.vfvm_linsolve(h, 0.0, 1.0e-10, 3, 0, C.byref(iters), C.byref(resn))  # warm-up: work vectors, NCCL channels
    t0 = time.perf_counter()
    rc = L.vfvm_assemble(h, 0.0, tstep, 0.0)
    assert rc == 0
    rc = L.vfvm_linsolve(h, 0.0, 1.0e-10, 5000, 0, C.byref(iters), C.byref(resn))
    ninf, n1 = C.c_double(), C.c_double()
    L.vfvm_newton_update(h, 1.0, C.byref(ninf), C.byref(n1))
    dt = time.perf_counter() - t0
    t = st.timings()
    return {"ms": dt * 1e3, "assemble_ms": float(t[0]), "linsolve_ms": float(t[1] + t[2]), "krylov": ("CG" if spd else "BiCGStab") + {v._lib.PRECON_JACOBI: "+Jacobi", v._lib.PRECON_BLOCKJACOBI: "+block-Jacobi", v._lib.PRECON_AMG: "+aggregation-AMG"}[precon] + (f" (W-cycle on levels 1-{wdepth})" if wcycle else ""), "reltol": 1e-10, "iters": iters.value,
            "resnorm": resn.value, "update_norm_inf": ninf.value, "rc": rc}


if __name__ == "__main__":
    main()
