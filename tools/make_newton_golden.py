#!/usr/bin/env python
"""Generates tests/golden/newton_samples_<workload>_<nx>.json: the solution after ONE Newton step of a bench workload at a fixed set of
sample nodes, used by bench.py's `parity.newton` and tests/test_gpu_large.py to check the device Newton step at full size.

    python tools/make_newton_golden.py --workload cfg3 --nx 193 --source oracle     # CPU: oracle assembly + oracle Krylov (minutes)
    python tools/make_newton_golden.py --workload cfg4 --nx 193 --source device     # one B200: device solve at reltol 1e-13

`--source oracle` is independent of the CUDA path (assembly = the restatement of the reference's edge loop, solve = the oracle's
multithreaded CG / BiCGStab driven to its attainable accuracy).  `--source device` is a self-consistency pin for sizes the CPU cannot
solve in reasonable time; the file records which it was.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NSAMPLES = 4096


def sample_nodes(num_nodes: int) -> np.ndarray:
    rng = np.random.default_rng(20261017)
    return np.sort(rng.choice(num_nodes, size=min(NSAMPLES, num_nodes), replace=False)).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--source", default="oracle", choices=["oracle", "device"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--reltol", type=float, default=None)
    a = ap.parse_args()
    system, kw, name = bench.make_system(a.workload, a.nx)
    nx = a.nx or bench.DEFAULT_NX[a.workload]
    U = bench.generic_state(system)
    nodes = sample_nodes(system.grid.num_nodes)
    t0 = time.time()
    if a.source == "oracle":
        from oracle import oracle as O

        o = O.OracleSystem(system)
        nth = len(os.sched_getaffinity(0))
        U0 = o.initialize(U)
        F, _ = o.assemble(U0, U, tstep=kw["tstep"], nthreads=nth, want_matrix=False)
        spd = bench.is_spd(system)
        reltol = a.reltol or 1.0e-14
        x, it, rel, sec = o.krylov_solve(F, "cg" if spd else "bicgstab", "jacobi" if system.num_species == 1 else "blockjacobi", reltol=reltol, maxiters=20000, nthreads=nth)
        sol = U0 - x.reshape(U0.shape, order="F")
        how = f"CPU oracle: assembly (restated edge loop) + {'CG+Jacobi' if spd else 'BiCGStab+node-block-Jacobi'} on {nth} threads, {it} iterations, true relative residual {rel:.2e}"
    else:
        import vfvm_b200 as v

        st = v.SystemState(system, device=0)
        sol, info = bench.newton_solution(st, system, U, kw["tstep"], reltol=a.reltol or 1.0e-13)
        st.close()
        how = f"device (1 GPU): {info}"
    out = a.out or os.path.join(ROOT, "tests", "golden", f"newton_samples_{a.workload}_{nx}.json")
    rec = {"workload": name, "nx": nx, "source": a.source, "how": how, "seconds": time.time() - t0, "nodes": nodes.tolist(),
           "solution": [[float(repr(float(sol[i, k]))) for k in nodes] for i in range(system.num_species)]}
    with open(out, "w") as f:
        json.dump(rec, f)
    print(f"wrote {out}: {how}")


if __name__ == "__main__":
    main()
