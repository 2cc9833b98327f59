import os, sys, time, ctypes as C, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfvm_b200 as v
from vfvm_b200 import _lib
import bench
wl = os.environ.get("WL", "cfg3"); nx = int(os.environ["NX"]) if "NX" in os.environ else None
system, kw, name = bench.make_system(wl, nx)
st = v.SystemState(system)
U = bench.generic_state(system)
st.set_vector(0, U); st.set_vector(1, U)
L, h = st.L, st.h
L.vfvm_copy_vector(h, 0, 1); L.vfvm_init_dirichlet(h, 0.0, 0.0)
assert L.vfvm_assemble(h, 0.0, kw["tstep"], 0.0) == 0
print(name, "asm ms", st.timings()[0])
it, rn = C.c_int(), C.c_double()
K = {"bicgstab": 0, "cg": 1, "gmres": 2}; P = {"none": 0, "jacobi": 1, "block": 2, "ilu0": 3, "ilu0mc": 4, "amg": 5}
for spec in os.environ.get("METHODS", "cg+jacobi,bicgstab+jacobi,cg+ilu0mc,bicgstab+ilu0mc,gmres+ilu0mc").split(","):
    k, p = spec.split("+")[:2]
    for kv in spec.split("+")[2:]:  # e.g. cg+amg+ALPHA=1.8+OMEGA=0.9
        key, val = kv.split("=")
        os.environ["VFVM_AMG_" + key] = val
    _lib.check(h, L.vfvm_linsolve_setup(h, K[k], P[p], 50))
    L.vfvm_linsolve(h, 0.0, 1e-10, 2, 0, C.byref(it), C.byref(rn))
    t0 = time.perf_counter()
    rc = L.vfvm_linsolve(h, 0.0, 1e-10, int(os.environ.get("MAXIT", "5000")), 0, C.byref(it), C.byref(rn))
    dt = time.perf_counter() - t0
    t = st.timings()
    x = st.get_vector(3)[:, : st.Nown]
    if "xref" not in globals():
        xref = x.copy()
    dx = np.abs(x - xref).max() / max(np.abs(xref).max(), 1e-300)
    print(f"{k:9s}+{p:7s} {spec} rc={rc} dx={dx:.1e} iters={it.value:5d} res={rn.value:.2e} setup={t[1]:9.2f} ms solve={t[2]:9.2f} ms  ({t[2]/max(1,it.value):.3f} ms/it) wall={dt*1e3:.1f}", flush=True)
