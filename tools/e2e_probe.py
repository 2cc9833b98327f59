"""e2e probe: PCIe one-way / two-way bandwidth with pinned buffers, and vfvm_eval_res_jac(host) for several chunk counts."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vfvm_b200 as v  # noqa: E402

nbytes = 57512456
h1 = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
h2 = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
d1 = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
d2 = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn in (("H2D", h2d), ("D2H", d2h), ("H2D+D2H concurrently", both)):
    ms = timeit(fn)
    print(f"{name}: {ms:.3f} ms  ({nbytes / ms / 1e6:.1f} GB/s per direction)")

system, kw, name = bench.make_system("cfg3", None)
st = v.SystemState(system)
U = bench.generic_state(system)
nd = st.n * st.N
hU = torch.empty(nd, dtype=torch.float64).pin_memory()
hF = torch.empty(nd, dtype=torch.float64).pin_memory()
hU.numpy()[:] = U.ravel(order="F")
L, h = st.L, st.h
for env in [{"VFVM_NO_PIPELINE": "1"}] + [{"VFVM_PIPE_CHUNKS": str(k), "VFVM_PIPE_GRID_PCT": str(p)} for k in (12, 16, 24, 32) for p in (20, 30, 40)]:
    for k in ("VFVM_NO_PIPELINE", "VFVM_PIPE_CHUNKS", "VFVM_PIPE_GRID_PCT"):
        os.environ.pop(k, None)
    os.environ.update(env)

    def step():
        assert L.vfvm_eval_res_jac(h, hU.data_ptr(), None, hF.data_ptr(), v._lib.HOST, 0.0, kw["tstep"], 0.0) == 0

    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    print(env, f"{(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
