import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vfvm_b200 as v
from vfvm_b200 import partition as P, physics as ph, _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx = int(os.environ.get("NX", "129"))
X = np.linspace(0, 1, nx)
s = v.System(v.simplexgrid(X, X, X), flux=ph.LinearDiffusion(), source=ph.XSinYExpZSource(1, 5.0))
v.enable_species(s, 1, [1]); v.boundary_dirichlet(s, 1, 5, 0.0); v.boundary_dirichlet(s, 1, 6, 0.0)
st, info = P.partitioned_state(s, rank, world, local)
L, h = st.L, st.h
U = np.asfortranarray(np.full((1, info.local_nodes.size), 0.5))
st.set_vector(0, U); st.set_vector(1, U)
L.vfvm_init_dirichlet(h, 0.0, 0.0)
assert L.vfvm_assemble(h, 0.0, float("inf"), 0.0) == 0
def timeit(name, fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    if rank == 0: print(f"{name:30s} {dt*1e6:9.1f} us", flush=True)
timeit("halo_exchange(+sync)", lambda: L.vfvm_halo_exchange(h, 0))
a, b = C.c_double(), C.c_double()
timeit("vector_norms (2 allreduce+sync)", lambda: L.vfvm_vector_norms(h, 0, C.byref(a), C.byref(b)))
x = np.ones(info.local_nodes.size); 
xd = torch.ones(info.local_nodes.size, dtype=torch.float64, device="cuda"); yd = torch.zeros(info.n_owned, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
timeit("spmv device (halo+kernel+sync)", lambda: L.vfvm_spmv(h, xd.data_ptr(), yd.data_ptr(), _lib.DEVICE))
it, rn = C.c_int(), C.c_double()
L.vfvm_linsolve_setup(h, 0, 1, 0)
for maxit in (100, 400, 800):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    L.vfvm_linsolve(h, 0.0, 1e-30, maxit, 0, C.byref(it), C.byref(rn))
    dt = time.perf_counter() - t0
    if rank == 0: print(f"linsolve {it.value} iters: {dt*1e3:.1f} ms -> {dt/it.value*1e6:.1f} us/iter", flush=True)
st.close(); dist.destroy_process_group()
