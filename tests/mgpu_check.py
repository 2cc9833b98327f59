"""Multi-GPU parity script (not collected by pytest): run as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_check.py
Each rank assembles / solves its node-owner partition on its own GPU (NCCL halo + all-reduce); rank 0 compares the
gathered result with the CPU oracle on the whole grid (Newton solution <= 1e-10, residual rows <= 1e-12)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import vfvm_b200 as v
from vfvm_b200 import partition as P
from vfvm_b200 import physics as ph


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    X = np.linspace(0, 1, 17)
    s = v.System(v.simplexgrid(X, X, X), flux=ph.PowerDiffusion(1.0e-1, 2), reaction=ph.PowerReaction(1.0, 2.0), source=ph.XSinYExpZSource(1, 5.0), storage=ph.LinearStorage(1.0))
    v.enable_species(s, 1, [1])
    v.boundary_dirichlet(s, 1, 5, 0.1)
    v.boundary_dirichlet(s, 1, 6, 0.2)
    st, info = P.partitioned_state(s, rank, world, local)
    N = s.grid.num_nodes
    rng = np.random.default_rng(1)
    Ug = np.asfortranarray(rng.uniform(0.1, 1.0, (1, N)))
    # ---- residual rows
    F = st.eval_res_jac(Ug[:, info.local_nodes], tstep=0.05)
    # ---- one implicit Euler step (Newton to convergence) with BiCGStab + Jacobi across ranks
    sol = v.solve_state(st, inival=np.asfortranarray(Ug[:, info.local_nodes]), tstep=0.05)
    # the same step with rank-local aggregation AMG inside BiCGStab (block-Jacobi across ranks, no communication in the preconditioner)
    sol_amg = v.solve_state(st, inival=np.asfortranarray(Ug[:, info.local_nodes]), tstep=0.05, method_linear=v.KrylovJL_BICGSTAB(precs=v.AMGPreconBuilder()),
                            reltol_linear=1e-13, abstol_linear=0.0, maxiters_linear=2000)
    own = slice(0, info.n_owned)
    assert np.max(np.abs(sol_amg[:, own] - sol[:, own])) < 1e-10, "AMG-preconditioned solve differs"

    gathered = [None] * world
    dist.all_gather_object(gathered, (info.local_nodes[own], F[:, own], sol[:, own]))
    if rank == 0:
        from oracle import oracle as O

        Fg, solg = np.zeros((1, N)), np.zeros((1, N))
        for nodes, f, u in gathered:
            Fg[:, nodes] = f
            solg[:, nodes] = u
        o = O.OracleSystem(s)
        Fo, _ = o.assemble(Ug, Ug, tstep=0.05)
        ref = o.solve_step(Ug, tstep=0.05)
        ef = np.max(np.abs(Fg - Fo) / np.maximum(np.abs(Fo), 1e-3))
        eu = np.max(np.abs(solg - ref))
        print(f"mgpu_check world={world} transport={'peer mailboxes' if st.peer else 'NCCL'}: residual rel err {ef:.2e}, Newton solution err {eu:.2e}")
        assert ef < 1e-11 and eu < 1e-10
        print("MGPU_OK")
    st.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
