"""diagnostic: the Newton history of tests/test_gpu_parity.py::test_amg_block_system_bipolar_newton under the switches given in the environment"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfvm_b200 as v
from vfvm_b200 import physics as ph

X = np.linspace(0, 1, 14)
g = v.simplexgrid(X, X, X)
bc = ph.BCondition()
for sp, val in ((1, 0.0), (2, 0.0), (3, 0.5)):
    bc.dirichlet(species=sp, region=5, value=val)
for sp, val in ((1, 0.1), (2, 0.1), (3, 0.2)):
    bc.dirichlet(species=sp, region=6, value=val)
s = v.System(g, flux=ph.BipolarSGFlux(), reaction=ph.BipolarReaction([1.0]), storage=ph.BipolarStorage(), bcondition=bc, species=[1, 2, 3])
every = os.environ.get("EVERY", "0") == "1"
sols = {}
for name, pc in (("block", v.BlockPreconBuilder()), ("amg", v.AMGPreconBuilder())):
    try:
        sols[name] = v.solve(s, inival=0.1, tstep=1.0e-2, method_linear=v.KrylovJL_BICGSTAB(precs=pc), reltol_linear=1e-13, abstol_linear=0.0, maxiters_linear=3000, verbose="n",
                             factorize_every_newtonstep=every)
        print(name, "converged", flush=True)
    except Exception as e:
        print(name, "FAILED", repr(e), flush=True)
if len(sols) == 2:
    print("diff", np.abs(sols["block"] - sols["amg"]).max())
