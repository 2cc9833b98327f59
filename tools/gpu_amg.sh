#!/bin/bash
export VFVM_AMG_VERBOSE=1 VFVM_AMG_ALPHA=1.75
WL=cfg1 METHODS="cg+jacobi,cg+amg" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -4
WL=cfg2 METHODS="bicgstab+jacobi,bicgstab+amg,cg+amg" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -5
WL=cfg4 METHODS="bicgstab+block,bicgstab+amg,gmres+amg" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -5
WL=cfg5 METHODS="cg+jacobi,cg+amg" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -4
