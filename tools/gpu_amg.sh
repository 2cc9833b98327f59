#!/bin/bash
WL=cfg3 METHODS="cg+amg+WDEPTH=2+ALPHA=2.0,cg+amg+WDEPTH=2+ALPHA=1.75+OMEGA=0.9,cg+amg+WDEPTH=2+OMEGA=0.8+COARSE_SWEEPS=2" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -3
export VFVM_AMG_COARSE_SWEEPS=4
WL=cfg2 METHODS="bicgstab+amg+WDEPTH=0,bicgstab+amg+WDEPTH=2,bicgstab+amg+WDEPTH=3" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -3
WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0,bicgstab+amg+WDEPTH=2" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
WL=cfg5 METHODS="cg+amg+WDEPTH=0,cg+amg+WDEPTH=2" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
WL=cfg1 METHODS="cg+amg+WDEPTH=0,cg+amg+WDEPTH=2" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
