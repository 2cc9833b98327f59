#!/bin/bash
WL=cfg3 METHODS="cg+amg+WDEPTH=0,cg+amg+WDEPTH=2" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
WL=cfg2 METHODS="bicgstab+amg+WDEPTH=0,bicgstab+amg+WDEPTH=3" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0,gmres+amg" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
WL=cfg5 METHODS="cg+amg+WDEPTH=0,cg+amg+WDEPTH=2" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
WL=cfg1 METHODS="cg+amg+WDEPTH=0,cg+amg+WDEPTH=2" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -2
VFVM_AMG_NO_GRAPH=1 WL=cfg1 METHODS="cg+amg+WDEPTH=0" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -1
