import sys, math, numpy as np
sys.path.insert(0,'.')
import vfvm_b200 as v
from vfvm_b200 import physics as ph
from oracle import oracle as O
X=np.linspace(0,1,8); g=v.simplexgrid(X,X,X)
v.cellmask(g,[0,0,0.3],[1,1,0.72],2); v.cellmask(g,[0,0,0.7],[1,1,1.0],3)
bc=ph.BCondition()
for sp,val in ((1,0.0),(2,0.0),(3,0.5+math.asinh(10.0/(2*math.sqrt(math.exp(-1.0)))))): bc.dirichlet(species=sp,region=5,value=val)
for sp,val in ((1,0.3),(2,0.3),(3,0.5+math.asinh(-10.0/(2*math.sqrt(math.exp(-1.0))))+0.3)): bc.dirichlet(species=sp,region=6,value=val)
s=v.System(g,flux=ph.BipolarSGFlux(),reaction=ph.BipolarReaction([10.0,0.0,-10.0]),storage=ph.BipolarStorage(),bcondition=bc,species=[1,2,3])
rng=np.random.default_rng(20261017); U=np.asfortranarray(rng.uniform(-0.5,0.5,(3,g.num_nodes))); U[2,:]=np.linspace(3.0,-3.0,g.num_nodes)
st=v.SystemState(s); F=st.eval_res_jac(U,U,tstep=1e-2); A=st.matrix('csc')
Fo,Ao=O.OracleSystem(s).assemble(U,U,tstep=1e-2)
coo=Ao.tocoo(); err=np.abs(A.data-Ao.data); rel=err/np.maximum(np.abs(Ao.data),1e-300)
off=(coo.row//3)!=(coo.col//3)
idx=np.argsort(-np.where(off,rel,0))[:10]
for k in idx: print(coo.row[k],coo.col[k],A.data[k],Ao.data[k],rel[k])
mag=np.where(np.abs(coo.data)<1e29,np.abs(coo.data),0.0); ts=np.bincount(coo.row,weights=mag,minlength=Ao.shape[0])
print("max err/termscale", (err/np.maximum(ts[coo.row],1e-300)).max(), "count rel>1e-12:", (rel[off]>1e-12).sum(), "of", off.sum())
