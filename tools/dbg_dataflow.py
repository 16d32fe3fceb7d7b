import os, sys
os.environ["PGO_TIMELINE"]="1"
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import posegraph_ceres_b200 as P
import oracle_py as O
g=P.datasets.kitti00() if len(sys.argv)>1 and sys.argv[1]=='kitti' else P.datasets.manhattan_loop()
G=P.Graph.from_dataset(g)
G.linearize(loss_type=1, loss_a=1.0)
_,_,_,grad=G.hessian()
_,_,ograd,ojac=O.evaluate(g,loss_type=1,loss_a=1.0)
d=np.random.default_rng(2).uniform(1e-3,1e-2,(g.n_poses,6))
rc,yref=O.normal_solve(g,ojac,d.ravel(),ograd.ravel())
o=P.default_options(); o.linear_solver_type=1; o.pcg_tolerance=1e-12
y,it,rel,ms=G.linear_solve(d,grad,o)
print('iters',it,'rel',rel,'ms',ms,'maxdiff',np.abs(y-yref).max(), 'nan', np.isnan(y).sum(), flush=True)
info=P.analyze_structure(g.n_poses,g.edge_ids,g.pose_const); print(info.factor_levels, info.factor_blocks, info.factor_max_degree)
