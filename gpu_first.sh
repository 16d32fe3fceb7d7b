cd /root/repo
python -m pytest tests -m gpu -q -k "not kitti00_matches" 2>&1 | tail -15
python - <<'PY'
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import numpy as np, posegraph_ceres_b200 as P, oracle_py as O
g=P.datasets.kitti00()
G=P.Graph.from_dataset(g)
for tol,maxit in ((1e-6,20000),(1e-10,200000)):
    G.set_poses(g.poses)
    o=P.default_options(); o.linear_solver_type=0; o.pcg_tolerance=tol; o.pcg_max_iterations=maxit
    t=time.time(); s,its=G.solve(o); dt=time.time()-t
    print("tol",tol,"time",dt,"iters",s.num_iterations,"pcg",s.total_pcg_iterations,"lin ms",s.time_linearize_ms,"solver ms",s.time_linear_solver_ms,s.message.decode(), s.final_cost)
    for it in its: print("  ",it.iteration,it.step_is_successful,it.cost,it.linear_solver_iterations,it.pcg_relative_residual,it.step_norm)
    ref,rs,rits=O.solve(g)
    print("  vs oracle pos diff", np.abs(G.get_poses()[:,:3]-ref[:,:3]).max(), "oracle iters", rs.num_iterations, rs.time_total_s)
c,ms=G.linearize(); print("linearize ms", ms)
x=np.random.default_rng(0).normal(size=(g.n_poses,6)); y,ms=G.spmv(x,None,100); print("spmv x100 ms",ms)
PY
