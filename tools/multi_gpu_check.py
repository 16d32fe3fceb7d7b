#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU): edge-sharded solve over NCCL vs the single-GPU solve and the CPU oracle.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, g in (("sphere60", P.datasets.sphere(6, 10, None)), ("sphere200", P.datasets.sphere(10, 20, None))):
    shard = P.datasets.shard_edges(g, rank, world)
    def stage(msg):
        print(f"[rank {rank}] {name}: {msg}", file=sys.stderr, flush=True)
    stage("create")
    G = P.Graph.from_dataset(shard, device=local)
    uid = [P.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    stage("init_comm")
    G.init_comm(uid[0], rank, world)
    stage("evaluate")
    # Problem::Evaluate across shards
    cost, _, grad, _ = G.evaluate()
    stage("solve")
    o = P.default_options()
    o.pcg_tolerance = 1e-12
    o.pcg_max_iterations = 100000
    o.verbose = 1 if rank == 0 else 0
    import time
    t0 = time.perf_counter()
    s, its = G.solve(o)
    if rank == 0:
        print(f"{name}: sharded solve took {time.perf_counter() - t0:.2f} s, {s.total_pcg_iterations} PCG iterations", flush=True)
    stage("solved")
    poses = G.get_poses()
    G.close()
    stage("closed")
    # all ranks must hold identical poses
    t = torch.from_numpy(poses).cuda()
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    spread = float((tmax - tmin).abs().max())
    if rank == 0:
        import oracle_py as O
        ocost, _, ograd, _ = O.evaluate(g)
        ref, rs, rits = O.solve(g)
        G1 = P.Graph.from_dataset(g, device=local)
        o1 = P.default_options()
        o1.linear_solver_type = P.LINEAR_PCG_BLOCK_JACOBI
        o1.pcg_tolerance = 1e-12
        o1.pcg_max_iterations = 100000
        s1, _ = G1.solve(o1)
        p1 = G1.get_poses()
        G1.close()
        d_or = np.abs(poses[:, :3] - ref[:, :3]).max()
        d_1 = np.abs(poses - p1).max()
        good = (abs(cost - ocost) <= 1e-10 * max(1, ocost) and np.abs(grad - ograd).max() <= 1e-9 * max(1, np.abs(ograd).max())
                and d_or <= 1e-4 and spread == 0.0 and s.termination_type == rs.termination_type and len(its) == len(rits))
        ok &= good
        print(f"{name}: world={world} cost {cost:.9f} (oracle {ocost:.9f}) iterations {len(its) - 1} (oracle {len(rits) - 1}, 1-GPU {s1.num_iterations - 1}) "
              f"pcg {s.total_pcg_iterations} |p - oracle| {d_or:.2e} m |p - 1GPU| {d_1:.2e} rank spread {spread:.1e} -> {'OK' if good else 'MISMATCH'}", flush=True)
dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
