#!/usr/bin/env python3
"""The multilevel-PCG solver (PGO_LINEAR_PCG_AMG) on the mesh-like configurations, one GPU or -- under torchrun, one rank
per GPU -- row-partitioned over NCCL:
   python tools/amg_check.py [--large] [--cases sphere,grid100,torus5k,...]
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/amg_check.py
Reports LM / PCG iterations, times, parity against the CPU oracle (where it finishes in seconds) and, multi-GPU, the
agreement with the one-GPU poses and between ranks."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402
from helpers import rot_angle_between  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--large", action="store_true")
ap.add_argument("--cases", default="")
ap.add_argument("--tol", type=float, default=0.0)
ap.add_argument("--no-oracle", action="store_true")
ap.add_argument("--verbose", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--no-1gpu", action="store_true", help="multi-GPU: skip the one-GPU comparison solve on rank 0")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

D = P.datasets
cases = {
    "sphere200": lambda: D.sphere(10, 20, None),
    "sphere": lambda: D.sphere(),
    "grid100": lambda: D.manhattan_grid(100, 100, 500),
    "torus5k": lambda: D.torus(5000, winds=50),
    "torus100k": lambda: D.torus(100000),
    "grid1000": lambda: D.manhattan_grid(1000, 1000, 50000),
}
default = ["sphere200", "sphere", "grid100", "torus5k"] + (["torus100k", "grid1000"] if args.large else [])
names = [c for c in args.cases.split(",") if c] or default
oracle_ok = {"sphere200", "sphere", "grid100", "torus5k"}
ok = True
for name in names:
    g = cases[name]()
    uid = None
    if world > 1:
        box = [P.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    t0 = time.perf_counter()
    G = P.Graph.from_dataset(g, device=local, unique_id=uid, rank=rank, world=world)
    t_create = time.perf_counter() - t0
    o = P.default_options()
    o.linear_solver_type = P.LINEAR_PCG_AMG
    if args.tol > 0:
        o.pcg_tolerance = args.tol
    o.verbose = args.verbose if rank == 0 else 0
    G.snapshot_poses()
    best = None
    for rep in range(args.reps):
        G.restore_poses()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s, its = G.solve(o)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    poses = G.get_poses()
    own, halo, ledges = G.local_sizes()
    G.close()
    lm = s.num_iterations - 1
    line = (f"{name}: world={world} N={g.n_poses} E={g.n_edges} | create {1e3 * t_create:.1f} ms, solve {1e3 * best:.1f} ms, {lm} LM iterations, "
            f"{s.total_pcg_iterations} PCG iterations ({s.total_pcg_iterations / max(lm, 1):.1f}/LM), amg levels {s.amg_levels} blocks {s.amg_blocks}, "
            f"cost {s.initial_cost:.4f} -> {s.final_cost:.6f} ({s.message.decode()[:40]}), linearize {s.time_linearize_ms:.1f} ms, solver {s.time_linear_solver_ms:.1f} ms, "
            f"launches {s.kernel_launches}")
    if world > 1:
        line += (f" | rank0 own {own} halo {halo} edges {ledges}, comm {s.comm_calls} calls {s.comm_bytes / 1e6:.2f} MB, per PCG it "
                 f"{s.comm_calls_per_pcg_iteration} calls {s.comm_bytes_per_pcg_iteration / 1e3:.1f} KB")
        t = torch.from_numpy(poses).cuda()
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        spread = float((tmax - tmin).abs().max())
        line += f" | rank spread {spread:.1e}"
        ok &= spread == 0.0
    if rank == 0:
        if world > 1 and not args.no_1gpu:
            G1 = P.Graph.from_dataset(g, device=local)
            s1, _ = G1.solve(o)
            p1 = G1.get_poses()
            G1.close()
            d1 = np.abs(poses - p1).max()
            line += f" | 1-GPU: {s1.num_iterations - 1} LM, {s1.total_pcg_iterations} PCG, |p - p_1gpu| {d1:.2e}"
            ok &= d1 <= 1e-6
        if name in oracle_ok and not args.no_oracle:
            import oracle_py as O
            O.set_num_threads(len(os.sched_getaffinity(0)))
            t0 = time.perf_counter()
            ref, rs, rits = O.solve(g)
            cdt = time.perf_counter() - t0
            dp = np.abs(poses[:, :3] - ref[:, :3]).max()
            dr = rot_angle_between(poses[:, 3:], ref[:, 3:]).max()
            good = dp <= 1e-4 and dr <= 1e-4 and s.termination_type == rs.termination_type
            ok &= good
            line += (f" | oracle {1e3 * cdt:.0f} ms {rs.num_iterations - 1} LM final {rs.final_cost:.6f}, max |dp| {dp:.2e} m angle {dr:.2e} rad "
                     f"-> {'OK' if good else 'MISMATCH'} speed-up {cdt / best:.1f}x")
        print(line, flush=True)
        if args.verbose:
            for it in its:
                print(f"    it {it.iteration:3d} cost {it.cost:.6e} radius {it.trust_region_radius:.2e} pcg {it.linear_solver_iterations} rel {it.pcg_relative_residual:.1e} "
                      f"ok {it.step_is_successful}")
if world > 1:
    dist.barrier()
if rank == 0:
    print("AMG_CHECK", "PASS" if ok else "FAIL", flush=True)
if world > 1:
    dist.destroy_process_group()
