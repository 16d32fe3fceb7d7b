"""world_size-2 gloo test of the multi-rank host logic: under the owner-computes row partition every rank evaluates the
edges that touch its own poses (cut edges on both sides), keeps its own rows of J^T r and accounts the cost of the edges
whose id_begin it owns; gathering the owned rows and summing the costs reproduces the single-rank quantities (what the
row-partitioned CUDA path does with NCCL).  The halo plan of the C++ host code is cross-checked between the ranks."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import oracle_py as O
    import posegraph_ceres_b200 as P
    D = P.datasets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = D.sphere(6, 10, None)
    part = D.partition_rows(g, rank, world)
    loc = part.local
    # owned rows of the gradient from ALL local edges; cost from the edges whose id_begin is owned
    _, _, grad, _ = O.evaluate(loc)
    import dataclasses
    mine = dataclasses.replace(loc, edge_ids=loc.edge_ids[part.cost_edges], edge_meas=loc.edge_meas[part.cost_edges],
                               edge_sqrt_info=loc.edge_sqrt_info[part.cost_edges])
    cost = O.evaluate(mine)[0] if mine.n_edges else 0.0
    full = torch.zeros(g.n_poses, 6, dtype=torch.float64)
    full[part.lo:part.hi] = torch.from_numpy(grad[:part.n_own])
    dist.all_reduce(full, op=dist.ReduceOp.SUM)          # disjoint slices: the sum is the all-gather
    c = torch.tensor([cost], dtype=torch.float64)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    # halo plan of the C++ host code: counts all-gathered and compared pairwise
    info = P.analyze_partition(g, rank, world)
    plan = torch.tensor([list(info.send_to[:world]), list(info.recv_from[:world])], dtype=torch.int64)
    plans = [torch.zeros_like(plan) for _ in range(world)]
    dist.all_gather(plans, plan)
    sums = torch.tensor([info.plan_checksum % 2 ** 62, info.recv_checksum % 2 ** 62, info.plan_checksum >> 62, info.recv_checksum >> 62], dtype=torch.int64)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if rank == 0:
        full_cost, _, full_grad, _ = O.evaluate(g)
        sym = all(int(plans[a][0][b]) == int(plans[b][1][a]) for a in range(world) for b in range(world))
        send_sum = (int(sums[0]) + (int(sums[2]) << 62)) % 2 ** 64
        recv_sum = (int(sums[1]) + (int(sums[3]) << 62)) % 2 ** 64
        out.put((abs(c.item() - full_cost), float(np.abs(full.numpy() - full_grad).max()), sym, send_sum == recv_sum,
                 info.consistent, info.n_own, len(part.halo_gid), info.n_halo))
    dist.destroy_process_group()


def test_row_partition_owner_computes_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    dc, dg, sym, sums_ok, consistent, n_own, n_halo_py, n_halo_c = out.get(timeout=10)
    assert dc < 1e-9 and dg < 1e-9
    assert sym and sums_ok and consistent == 1
    assert n_own == 30 and n_halo_py == n_halo_c and n_halo_c > 0
