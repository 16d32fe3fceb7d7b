"""world_size-2 gloo test of the multi-rank host logic: contiguous edge shards + all-reduce(sum) of
the per-shard J^T r / cost reproduce the single-rank quantities (what the NCCL path does on GPUs)."""
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
    import posegraph_ceres_b200.datasets as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = D.sphere(6, 10, None)
    shard = D.shard_edges(g, rank, world)
    cost, _, grad, _ = O.evaluate(shard)
    t = torch.from_numpy(np.concatenate([[cost], grad.ravel()]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        full_cost, _, full_grad, _ = O.evaluate(g)
        out.put((abs(t[0].item() - full_cost), float(np.abs(t[1:].numpy() - full_grad.ravel()).max()), shard.n_edges, g.n_edges))
    dist.destroy_process_group()


def test_edge_sharding_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    dc, dg, ne, total = out.get(timeout=10)
    assert dc < 1e-9 and dg < 1e-9 and ne == total // 2
