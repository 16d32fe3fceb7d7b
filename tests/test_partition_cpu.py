"""Host logic of the multi-GPU path (no GPU): the owner-computes row partition, its halo exchange plan and the slices
of the multilevel hierarchy, checked through pgo_analyze_partition / pgo_amg_aggregates (csrc/pgo_b200.cu,
csrc/pgo_amg_host.hpp) and against the Python mirror datasets.partition_rows."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def D(pgo):
    return pgo.datasets


def _graphs(D):
    return [D.sphere(10, 20, None), D.manhattan_grid(30, 30, 60), D.torus(2000, winds=20), D.kitti00()]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_partition_plans_are_consistent_across_ranks(pgo, D, world):
    for g in _graphs(D):
        infos = [pgo.analyze_partition(g, r, world) for r in range(world)]
        assert all(i.consistent == 1 for i in infos), g.name
        assert sum(i.n_own for i in infos) == g.n_poses
        # every cut edge is evaluated by exactly its two owners, every other edge once
        assert sum(i.n_local_edges for i in infos) == g.n_edges + sum(i.n_cut_edges for i in infos) // 2
        assert sum(i.n_cut_edges for i in infos) % 2 == 0
        for a in range(world):
            for b in range(world):
                assert infos[a].send_to[b] == infos[b].recv_from[a], (g.name, a, b)
        # order-sensitive: what a sends to b, position by position, is what b files into its halo, on every level
        assert sum(i.plan_checksum for i in infos) % 2 ** 64 == sum(i.recv_checksum for i in infos) % 2 ** 64
        # the hierarchy is the same global object on every rank
        for i in infos[1:]:
            assert i.amg_levels == infos[0].amg_levels
            assert list(i.level_nodes) == list(infos[0].level_nodes)
        nl = infos[0].amg_levels
        for l in range(nl):
            if infos[0].level_replicated[l]:
                assert all(i.level_own[l] == infos[0].level_nodes[l] and i.level_halo[l] == 0 for i in infos)
            else:
                assert sum(i.level_own[l] for i in infos) == infos[0].level_nodes[l]


def test_partition_matches_python_mirror(pgo, D):
    g = D.manhattan_grid(30, 30, 60)
    for world in (2, 5):
        for r in range(world):
            info = pgo.analyze_partition(g, r, world)
            part = D.partition_rows(g, r, world)
            assert (info.n_own, info.n_halo, info.n_local_edges) == (part.n_own, len(part.halo_gid), part.local.n_edges)


def test_world_one_is_the_whole_graph(pgo, D):
    g = D.sphere(10, 20, None)
    info = pgo.analyze_partition(g, 0, 1)
    assert (info.n_own, info.n_halo, info.n_local_edges, info.n_cut_edges, info.n_neighbours) == (g.n_poses, 0, g.n_edges, 0, 0)
    assert info.consistent == 1 and info.amg_levels >= 2


def test_aggregates_are_connected_patches_that_shrink_the_graph(pgo, D):
    """Every aggregate is a connected set of poses; constant poses stay out; the hierarchy ends in a level small enough to
    be inverted densely (<= max(16, N / 16), at most 512 nodes: amg_host_params)."""
    import scipy.sparse as sp
    import scipy.sparse.csgraph as csg
    for g in (D.sphere(), D.manhattan_grid(60, 60, 200), D.torus(5000, winds=50)):
        sizes, aggs = pgo.amg_aggregates(g, 1)
        assert sizes[0] == g.n_poses and sizes[-1] <= min(512, max(16, g.n_poses // 16)) and len(sizes) >= 3
        assert all(sizes[k + 1] < 0.6 * sizes[k] for k in range(len(sizes) - 1)), sizes
        a0 = aggs[0]
        assert (a0[g.pose_const != 0] == -1).all() and (a0[g.pose_const == 0] >= 0).all()
        assert a0.max() + 1 == sizes[1]
        # connectivity of the level-0 patches in the pose graph
        n = g.n_poses
        adj = sp.coo_matrix((np.ones(g.n_edges), (g.edge_ids[:, 0], g.edge_ids[:, 1])), shape=(n, n))
        same = a0[g.edge_ids[:, 0]] == a0[g.edge_ids[:, 1]]
        inner = sp.coo_matrix((np.ones(same.sum()), (g.edge_ids[same, 0], g.edge_ids[same, 1])), shape=(n, n))
        ncomp, lab = csg.connected_components(inner, directed=False)
        variable = a0 >= 0
        # one connected component per aggregate (plus the singletons of the constant poses)
        assert len(np.unique(lab[variable])) == sizes[1]
        del adj


def test_partitioned_aggregates_do_not_cross_ranks(pgo, D):
    g = D.manhattan_grid(40, 40, 100)
    for world in (2, 4):
        sizes, aggs = pgo.amg_aggregates(g, world)
        a0 = aggs[0]
        owner = (np.arange(g.n_poses) * world) // g.n_poses      # contiguous ranges n r / W
        # exact ranges as the library computes them
        off = [(g.n_poses * r) // world for r in range(world + 1)]
        owner = np.searchsorted(off, np.arange(g.n_poses), side="right") - 1
        for agg_id in np.unique(a0[a0 >= 0]):
            assert len(np.unique(owner[a0 == agg_id])) == 1
