#!/usr/bin/env python3
"""CPU prototype (numpy/scipy + the oracle's Jacobians) of the aggregation-AMG preconditioner for the damped normal
equations of a pose graph.  Design evidence only: it answers "how many PCG iterations does a rigid-body-mode
aggregation cycle need on sphere / grid / torus" before the CUDA version is written.  Not part of the product.

Experiments recorded in DESIGN.md section 3.3 (all with --agg cxx --omega 0.85, i.e. the library's own aggregates):
  --dense-below N          solve the first level of <= N nodes exactly           (built: amg_dense_gj_kernel)
  --gamma 2 [--gamma-depth d]   visit levels 1..d twice (truncated W-cycle)      (built: amg_vcycle)
  --at-truth               linearise at the optimum: the long-range loop edges carry full weight there (Huber
                           down-weights them at the noisy initial poses) and the 1000 x 1000 grid needs 214 iterations
                           instead of 86; --loops N varies their number
  --smooth-p 0.66          smoothed aggregation (prolongator smoothing with the full matrix): 15-20 iterations
                           everywhere, operator complexity 6-16 unfiltered                          (next step)
  --smooth-p 0.66 --smooth-filter   the same with a filtered matrix (near blocks, far ones lumped rigidly into the
                           diagonal): complexity 2.5-4, but 29 instead of 16 iterations on the 100 x 100 grid at the optimum
                           and PCG breaks down on the 200^2 / 400^2 grids there; --filter-nolump (far blocks dropped,
                           full diagonal kept, i.e. rigid motions no longer reproduced at the loop ends): 86 / 177 / 258
  --pair-levels k, --local-scale s, --geo-smooth w     12x12 pair smoother over the loop edges, rescaled local part of
                           the coarse operators, topology-only prolongator weights: tried, none helps
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import posegraph_ceres_b200.datasets as D  # noqa: E402


def assemble(g, jac):
    E, N = g.n_edges, g.n_poses
    Ja = jac[:, 0].reshape(E, 6, 6)
    Jb = jac[:, 1].reshape(E, 6, 6)
    a, b = g.edge_ids[:, 0], g.edge_ids[:, 1]
    rows = np.concatenate([a, b, a, b])
    cols = np.concatenate([a, b, b, a])
    blk = np.concatenate([np.einsum("eki,ekj->eij", Ja, Ja), np.einsum("eki,ekj->eij", Jb, Jb),
                          np.einsum("eki,ekj->eij", Ja, Jb), np.einsum("eki,ekj->eij", Jb, Ja)])
    # expand to COO scalars
    r = (rows[:, None, None] * 6 + np.arange(6)[None, :, None]) + np.zeros((1, 1, 6), int)
    c = (cols[:, None, None] * 6 + np.arange(6)[None, None, :]) + np.zeros((1, 6, 1), int)
    H = sp.coo_matrix((blk.ravel(), (r.ravel(), c.ravel())), shape=(6 * N, 6 * N)).tocsr()
    return H


def aggregate(adj, n, target=None, rng=None):
    """Greedy aggregation: pass 1 roots with all-unaggregated neighbourhoods, pass 2 attach leftovers to a neighbour
    aggregate, pass 3 leftovers form their own."""
    agg = -np.ones(n, int)
    indptr, indices = adj.indptr, adj.indices
    na = 0
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = indices[indptr[i]:indptr[i + 1]]
        if np.all(agg[nb] < 0):
            agg[i] = na
            agg[nb] = na
            na += 1
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = indices[indptr[i]:indptr[i + 1]]
        cand = agg[nb]
        cand = cand[cand >= 0]
        if len(cand):
            agg[i] = -2 - cand[0]
    m = agg <= -2
    agg[m] = -agg[m] - 2
    for i in range(n):
        if agg[i] < 0:
            agg[i] = na
            na += 1
    return agg, na


def pairwise_aggregate(adj_w, n, passes=2):
    """Notay-style pairwise matching by strongest connection, repeated `passes` times (aggregates of <= 2^passes)."""
    agg = np.arange(n)
    na = n
    A = adj_w
    for _ in range(passes):
        A = A.tocsr()
        m = A.shape[0]
        match = -np.ones(m, int)
        new = -np.ones(m, int)
        k = 0
        deg = np.diff(A.indptr)
        for i in np.argsort(deg, kind="stable"):
            if new[i] >= 0:
                continue
            nb = A.indices[A.indptr[i]:A.indptr[i + 1]]
            w = A.data[A.indptr[i]:A.indptr[i + 1]]
            best, bw = -1, 0.0
            for j, ww in zip(nb, w):
                if j != i and new[j] < 0 and ww > bw:
                    best, bw = j, ww
            new[i] = k
            if best >= 0:
                new[best] = k
            k += 1
        agg = new[agg]
        P = sp.coo_matrix((np.ones(m), (np.arange(m), new)), shape=(m, k)).tocsr()
        A = (P.T @ A @ P).tocsr()
        A.setdiag(0)
        A.eliminate_zeros()
        na = k
    return agg, na


def block_diag_inv(A, n):
    """inverse of the 6x6 diagonal blocks -> BSR"""
    Ab = A.tobsr(blocksize=(6, 6))
    Dinv = np.zeros((n, 6, 6))
    rows = np.repeat(np.arange(n), np.diff(Ab.indptr))
    sel = np.nonzero(Ab.indices == rows)[0]
    Dinv[rows[sel]] = np.linalg.inv(Ab.data[sel])
    return sp.bsr_matrix((Dinv, np.arange(n), np.arange(n + 1)), shape=(6 * n, 6 * n)).tocsr()


def pair_block_diag_inv(A, n, pos, theta, levels_left):
    """Block-Jacobi inverse whose blocks are 12x12 for pairs of nodes joined by a geometrically WEAK (long-range) edge
    (greedy matching, strongest algebraic coupling first), 6x6 elsewhere."""
    Ab = A.tobsr(blocksize=(6, 6))
    rows = np.repeat(np.arange(n), np.diff(Ab.indptr))
    cols = Ab.indices
    off = rows != cols
    d2 = np.sum((pos[rows] - pos[cols]) ** 2, axis=1) + 1e-12
    wgt = np.where(off, 1.0 / d2, 0.0)
    W = sp.csr_matrix((wgt, (rows, cols)), shape=(n, n))
    rmax = W.max(axis=1).toarray().ravel()
    weak = off & (wgt < theta * np.maximum(rmax[rows], rmax[cols])) & (rows < cols)
    nrm = np.linalg.norm(Ab.data.reshape(-1, 36), axis=1)
    cand = np.nonzero(weak)[0]
    cand = cand[np.argsort(-nrm[cand])]
    mate = -np.ones(n, int)
    for k in cand:
        i, j = rows[k], cols[k]
        if mate[i] < 0 and mate[j] < 0:
            mate[i] = j; mate[j] = i
    # dense lookup of blocks
    diag_sel = np.nonzero(~off)[0]
    Dblk = np.zeros((n, 6, 6)); Dblk[rows[diag_sel]] = Ab.data[diag_sel]
    key = {}
    for k in np.nonzero(off & (mate[rows] == cols))[0]:
        key[(rows[k], cols[k])] = Ab.data[k]
    single = np.nonzero(mate < 0)[0]
    inv_single = np.linalg.inv(Dblk[single])
    # singles as BSR
    S = sp.bsr_matrix((inv_single, single, np.arange(len(single) + 1)), shape=(6 * len(single), 6 * n))
    R = sp.csr_matrix((np.ones(6 * len(single)), (np.repeat(single, 6) * 6 + np.tile(np.arange(6), len(single)), np.arange(6 * len(single)))), shape=(6 * n, 6 * len(single)))
    Dinv = (R @ S).tocsr()
    pairs = [(i, mate[i]) for i in range(n) if mate[i] > i]
    if pairs:
        pi = np.array([p[0] for p in pairs]); pj = np.array([p[1] for p in pairs])
        B = np.zeros((len(pairs), 12, 12))
        B[:, :6, :6] = Dblk[pi]; B[:, 6:, 6:] = Dblk[pj]
        for t, (i, j) in enumerate(pairs):
            B[t, :6, 6:] = key[(i, j)]; B[t, 6:, :6] = key[(j, i)]
        Bi = np.linalg.inv(B)
        idx = np.concatenate([pi[:, None] * 6 + np.arange(6)[None, :], pj[:, None] * 6 + np.arange(6)[None, :]], axis=1)   # (np, 12)
        rr = np.repeat(idx[:, :, None], 12, axis=2).ravel(); cc = np.repeat(idx[:, None, :], 12, axis=1).ravel()
        Dinv = Dinv + sp.csr_matrix((Bi.ravel(), (rr, cc)), shape=(6 * n, 6 * n))
    print(f"    pair smoother: {len(pairs)} pairs of {n} nodes")
    return Dinv.tocsr()


class Level:
    pass


def build_hierarchy(A, pos, n, scale_inv, args, active, A_far=None):
    """A: (6n x 6n) csr incl. LM diagonal, in SCALED coordinates; pos: (n,3) positions; scale_inv (n,6): 1/scale
    (the rigid-body modes of the unscaled problem are B_i = [[I, -2[p_i - c]x],[0, I]]; scaled: S^-1 B)."""
    levels = []
    cur_A, cur_pos, cur_n = A, pos, n
    cur_far = A_far
    cur_Sinv = scale_inv
    cur_active = active
    while True:
        L = Level()
        L.A = cur_A
        L.n = cur_n
        L.Dinv = block_diag_inv(cur_A, cur_n) if cur_n > 0 else None
        L.Dsm = L.Dinv
        if args.pair_levels > len(levels) and cur_n > args.dense_below and cur_n > args.coarsest:
            L.Dsm = pair_block_diag_inv(cur_A, cur_n, cur_pos, max(args.theta, 0.3), 0)
        levels.append(L)
        if (cur_n <= args.coarsest and not (args.agg == "cxx" and len(levels) - 1 < len(args.cxx_aggs))) or cur_n <= args.dense_below:
            L.dense = np.linalg.pinv(cur_A.toarray())
            break
        Ab = cur_A.tobsr(blocksize=(6, 6))
        w = np.linalg.norm(Ab.data.reshape(-1, 36), axis=1)
        adj = sp.csr_matrix((w, Ab.indices, Ab.indptr), shape=(cur_n, cur_n))
        adj.setdiag(0)
        adj.eliminate_zeros()
        if args.theta > 0:
            # geometric strength: 1 / |p_i - p_j|^2, keep neighbours within theta of the row maximum
            coo = adj.tocoo()
            d2 = np.sum((cur_pos[coo.row] - cur_pos[coo.col]) ** 2, axis=1) + 1e-12
            wgt = 1.0 / d2
            W = sp.csr_matrix((wgt, (coo.row, coo.col)), shape=adj.shape)
            rmax = W.max(axis=1).toarray().ravel()
            keep = wgt >= args.theta * np.maximum(rmax[coo.row], rmax[coo.col])
            adj = sp.csr_matrix((wgt[keep], (coo.row[keep], coo.col[keep])), shape=adj.shape)
        if args.agg == "cxx":
            agg = args.cxx_aggs[len(levels) - 1].astype(int).copy()
            na = int(agg.max()) + 1
            if len(agg) < cur_n:       # the parking node of the inactive poses (below) rides along
                agg = np.concatenate([agg, -np.ones(cur_n - len(agg), int)])
            if (agg < 0).any():        # inactive nodes: park them in an extra (empty-operator) aggregate
                agg[agg < 0] = na
                na += 1
        elif args.agg == "greedy":
            agg, na = aggregate(adj, cur_n)
        else:
            agg, na = pairwise_aggregate(adj, cur_n, args.passes)
        cnt = np.bincount(agg, minlength=na).astype(float)
        cpos = np.stack([np.bincount(agg, weights=cur_pos[:, k], minlength=na) / cnt for k in range(3)], axis=1)
        # tentative prolongator blocks P_i = Sinv_i * [[I, -2 [p_i - c_I]x], [0, I]]
        d = cur_pos - cpos[agg]
        Pb = np.zeros((cur_n, 6, 6))
        Pb[:, np.arange(6), np.arange(6)] = 1.0
        if args.modes == 6:
            # dp = omega x d with omega = 2 delta  ->  dp = -2 [d]x delta
            Pb[:, 0, 4] = 2 * d[:, 2]; Pb[:, 0, 5] = -2 * d[:, 1]
            Pb[:, 1, 3] = -2 * d[:, 2]; Pb[:, 1, 5] = 2 * d[:, 0]
            Pb[:, 2, 3] = 2 * d[:, 1]; Pb[:, 2, 4] = -2 * d[:, 0]
        Pb = Pb * cur_Sinv[:, :, None]
        Pb = Pb * cur_active[:, None, None]
        if args.geo_smooth > 0:
            # scalar partition-of-unity weights from one damped-Jacobi step on the strong-connection graph Laplacian; every
            # (node, aggregate) weight multiplies the rigid-body transfer block T(p_i - c_J): rigid motions are reproduced exactly
            Aw = adj.copy().tocsr()
            if args.geo_unweighted:
                Aw.data[:] = 1.0
            deg = np.asarray(Aw.sum(axis=1)).ravel()
            deg[deg == 0] = 1.0
            P0 = sp.csr_matrix((np.ones(cur_n), (np.arange(cur_n), agg)), shape=(cur_n, na))
            Wt = ((1 - args.geo_smooth) * P0 + args.geo_smooth * (sp.diags(1.0 / deg) @ (Aw @ P0))).tocoo()
            dd = cur_pos[Wt.row] - cpos[Wt.col]
            nb = len(Wt.row)
            Pb = np.zeros((nb, 6, 6))
            Pb[:, np.arange(6), np.arange(6)] = 1.0
            Pb[:, 0, 4] = 2 * dd[:, 2]; Pb[:, 0, 5] = -2 * dd[:, 1]
            Pb[:, 1, 3] = -2 * dd[:, 2]; Pb[:, 1, 5] = 2 * dd[:, 0]
            Pb[:, 2, 3] = 2 * dd[:, 1]; Pb[:, 2, 4] = -2 * dd[:, 0]
            Pb = Pb * (cur_Sinv[Wt.row][:, :, None]) * (cur_active[Wt.row] * Wt.data)[:, None, None]
            order = np.lexsort((Wt.col, Wt.row))
            indptr = np.concatenate([[0], np.cumsum(np.bincount(Wt.row, minlength=cur_n))])
            P = sp.bsr_matrix((Pb[order], Wt.col[order], indptr), shape=(6 * cur_n, 6 * na)).tocsr()
        else:
            P = sp.bsr_matrix((Pb, agg, np.arange(cur_n + 1)), shape=(6 * cur_n, 6 * na)).tocsr()
        if args.smooth_p > 0 and args.smooth_filter:
            # FILTERED prolongator smoothing: only the geometrically strong (local) blocks of A take part; every dropped
            # block A_ij is lumped into the diagonal through the rigid transfer T(p_j - p_i) (scaled: S_j^-1 T S_i), so
            # the filtered matrix annihilates exactly what A annihilates and the smoothed P still reproduces rigid motions
            Ab2 = cur_A.tobsr(blocksize=(6, 6))
            rows2 = np.repeat(np.arange(cur_n), np.diff(Ab2.indptr)); cols2 = Ab2.indices
            is_diag = rows2 == cols2
            # "far" = much farther than the nearest neighbours of both ends (a looser test than the aggregation's theta:
            # with theta itself most coarse-level neighbours count as weak and the lumped diagonal loses its stiffness)
            d2f = np.sum((cur_pos[rows2] - cur_pos[cols2]) ** 2, axis=1) + 1e-12
            wf = np.where(is_diag, 0.0, 1.0 / d2f)
            rmaxf = sp.csr_matrix((wf, (rows2, cols2)), shape=(cur_n, cur_n)).max(axis=1).toarray().ravel()
            keep = is_diag | (wf >= args.filter_theta * np.minimum(rmaxf[rows2], rmaxf[cols2]))
            weak = ~keep
            Dblk = np.zeros((cur_n, 6, 6)); Dblk[rows2[is_diag]] = Ab2.data[is_diag]
            if weak.any() and not args.filter_nolump:
                dd = cur_pos[cols2[weak]] - cur_pos[rows2[weak]]
                T = np.zeros((int(weak.sum()), 6, 6)); T[:, np.arange(6), np.arange(6)] = 1.0
                T[:, 0, 4] = 2 * dd[:, 2]; T[:, 0, 5] = -2 * dd[:, 1]
                T[:, 1, 3] = -2 * dd[:, 2]; T[:, 1, 5] = 2 * dd[:, 0]
                T[:, 2, 3] = 2 * dd[:, 1]; T[:, 2, 4] = -2 * dd[:, 0]
                Si = np.where(cur_Sinv > 0, 1.0 / np.maximum(cur_Sinv, 1e-300), 0.0)
                Ts = cur_Sinv[cols2[weak]][:, :, None] * T * Si[rows2[weak]][:, None, :]
                np.add.at(Dblk, rows2[weak], np.einsum("bij,bjk->bik", Ab2.data[weak], Ts))
            data = Ab2.data[keep].copy()
            kd = is_diag[keep]
            data[kd] = Dblk[rows2[keep][kd]]
            order = np.lexsort((cols2[keep], rows2[keep]))
            indptr = np.concatenate([[0], np.cumsum(np.bincount(rows2[keep], minlength=cur_n))])
            AF = sp.bsr_matrix((data[order], cols2[keep][order], indptr), shape=cur_A.shape).tocsr()
            DFinv = sp.bsr_matrix((np.linalg.inv(Dblk + 1e-300 * np.eye(6)[None]), np.arange(cur_n), np.arange(cur_n + 1)), shape=cur_A.shape).tocsr()
            P = P - args.smooth_p * (DFinv @ (AF @ P))
        elif args.smooth_p > 0:
            P = P - args.smooth_p * (L.Dinv @ (cur_A @ P))
        L.P = P
        if cur_far is not None and args.local_scale != 1.0:
            far_c = (P.T @ cur_far @ P).tocsr()
            near_c = (P.T @ (cur_A - cur_far) @ P).tocsr()
            cur_A = (near_c / args.local_scale + far_c).tocsr()
            cur_far = far_c
        else:
            cur_A = (P.T @ cur_A @ P).tocsr()
        cur_pos, cur_n = cpos, na
        cur_Sinv = np.ones((na, 6))
        cur_active = np.ones(na)
    return levels


def vcycle(levels, k, r, args):
    L = levels[k]
    if k == len(levels) - 1:
        return L.dense @ r
    om = args.omega
    x = om * (L.Dsm @ r)
    for _ in range(args.nu - 1):
        x = x + om * (L.Dsm @ (r - L.A @ x))
    rc = L.P.T @ (r - L.A @ x)
    ec = vcycle(levels, k + 1, rc, args)
    if args.gamma == 2 and k + 1 < len(levels) - 1 and k + 1 <= args.gamma_depth:
        # W-cycle-ish second visit
        rc2 = rc - levels[k + 1].A @ ec
        ec = ec + vcycle(levels, k + 1, rc2, args)
    x = x + args.over * (L.P @ ec)
    for _ in range(args.nu):
        x = x + om * (L.Dsm @ (r - L.A @ x))
    return x


def pcg(A, b, M, tol, maxit):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    rz0 = rz
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = M(r)
        rz_new = r @ z
        if rz_new <= tol * tol * rz0:
            return x, it
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", default="sphere")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--radius", type=float, nargs="+", default=[1e4, 1e6, 1e8])
    ap.add_argument("--agg", default="greedy")
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--modes", type=int, default=6)
    ap.add_argument("--omega", type=float, default=0.7)
    ap.add_argument("--nu", type=int, default=1)
    ap.add_argument("--over", type=float, default=1.0)
    ap.add_argument("--gamma", type=int, default=1)
    ap.add_argument("--gamma-depth", type=int, default=99, help="levels 1..depth are visited twice (truncated W-cycle)")
    ap.add_argument("--smooth-p", type=float, default=0.0)
    ap.add_argument("--smooth-filter", action="store_true", help="--smooth-p with the filtered (near blocks + rigidly lumped diagonal) matrix")
    ap.add_argument("--filter-nolump", action="store_true", help="--smooth-filter keeping the FULL diagonal blocks (far blocks dropped, not lumped)")
    ap.add_argument("--filter-theta", type=float, default=0.05, help="a block is far (dropped from the smoothing matrix) when 1/d^2 < this x the smaller row maximum")
    ap.add_argument("--geo-smooth", type=float, default=0.0, help="topology-only prolongator smoothing weight")
    ap.add_argument("--geo-unweighted", action="store_true")
    ap.add_argument("--coarsest", type=int, default=8)
    ap.add_argument("--dense-below", type=int, default=0, help="levels of at most this many nodes are solved exactly")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--theta", type=float, default=0.0)
    ap.add_argument("--loops", type=int, default=-1)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--at-truth", action="store_true")
    ap.add_argument("--no-jacobi", action="store_true")
    ap.add_argument("--local-scale", type=float, default=1.0, help="coarse operators: local (non-loop) part divided by this per level")
    ap.add_argument("--pair-levels", type=int, default=0, help="levels 0..k-1 smooth with 12x12 blocks over long-range edge pairs")
    args = ap.parse_args()
    if args.graph == "sphere":
        g = D.sphere()
    elif args.graph == "grid":
        s = args.n or 100
        g = D.manhattan_grid(s, s, max(1, s * s // 20) if args.loops < 0 else args.loops)
    elif args.graph == "torus":
        g = D.torus(args.n or 10000)
    elif args.graph == "kitti":
        g = D.kitti00()
    poses = g.truth if args.at_truth and g.truth is not None else g.poses
    if args.agg == "cxx":
        import posegraph_ceres_b200 as P
        sizes, args.cxx_aggs = P.amg_aggregates(g, args.world)
        args.coarsest = sizes[-1] + 1
        print("cxx hierarchy", sizes)
    t0 = time.time()
    cost, res, grad, jac = O.evaluate(g, poses=poses)
    H = assemble(g, jac)
    N = g.n_poses
    active = (1 - g.pose_const).astype(float)
    diag = H.diagonal().reshape(N, 6)
    scale = active[:, None] / (1.0 + np.sqrt(diag))
    S = sp.diags(scale.ravel())
    Hs = (S @ H @ S).tocsr()
    gs = (scale * grad).ravel()
    print(f"{g.name}: N={N} E={g.n_edges} assemble {time.time() - t0:.1f}s cost {cost:.3f}")
    dsc = np.clip(Hs.diagonal(), 1e-6, 1e32)
    scale_inv = np.where(scale > 0, 1.0 / np.maximum(scale, 1e-300), 0.0)
    for radius in args.radius:
        A = (Hs + sp.diags(dsc / radius)).tocsr()
        # constant pose rows: identity-ish (diag = d only) -- fine
        Dinv = block_diag_inv(A, N)
        t0 = time.time()
        it_j = -1
        if not args.no_jacobi:
            _, it_j = pcg(A, gs, lambda r: Dinv @ r, args.tol, 20000)
        tj = time.time() - t0
        t0 = time.time()
        A_far = None
        if args.local_scale != 1.0:
            far = np.abs(g.edge_ids[:, 0] - g.edge_ids[:, 1]) > 1
            if args.graph == "grid":
                s_ = args.n or 100
                far = np.arange(g.n_edges) >= (s_ * s_ - 1) + (s_ - 1) * s_ - s_   # the random loops come last
                far &= np.abs(g.edge_ids[:, 0] - g.edge_ids[:, 1]) > 2 * s_
            class _G: pass
            gf = _G()
            gf.edge_ids = g.edge_ids[far]; gf.n_edges = int(far.sum()); gf.n_poses = g.n_poses
            Hf = assemble(gf, jac[far])
            A_far = (S @ Hf @ S).tocsr()
            print(f"    far edges: {int(far.sum())}")
        levels = build_hierarchy(A, poses[:, :3], N, scale_inv, args, active, A_far)
        tb = time.time() - t0
        sizes = [L.n for L in levels]
        nnz = [L.A.nnz // 36 for L in levels]
        t0 = time.time()
        x, it_a = pcg(A, gs, lambda r: vcycle(levels, 0, r, args), args.tol, 2000)
        ta = time.time() - t0
        rel = np.linalg.norm(A @ x - gs) / np.linalg.norm(gs)
        print(f"  radius {radius:.0e}: block-Jacobi PCG {it_j} it ({tj:.1f}s) | AMG-PCG {it_a} it ({ta:.1f}s, build {tb:.1f}s) "
              f"levels {sizes} blocks {nnz} opcx {sum(nnz) / nnz[0]:.2f} true rel res {rel:.1e}")


if __name__ == "__main__":
    main()
