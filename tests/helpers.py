import numpy as np


def hessian_from_jac(g, jac, scale=None):
    """Dense-block reference H = J^T J (dict of 6x6 blocks) and gradient pieces from oracle Jacobians."""
    E = g.n_edges
    Ja = jac[:, 0].reshape(E, 6, 6)
    Jb = jac[:, 1].reshape(E, 6, 6)
    if scale is not None:
        Ja = Ja * scale[g.edge_ids[:, 0]][:, None, :]
        Jb = Jb * scale[g.edge_ids[:, 1]][:, None, :]
    blocks = {}

    def add(i, j, m):
        blocks[(i, j)] = blocks.get((i, j), 0) + m

    for e in range(E):
        a, b = int(g.edge_ids[e, 0]), int(g.edge_ids[e, 1])
        add(a, a, Ja[e].T @ Ja[e])
        add(b, b, Jb[e].T @ Jb[e])
        add(a, b, Ja[e].T @ Jb[e])
        add(b, a, Jb[e].T @ Ja[e])
    return blocks


def bsr_to_dict(rp, ci, vals):
    out = {}
    for i in range(len(rp) - 1):
        for p in range(rp[i], rp[i + 1]):
            out[(i, int(ci[p]))] = vals[p]
    return out


def rot_angle_between(qa, qb):
    """angle (rad) of qa^-1 qb for (n,4) xyzw arrays (need not be exactly unit)"""
    qa = qa / np.linalg.norm(qa, axis=1, keepdims=True)
    qb = qb / np.linalg.norm(qb, axis=1, keepdims=True)
    d = np.abs(np.sum(qa * qb, axis=1)).clip(0, 1)
    return 2 * np.arccos(d)
