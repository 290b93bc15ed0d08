"""Deterministic, array-based synthetic networks for the benchmark configurations
(SURVEY.md 8d).  The reference's own tree generator (data/generate_arterial_tree.py:34-188) is
O(E^2) (segment-intersection rejection with deep copies) and cannot reach 500k edges; this one
applies the same Murray-law rules level by level, vectorised, WITHOUT the rejection step, and
returns plain arrays (node order = BFS creation order = the reference's node/edge order).
"""
import numpy as np

__all__ = ["murray_tree", "tumour_like_graph", "partition_tree_edges"]


def murray_tree(levels, radius0=0.1, gam=0.9, L0=10.0, gdim=3):
    """Full binary arterial tree with `levels` generations (E = 2^levels - 1 edges).

    Rules of data/generate_arterial_tree.py:96-131:  D2 = D0 (gam^3+1)^(-1/3), D1 = gam D2,
    L2 = lmbda D2, L1 = 0.6 lmbda D1 (lmbda = L0/D_root), bifurcation angles from the
    minimum-energy formulas (:111-119), branch side signs alternating -1,+1 per parent
    (`directions = np.tile([-1, 1], ...)`, demo/Tree Profiling.ipynb).
    Returns pos [V,gdim], edges [E,2] (int64), radius [E], level [E].
    """
    assert levels >= 1 and gdim in (2, 3)
    D_root = 2.0 * radius0
    lmbda = L0 / D_root
    pos = [np.array([[0.0, 0.0], [0.0, L0]])]
    edges = [np.array([[0, 1]], dtype=np.int64)]
    radius = [np.array([radius0])]
    level = [np.zeros(1, dtype=np.int64)]
    # state of the previous generation's edges
    end_node = np.array([1], dtype=np.int64)
    end_pos = np.array([[0.0, L0]])
    direc = np.array([[0.0, 1.0]])
    D0 = np.array([D_root])
    next_node = 2
    for g in range(1, levels):
        npar = len(end_node)
        D2 = D0 * (gam ** 3 + 1) ** (-1.0 / 3.0)
        D1 = gam * D2
        L2 = lmbda * D2
        L1 = lmbda * 0.6 * D1
        cos1 = (D0 ** 4 + D1 ** 4 - (D0 ** 3 - D1 ** 3) ** (4.0 / 3.0)) / (2 * D0 ** 2 * D1 ** 2)
        cos2 = (D0 ** 4 + D2 ** 4 - (D0 ** 3 - D2 ** 3) ** (4.0 / 3.0)) / (2 * D0 ** 2 * D2 ** 2)
        a1 = np.arccos(np.clip(cos1, -1, 1))
        a2 = np.arccos(np.clip(cos2, -1, 1))
        k = (2 ** (g - 1) - 1) + np.arange(npar)            # index of the parent in BFS order
        sign1 = np.where(k % 2 == 0, -1.0, 1.0)

        def rotate(d, ang):
            c, s = np.cos(ang), np.sin(ang)
            return np.stack([c * d[:, 0] - s * d[:, 1], s * d[:, 0] + c * d[:, 1]], axis=1)

        d1 = rotate(direc, sign1 * a1)
        d2 = rotate(direc, -sign1 * a2)
        p1 = end_pos + d1 * L1[:, None]
        p2 = end_pos + d2 * L2[:, None]
        new_pos = np.empty((2 * npar, 2))
        new_pos[0::2], new_pos[1::2] = p1, p2
        new_nodes = next_node + np.arange(2 * npar, dtype=np.int64)
        e = np.empty((2 * npar, 2), dtype=np.int64)
        e[:, 0] = np.repeat(end_node, 2)
        e[:, 1] = new_nodes
        r = np.empty(2 * npar)
        r[0::2], r[1::2] = D1 / 2, D2 / 2
        pos.append(new_pos); edges.append(e); radius.append(r)
        level.append(np.full(2 * npar, g, dtype=np.int64))
        nd = np.empty((2 * npar, 2))
        nd[0::2], nd[1::2] = d1, d2
        DD = np.empty(2 * npar)
        DD[0::2], DD[1::2] = D1, D2
        end_node, end_pos, direc, D0 = new_nodes, new_pos, nd, DD
        next_node += 2 * npar
    pos = np.vstack(pos)
    if gdim == 3:
        pos = np.hstack([pos, np.zeros((len(pos), 1))])
    return (np.ascontiguousarray(pos), np.vstack(edges), np.concatenate(radius), np.concatenate(level))


def tumour_like_graph(nx_=22, ny=22, nz=22, drop=0.35, jitter=0.3, seed=0):
    """~10k-edge 3-D random spatial graph for the pulsatile configuration:
    a jittered cubic lattice with random edge deletions, node degree <= 6, largest connected
    component kept, labels compacted.  Returns pos [V,3], edges [E,2], radius [E] (log-uniform
    5-25 micrometres, scaled by 10e-3 like demo/Pulsatile...ipynb:211)."""
    rng = np.random.default_rng(seed)
    idx = np.arange(nx_ * ny * nz).reshape(nx_, ny, nz)
    g = np.stack(np.meshgrid(np.arange(nx_), np.arange(ny), np.arange(nz), indexing="ij"), axis=-1)
    pos = g.reshape(-1, 3).astype(np.float64) + rng.uniform(-jitter, jitter, size=(idx.size, 3))
    e = np.vstack([
        np.stack([idx[:-1].ravel(), idx[1:].ravel()], axis=1),
        np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], axis=1),
        np.stack([idx[:, :, :-1].ravel(), idx[:, :, 1:].ravel()], axis=1)])
    e = e[rng.random(len(e)) > drop]
    # largest connected component (union-find, vectorised label propagation)
    lab = np.arange(idx.size)
    while True:
        m = np.minimum(lab[e[:, 0]], lab[e[:, 1]])
        new = lab.copy()
        np.minimum.at(new, e[:, 0], m)
        np.minimum.at(new, e[:, 1], m)
        if (new == lab).all():
            break
        lab = new
    keep_lab = np.bincount(lab).argmax()
    keep = lab == keep_lab
    e = e[keep[e[:, 0]] & keep[e[:, 1]]]
    remap = -np.ones(idx.size, dtype=np.int64)
    remap[keep] = np.arange(keep.sum())
    e = remap[e]
    flip = rng.random(len(e)) < 0.5
    e[flip] = e[flip][:, ::-1]
    e = e[np.lexsort((e[:, 1], e[:, 0]))]
    radius = np.exp(rng.uniform(np.log(5.0), np.log(25.0), size=len(e))) * 10e-3
    return np.ascontiguousarray(pos[keep] * 50.0 * 10e-3), e, radius


def partition_tree_edges(edges, weights, nparts):
    """Edge-balanced contiguous partition (by edge index) -> part id per edge.  In BFS edge
    order a contiguous index range of a tree level is a set of complete subtrees' slices; used
    only as the simple baseline partitioner (see partition.py for the subtree-aware one)."""
    w = np.asarray(weights, dtype=np.float64)
    c = np.cumsum(w)
    return np.minimum((c - 0.5 * w) / c[-1] * nparts, nparts - 1).astype(np.int64)
