"""Edge partition of a network for the multi-GPU solve ("the graph partitions naturally by edges,
so large networks are split into edge-balanced subtrees").

A rank owns a set of edges and, exclusively, their q and p unknowns.  What makes a partition GOOD for the
solver is the rake forest of the graph-level Schur system (schur_plan.SchurPlan): a raked node whose whole
subtree -- all edges at it and at its rake descendants -- sits on one rank is eliminated by that rank alone;
every other bifurcation is REPLICATED (processed redundantly by all ranks after the one exchange of a solve).
So the partitioner cuts the rake forest into subtrees: starting from the forest roots it keeps splitting the
heaviest piece (its root joins the replicated set, its children become pieces) until a longest-processing-time
packing of the pieces into `world` bins is balanced; the edges that belong to replicated nodes, and all edges of
a cyclic core, are dealt out to fill the bins.  A tree on 2^k ranks is cut k levels below its root (C5 on eight
GPUs: 8 subtrees, 7 replicated bifurcations); a honeycomb has no rakeable node and degenerates to contiguous
index ranges with every bifurcation replicated.

Host side, numpy, once per network.
"""
import heapq

import numpy as np

from .schur_plan import SchurPlan, node_owners

__all__ = ["bifurcation_index", "partition_edges", "contiguous_partition"]


def bifurcation_index(V, edges):
    """(bix [V]: index among the nodes of degree >= 2 in node order, -1 otherwise; B) -- fenics_graph.py:161-184"""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    deg = np.bincount(edges.ravel(), minlength=V)
    isb = deg > 1
    return np.where(isb, np.cumsum(isb) - 1, -1), int(isb.sum())


def contiguous_partition(E, world):
    """equal contiguous index ranges (chunk = ceil(E / world)) -- the layout of round 1, still what a network
    without tree-like parts gets"""
    chunk = -(-E // world)
    return np.minimum(np.arange(E, dtype=np.int64) // chunk, world - 1).astype(np.int32)


def partition_edges(edges, V, world, weights=None, plan=None, tol=0.03, max_replicated=1024):
    """edge -> rank (int32 [E]) and the plan it was derived from.  `weights`: work per edge (default 1: every
    edge has the same number of cells)."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    E = len(edges)
    w_e = np.ones(E) if weights is None else np.asarray(weights, dtype=np.float64)
    if world <= 1:
        return np.zeros(E, dtype=np.int32), plan
    bix, B = bifurcation_index(V, edges)
    if plan is None:
        plan = SchurPlan(edges, bix, B)
    if B == 0 or not plan.levels:
        return contiguous_partition(E, world), plan
    bu, bv = bix[edges[:, 0]], bix[edges[:, 1]]
    parent, pedge = plan.parent.astype(np.int64), plan.pedge.astype(np.int64)
    raked = np.zeros(B, dtype=bool)
    for nodes in plan.levels:
        raked[nodes] = True
    # the raked node an edge hangs on: the child end of a rake link, the bifurcation end of a boundary edge
    own = np.full(E, -1, dtype=np.int64)
    link_child = np.nonzero(pedge >= 0)[0]
    own[pedge[link_child]] = link_child
    bnd_u, bnd_v = (bu < 0) & (bv >= 0), (bv < 0) & (bu >= 0)
    own[bnd_u] = bv[bnd_u]
    own[bnd_v] = bu[bnd_v]
    own[(own >= 0) & ~raked[np.maximum(own, 0)]] = -1          # edges hanging on core nodes are dealt out
    wn = np.bincount(own[own >= 0], weights=w_e[own >= 0], minlength=B)
    sub = wn.copy()                                            # weight of the whole rake subtree
    for nodes in plan.levels:
        par = parent[nodes]
        has = (par >= 0) & raked[np.maximum(par, 0)]
        np.add.at(sub, par[has], sub[nodes[has]])
    core = ~raked
    par_core = np.where(parent >= 0, core[np.maximum(parent, 0)], True)
    pieces = [(-sub[i], int(i)) for i in np.nonzero(raked & par_core)[0]]   # forest roots + children of the core
    heapq.heapify(pieces)
    target = w_e.sum() / world
    ch_ptr, ch_idx = plan.ch_ptr, plan.ch_idx
    loose = float(w_e[own < 0].sum())                          # weight that can be dealt out freely
    n_repl = 0

    def pack(pcs):
        load = np.zeros(world)
        where = {}
        for negw, i in sorted(pcs):
            r = int(np.argmin(load))
            load[r] += -negw
            where[i] = r
        return load, where

    while True:
        load, where = pack(pieces)
        # the loose weight fills the bins from below
        lvl = np.sort(load)
        room = float(np.maximum(lvl[-1] - lvl, 0.0).sum())
        peak = lvl[-1] if loose <= room else lvl[-1] + (loose - room) / world
        if peak <= target * (1.0 + tol) or n_repl >= max_replicated:
            break
        negw, i = pieces[0]
        kids = ch_idx[ch_ptr[i]:ch_ptr[i + 1]]
        kids = kids[raked[kids]]
        if len(kids) == 0:
            break
        heapq.heappop(pieces)
        for c in kids:
            heapq.heappush(pieces, (-sub[c], int(c)))
        loose += wn[i]
        own[own == i] = -1
        n_repl += 1
    # ---- edges of the pieces: a node inherits the piece of its parent (top-down)
    piece_rank = np.full(B, -1, dtype=np.int64)
    for i, r in where.items():
        piece_rank[i] = r
    root_of_piece = np.zeros(B, dtype=bool)
    root_of_piece[list(where)] = True
    for nodes in reversed(plan.levels):
        inner = nodes[~root_of_piece[nodes]]
        par = parent[inner]
        has = par >= 0
        piece_rank[inner[has]] = np.where(piece_rank[par[has]] >= 0, piece_rank[par[has]], -1)
    epart = np.full(E, -1, dtype=np.int64)
    owned = own >= 0
    epart[owned] = piece_rank[own[owned]]
    # ---- the rest in index order, each rank taking what it lacks
    free = np.nonzero(epart < 0)[0]
    if len(free):
        load = np.bincount(epart[epart >= 0], weights=w_e[epart >= 0], minlength=world)
        fw = np.cumsum(w_e[free])
        total = load.sum() + fw[-1]
        want = np.maximum(total / world - load, 0.0)
        want *= fw[-1] / max(want.sum(), 1e-300)
        cuts = np.cumsum(want)
        epart[free] = np.minimum(np.searchsorted(cuts, fw - 0.5 * w_e[free], side="left"), world - 1)
    return epart.astype(np.int32), plan
