"""Oracle, elimination identity (SURVEY.md 8a "Elimination identity"): an independent numpy route
from the mixed right-hand side to the solution that never forms or factorises the monolithic
matrix -- per-edge prefix sums, a graph-level SPD solve on the bifurcation multipliers, per-edge
back-substitution.  Derived from the forms of MixedHydraulicNetwork
(graphnics/flowmodels/flow_models.py:233-236,251-262), not from the CUDA code; it works edge by
edge in physical order found by SORTING the submesh vertices along the edge.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used to check the LU route against a second
derivation and to play the ranks of the edge-partitioned algorithm on the CPU (gloo tests).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem

__all__ = ["edge_order", "eliminate_edges", "schur_system", "backsub_edges", "solve_by_elimination"]


def edge_order(om):
    """(node_ids [E, m+1], cell_ids [E, m], h [E, m]): submesh-local vertex ids / parent-cell slots of
    every edge in physical order from u to v."""
    E, m = om.E, om.m
    u = om.pos[om.edges[:, 0]]
    dist = np.linalg.norm(om.sub_coords - u[:, None, :], axis=2)        # [E, m+1]
    node_ids = np.argsort(dist, axis=1, kind="stable")
    rank = np.empty_like(node_ids)
    ar = np.arange(E)[:, None]
    rank[ar, node_ids] = np.arange(m + 1)[None, :]
    cpos = np.minimum(rank[ar, om.sub_cells[:, :, 0]], rank[ar, om.sub_cells[:, :, 1]])   # [E, m] position of cell
    cell_ids = np.argsort(cpos, axis=1, kind="stable")
    xs = om.sub_coords[ar, node_ids]
    h = np.linalg.norm(xs[:, 1:] - xs[:, :-1], axis=2)
    return node_ids, cell_ids, h


def eliminate_edges(om, b, a_edge, sigma=1.0, edges=None, order=None):
    """edge scalars (cond, rsrc, ftot) of the edges `edges` (default all) for the operator
    [a M, -s B^T, s C^T; s B, 0, 0; s C, 0, 0]"""
    lay = fem.mixed_layout(om)
    node_ids, cell_ids, h = order if order is not None else edge_order(om)
    es = np.arange(om.E) if edges is None else np.asarray(edges)
    m = om.m
    bq = b[lay["q_off"][es][:, None] + node_ids[es]]                     # [k, m+1] physical order
    F = b[lay["p_off"] + om.parent_cell[es[:, None], cell_ids[es]]] / sigma
    hh = h[es]
    a = np.asarray(a_edge, dtype=np.float64)[es]
    cum = np.concatenate([np.zeros((len(es), 1)), np.cumsum(F, axis=1)], axis=1)      # cumF at nodes
    s2 = (hh * (cum[:, :-1] + cum[:, 1:]) * 0.5).sum(axis=1)
    L = hh.sum(axis=1)
    return 1.0 / (a * L), bq.sum(axis=1) - a * s2, cum[:, -1]


def schur_system(om, cond, rsrc, ftot, b_lam=None, sigma=1.0):
    """S P = r on the bifurcation multipliers P = sigma*lambda (Kirchhoff at every bifurcation)"""
    B = om.B
    bix = -np.ones(om.V, dtype=np.int64)
    bix[om.bif_ixs] = np.arange(B)
    bu, bv = bix[om.edges[:, 0]], bix[om.edges[:, 1]]
    diag = np.zeros(B)
    rhs = np.zeros(B) if b_lam is None else -np.asarray(b_lam) / sigma
    np.add.at(diag, bu[bu >= 0], cond[bu >= 0])
    np.add.at(diag, bv[bv >= 0], cond[bv >= 0])
    np.add.at(rhs, bv[bv >= 0], (cond * rsrc + ftot)[bv >= 0])           # edge INTO the node
    np.add.at(rhs, bu[bu >= 0], (-cond * rsrc)[bu >= 0])                 # edge OUT of the node
    inner = (bu >= 0) & (bv >= 0)
    S = sp.coo_matrix((np.concatenate([-cond[inner], -cond[inner]]),
                       (np.concatenate([bu[inner], bv[inner]]), np.concatenate([bv[inner], bu[inner]]))),
                      shape=(B, B)).tocsr() + sp.diags(diag)
    return S, rhs, bix


def backsub_edges(om, b, a_edge, cond, rsrc, P, bix, sigma=1.0, edges=None, order=None, x=None):
    """q and p of the edges `edges` from the multipliers; writes into x (monolithic layout)"""
    lay = fem.mixed_layout(om)
    node_ids, cell_ids, h = order if order is not None else edge_order(om)
    es = np.arange(om.E) if edges is None else np.asarray(edges)
    x = np.zeros(lay["N"]) if x is None else x
    a = np.asarray(a_edge, dtype=np.float64)[es][:, None]
    qd = lay["q_off"][es][:, None] + node_ids[es]
    pd = lay["p_off"] + om.parent_cell[es[:, None], cell_ids[es]]
    G = b[qd]
    F = b[pd] / sigma
    hh = h[es]
    bu, bv = bix[om.edges[es, 0]], bix[om.edges[es, 1]]
    Pu = np.where(bu >= 0, P[np.maximum(bu, 0)], 0.0)
    Pv = np.where(bv >= 0, P[np.maximum(bv, 0)], 0.0)
    q0 = cond[es] * (Pu - Pv + rsrc[es])
    q = q0[:, None] + np.concatenate([np.zeros((len(es), 1)), np.cumsum(F, axis=1)], axis=1)
    # (a M q)_j = a [ h_{j-1} (q_{j-1} + 2 q_j) + h_j (2 q_j + q_{j+1}) ] / 6
    mq = np.zeros_like(q)
    mq[:, :-1] += a * hh * (2 * q[:, :-1] + q[:, 1:]) / 6.0
    mq[:, 1:] += a * hh * (q[:, :-1] + 2 * q[:, 1:]) / 6.0
    D = G - mq                                                           # P_j - P_{j-1}, P_{-1} = P(u)
    Pc = Pu[:, None] + np.cumsum(D[:, :-1], axis=1)                      # cells 0..m-1
    x[qd] = q
    x[pd] = Pc / sigma
    return x


def solve_by_elimination(om, b, Res=1.0, sigma=1.0):
    a = np.broadcast_to(np.asarray(Res, dtype=np.float64), (om.E,))
    lay = fem.mixed_layout(om)
    order = edge_order(om)
    cond, rsrc, ftot = eliminate_edges(om, b, a, sigma, order=order)
    x = np.zeros(lay["N"])
    if om.B:
        S, rhs, bix = schur_system(om, cond, rsrc, ftot, b[lay["lam_off"]:], sigma)
        P = spla.spsolve(S.tocsc(), rhs)
        x[lay["lam_off"]:] = P / sigma
    else:
        P, bix = np.zeros(1), -np.ones(om.V, dtype=np.int64)
    return backsub_edges(om, b, a, cond, rsrc, P, bix, sigma, order=order, x=x)
