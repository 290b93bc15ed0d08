"""Oracle, graph level: arrays from a networkx DiGraph, node tables, tangents.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy only.
"""
import numpy as np

__all__ = [
    "BIF_IN", "BIF_OUT", "BOUN_IN", "BOUN_OUT",
    "graph_arrays", "node_tables", "edge_tangents", "edge_lengths", "vertex_degrees",
]

# reference: graphnics/fenics_graph.py:22-25
BIF_IN = 1
BIF_OUT = 2
BOUN_IN = 3
BOUN_OUT = 4


def graph_arrays(G):
    """(pos[V,gdim] f64, edges[E,2] i64) in the iteration orders the reference uses.

    reference: fenics_graph.py:71 (coords by node ITERATION order) and :72 (cells hold
    node LABELS, in `G.edges()` iteration order).  Labels must be 0..V-1 in insertion order
    (all reference generators guarantee it); we assert it instead of silently mis-indexing.
    """
    nodes = list(G.nodes())
    assert nodes == list(range(len(nodes))), "node labels must be 0..V-1 in insertion order"
    pos = np.asarray([G.nodes[v]["pos"] for v in nodes], dtype=np.float64)
    edges = np.asarray([[u, v] for u, v in G.edges()], dtype=np.int64).reshape(-1, 2)
    return pos, edges


def node_tables(num_nodes, edges):
    """Bifurcation / boundary node lists.  reference: fenics_graph.py:161-184.

    degree = #in-edges + #out-edges; ==1 -> boundary, >1 -> bifurcation (node order).
    Returns (bifurcation_ixs, boundary_ixs, degree).
    """
    deg = np.zeros(num_nodes, dtype=np.int64)
    for u, v in np.asarray(edges).reshape(-1, 2):
        deg[u] += 1
        deg[v] += 1
    bif = [int(v) for v in range(num_nodes) if deg[v] > 1]
    bnd = [int(v) for v in range(num_nodes) if deg[v] == 1]
    return bif, bnd, deg


def edge_tangents(pos, edges):
    """Unit tangents, bit-for-bit the reference arithmetic.  fenics_graph.py:239-246:
    t = pos[v]-pos[u]; norm = np.linalg.norm(t); inv = 1.0/norm; t *= inv.
    """
    out = np.empty((len(edges), pos.shape[1]), dtype=np.float64)
    for i, (u, v) in enumerate(edges):
        t = np.asarray(pos[v]) - np.asarray(pos[u], dtype=np.float64)
        nrm = np.linalg.norm(t)
        inv = 1.0 / nrm
        t *= inv
        out[i] = t
    return out


def edge_lengths(pos, edges):
    """reference: fenics_graph.py:187-198."""
    return np.asarray([np.linalg.norm(pos[v] - pos[u]) for u, v in edges], dtype=np.float64)


def vertex_degrees(num_nodes, pos, edges):
    """Weighted vertex degree = sum of incident edge lengths / 2 (fenics_graph.py:200-226).

    The reference adds in-edges first, then out-edges, each in adjacency order.
    """
    L = edge_lengths(pos, edges)
    out = np.zeros(num_nodes)
    for v in range(num_nodes):
        l_v = 0
        for i, (a, b) in enumerate(edges):
            if b == v:
                l_v += L[i]
        for i, (a, b) in enumerate(edges):
            if a == v:
                l_v += L[i]
        out[v] = l_v / 2
    return out
