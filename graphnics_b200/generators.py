"""Canned network generators, drop-in for graphnics/generators.py:14-151 and
data/generate_arterial_tree.py:34-188.  Pure networkx/numpy host code (graph definition only);
each returns a FenicsGraph with `make_mesh()` already called, like the reference.
"""
import random

import networkx as nx
import numpy as np

from .fenics_graph import FenicsGraph, copy_from_nx_graph

__all__ = ["line_graph", "honeycomb", "Y_bifurcation", "YY_bifurcation", "make_Y_bifurcation",
           "make_arterial_tree"]


def _finish(G, mesh=True, n=1):
    if mesh:
        G.make_mesh(n)
    return G


def line_graph(n, dim=2, dx=1, mesh=True):
    """n nodes on the x axis, spacing dx (generators.py:14-34)."""
    G = FenicsGraph()
    G.add_nodes_from(range(0, n))
    for i in range(0, n):
        G.nodes[i]["pos"] = [i * dx] + [0] * (dim - 1)
    for i in range(0, n - 1):
        G.add_edge(i, i + 1)
    return _finish(G, mesh)


def honeycomb(n, m, mesh=True):
    """Hexagonal lattice with an inlet below node 0 and an outlet above the node furthest from
    the origin (generators.py:37-85)."""
    H = nx.convert_node_labels_to_integers(nx.hexagonal_lattice_graph(n, m))
    inlet = len(H.nodes)
    H.add_node(inlet)
    H.nodes[inlet]["pos"] = [0, -1]
    H.add_edge(inlet, 0)
    coords = np.asarray(list(nx.get_node_attributes(H, "pos").values()))
    far = int(np.argmax(np.linalg.norm(coords, axis=1), axis=0))
    outlet = len(H.nodes)
    H.add_node(outlet)
    H.nodes[outlet]["pos"] = coords[far, :] + np.asarray([0.7, 1])
    H.add_edge(outlet, far)
    G = copy_from_nx_graph(H)
    if (0, inlet) in G.edges():          # inlet edge must point inwards
        G.remove_edge(0, inlet)
        G.add_edge(inlet, 0)
    return _finish(G, mesh)


def Y_bifurcation(dim=2, mesh=True):
    """Y-shaped bifurcation (generators.py:88-111)."""
    G = FenicsGraph()
    G.add_nodes_from([0, 1, 2, 3])
    pad = [0] * (dim - 2)
    for i, p in enumerate([[0, 0], [0, 0.5], [-0.5, 1], [0.5, 1]]):
        G.nodes[i]["pos"] = p + pad
    G.add_edges_from([(0, 1), (1, 2), (1, 3)])
    return _finish(G, mesh)


make_Y_bifurcation = Y_bifurcation      # name used by BASELINE.json


def YY_bifurcation(dim=2, mesh=True):
    """Two generations of Y bifurcations (generators.py:114-151)."""
    G = FenicsGraph()
    G.add_nodes_from(range(8))
    pad = [0] * (dim - 2)
    pts = [[0, 0], [0, 0.5], [-0.5, 1], [0.5, 1], [-0.75, 1.5], [-0.25, 1.5], [0.25, 1.5], [0.75, 1.5]]
    for i, p in enumerate(pts):
        G.nodes[i]["pos"] = p + pad
    G.add_edges_from([(0, 1), (1, 2), (1, 3), (2, 4), (2, 5), (3, 6), (3, 7)])
    return _finish(G, mesh)


# ------------------------------------------------------------------------------------------
# Murray-law arterial tree with segment-intersection rejection
# ------------------------------------------------------------------------------------------
def _orient(p, q, r):
    val = (q[..., 1] - p[..., 1]) * (r[..., 0] - q[..., 0]) - (q[..., 0] - p[..., 0]) * (r[..., 1] - q[..., 1])
    return np.where(val > 0, 1, np.where(val < 0, 2, 0))


def _on_segment(p, q, r):
    return ((q[..., 0] <= np.maximum(p[..., 0], r[..., 0])) & (q[..., 0] >= np.minimum(p[..., 0], r[..., 0])) &
            (q[..., 1] <= np.maximum(p[..., 1], r[..., 1])) & (q[..., 1] >= np.minimum(p[..., 1], r[..., 1])))


def _segments_intersect(A, B, C, D):
    """vectorised over rows of A,B (existing edges) against one candidate C->D; same predicate as
    data/generate_arterial_tree.py:221-253 (general position + the four collinear cases)."""
    C = np.broadcast_to(C, A.shape)
    D = np.broadcast_to(D, A.shape)
    o1, o2, o3, o4 = _orient(A, B, C), _orient(A, B, D), _orient(C, D, A), _orient(C, D, B)
    hit = (o1 != o2) & (o3 != o4)
    hit |= (o1 == 0) & _on_segment(A, C, B)
    hit |= (o2 == 0) & _on_segment(A, D, B)
    hit |= (o3 == 0) & _on_segment(C, A, D)
    hit |= (o4 == 0) & _on_segment(C, B, D)
    return hit


def make_arterial_tree(N, radius0=1, gam=0.8, L0=10, directions=False, uniform_lengths=False, mesh=True):
    """N generations of a planar Murray-law tree (data/generate_arterial_tree.py:34-188):
    D2 = D0 (gam^3+1)^(-1/3), D1 = gam D2, L2 = lmbda D2, L1 = 0.6 lmbda D1, energy-optimal
    bifurcation angles, side signs from `directions` (consumed, like the reference) or random;
    a daughter is dropped if it crosses any non-neighbouring vessel."""
    if gam > 1:
        raise Exception("Please choose a value for gamma lower or equal to 1")
    D_root = 2 * radius0
    lmbda = L0 / D_root
    pos = [np.array([0.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0]) * L0]
    edges = [(0, 1)]
    radius = {(0, 1): D_root / 2}
    previous = [(0, 1)]
    for igen in range(1, N):
        current = []
        for e in previous:
            pa, pb = pos[e[0]], pos[e[1]]
            D0 = radius[e] * 2
            D2 = D0 * (gam ** 3 + 1) ** (-1 / 3)
            D1 = gam * D2
            L2, L1 = lmbda * D2, lmbda * 0.6 * D1
            if uniform_lengths:
                L1, L2 = L0, L0
            cos1 = (D0 ** 4 + D1 ** 4 - (D0 ** 3 - D1 ** 3) ** (4 / 3)) / (2 * D0 ** 2 * D1 ** 2)
            cos2 = (D0 ** 4 + D2 ** 4 - (D0 ** 3 - D2 ** 3) ** (4 / 3)) / (2 * D0 ** 2 * D2 ** 2)
            ang1, ang2 = np.degrees(np.arccos(cos1)), np.degrees(np.arccos(cos2))
            if not directions:
                sign1 = random.choice((-1, 1))
            else:
                sign1 = directions[0]
                del directions[0]
            for sign, ang, L, D in ((sign1, ang1, L1, D1), (-sign1, ang2, L2, D2)):
                d = (pb - pa)[:2]
                th = np.radians(sign * ang)
                nd = np.array([np.cos(th) * d[0] - np.sin(th) * d[1], np.sin(th) * d[0] + np.cos(th) * d[1]])
                nd = nd / np.linalg.norm(nd)
                new = np.array([pb[0] + nd[0] * L, pb[1] + nd[1] * L, pb[2]])
                P = np.asarray(pos)
                ea = np.asarray(edges)
                non_nb = (ea[:, 0] != e[1]) & (ea[:, 1] != e[1])
                A, B = P[ea[non_nb, 0]][:, :2], P[ea[non_nb, 1]][:, :2]
                if not _segments_intersect(A, B, pb[:2], new[:2]).any():
                    k = len(pos)
                    pos.append(new)
                    edges.append((e[1], k))
                    radius[(e[1], k)] = D / 2
                    current.append((e[1], k))
        previous = current
    H = nx.DiGraph()
    H.add_nodes_from(range(len(pos)))
    for i, p in enumerate(pos):
        H.nodes[i]["pos"] = p
    # DiGraph.edges() order of the reference: by source node, then insertion
    for u, v in edges:
        H.add_edge(u, v, radius=radius[(u, v)])
    G = copy_from_nx_graph(H)
    return _finish(G, mesh, 3)
