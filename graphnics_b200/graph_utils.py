"""Graph pre-processing helpers, drop-in for graphnics/graph_utils.py:19-271:
branch colouring, Murray's-law radii, distance from a source node.  Host-side networkx/numpy code --
none of it is on the assemble+solve path."""
import networkx as nx
import numpy as np

from .coefficients import UserExpression
from .fenics_graph import copy_from_nx_graph

__all__ = ["color_graph", "plot_graph_color", "assign_radius_using_Murrays_law", "DistFromSource"]


class _Copy:
    """copy of bifurcation node `b` that keeps only the edge towards `other`"""
    __slots__ = ("b", "other")

    def __init__(self, b, other):
        self.b, self.other = b, other


def color_graph(G):
    """Colour the branches of a graph: edge attribute "color" (int), graph_utils.py:19-109.

    Same construction as the reference so that the colours agree edge by edge: every bifurcation point
    (undirected degree > 2) is pulled apart into one copy per incident edge, the connected pieces of the
    resulting graph that are simple chains are the branches (numbered in networkx's component order), and
    whatever stays uncoloured -- edges between two bifurcation points whose piece is not a chain -- gets
    a colour of its own at the end."""
    U = nx.Graph(G)
    D = nx.Graph(G.copy(as_view=False))
    bifs = [n for n in U.nodes() if U.degree(n) > 2]
    is_bif = set(bifs)
    for b in bifs:
        for e in U.edges(b):
            other = e[1] if e[0] == b else e[0]
            c = _Copy(b, other)
            D.add_node(c)
            D.add_edge(c, other)
            if D.has_edge(b, other):
                D.remove_edge(b, other)
            else:
                assert other in is_bif, "Graph coloring algorithm may be malfunctioning"
    orig = lambda x: x.b if isinstance(x, _Copy) else x
    n_branches = 0
    for comp in nx.connected_components(D):
        sub = D.subgraph(comp)
        degs = [d for _, d in sub.degree()]
        if min(degs) > 0 and max(degs) < 3:
            for x, y in sub.edges():
                u, v = orig(x), orig(y)
                G.edges()[(u, v) if G.has_edge(u, v) else (v, u)]["color"] = n_branches
            n_branches += 1
    for e in G.edges():
        if "color" not in G.edges()[e]:
            G.edges()[e]["color"] = n_branches
            n_branches += 1


def plot_graph_color(G):
    """Plots the colouring stored in the edge attribute 'color' (graph_utils.py:112-123); needs matplotlib."""
    pos = {n: np.asarray(p)[:2] for n, p in nx.get_node_attributes(G, "pos").items()}
    colors = nx.get_edge_attributes(G, "color")
    nx.draw_networkx_edge_labels(G, pos, colors)
    nx.draw_networkx(G, pos)


def assign_radius_using_Murrays_law(G, start_node, start_radius):
    """Radii from Murray's law r_p^3 = sum_d r_d^3, daughters weighted by the length of the subnetwork
    sprouting from them (graph_utils.py:126-202).  Returns the breadth-first tree of G from start_node
    (a networkx DiGraph, edges in BFS order) with "pos", "length" and "radius" attributes.

    The reference recomputes every subtree length with shortest-path searches (O(V^2)); here they are
    accumulated bottom-up once."""
    T = nx.bfs_tree(G, source=start_node)
    for n in T.nodes():
        T.nodes()[n]["pos"] = G.nodes()[n]["pos"]
    for u, v in T.edges():
        T.edges()[(u, v)]["length"] = float(np.linalg.norm(np.asarray(T.nodes()[v]["pos"], dtype=float)
                                                           - np.asarray(T.nodes()[u]["pos"], dtype=float)))
    assert len(list(G.edges(start_node))) == 1, "start node has to have a single edge sprouting from it"
    below = {n: 0.0 for n in T.nodes()}                        # total edge length of the subtree hanging from n
    for n in reversed(list(nx.topological_sort(T))):
        for c in T.successors(n):
            below[n] += below[c] + T.edges()[(n, c)]["length"]
    for i, (u, v) in enumerate(T.edges()):
        if i == 0:
            T.edges()[(u, v)]["radius"] = start_radius
            continue
        (pu, _), = T.in_edges(u)
        fraction = (below[v] + T.edges()[(u, v)]["length"]) / below[u]
        T.edges()[(u, v)]["radius"] = fraction ** (1 / 3) * T.edges()[(pu, u)]["radius"]
    return T


class DistFromSource(UserExpression):
    """Distance along the graph from `source_node` (graph_utils.py:205-271): nodal distances over the
    breadth-first tree, linear along every edge.  `dist(x)` / `eval(values, x)` locate the edge nearest to x."""

    def __init__(self, G, source_node, **kwargs):
        self.G, self.source = G, source_node
        super().__init__(**kwargs)
        T = nx.bfs_tree(G, source_node)
        pos = {n: np.asarray(G.nodes()[n]["pos"], dtype=float) for n in G.nodes()}
        for u, v in T.edges():
            T.edges()[(u, v)]["length"] = float(np.linalg.norm(pos[v] - pos[u]))
        dist = nx.shortest_path_length(T, source_node, weight="length")
        if len(nx.get_edge_attributes(G, "length")) == 0 and hasattr(G, "compute_edge_lengths") and G._dn is not None:
            G.compute_edge_lengths()
        nodes = list(G.nodes())
        self._P = np.asarray([pos[n] for n in nodes])
        self._d = np.asarray([dist.get(n, np.nan) for n in nodes])
        ix = {n: i for i, n in enumerate(nodes)}
        self._e = np.asarray([[ix[u], ix[v]] for u, v in G.edges()], dtype=np.int64).reshape(-1, 2)

    def _value(self, x):
        x = np.asarray(x, dtype=float).ravel()
        gd = self._P.shape[1]
        x = np.concatenate([x, np.zeros(max(0, gd - x.size))])[:gd]
        a, b = self._P[self._e[:, 0]], self._P[self._e[:, 1]]
        ab = b - a
        s = np.clip(((x - a) * ab).sum(1) / (ab * ab).sum(1), 0.0, 1.0)
        k = int(np.argmin(np.linalg.norm(a + s[:, None] * ab - x, axis=1)))
        da, db = self._d[self._e[k, 0]], self._d[self._e[k, 1]]
        return float((1 - s[k]) * da + s[k] * db)

    def eval(self, values, x):
        values[0] = self._value(x)

    def __call__(self, *x):
        return self._value(x[0] if len(x) == 1 and np.ndim(x[0]) > 0 else x)

    def value_shape(self):
        return ()
