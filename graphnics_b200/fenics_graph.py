"""FenicsGraph: drop-in for graphnics/fenics_graph.py:28-311 -- a networkx.DiGraph whose
`make_mesh` / `make_submeshes` / tangent computations run as CUDA kernels (graphnics_b200.device)
instead of DOLFIN `MeshEditor`/`refine`/`adapt`, xii `EmbeddedMesh` and Python callbacks.

Same attributes after `make_mesh(n)`: geom_dim, num_edges, mesh, mf, bifurcation_ixs,
num_bifurcations, boundary_ixs, tangents, tangent, edges[e]['tangent']; after
`make_submeshes()`: edges[e]['submesh'], edges[e]['vf'].
"""
from functools import partial
from operator import itemgetter

import networkx as nx
import numpy as np

from .coefficients import (UserExpression, UFLExpr, Constant, Expression, dot, grad, tangent_components,
                           as_coefficient)
from .device import DeviceNetwork
from .mesh import Mesh1D, SubMesh, CellMarker, VertexMarker

__all__ = ["BIF_IN", "BIF_OUT", "BOUN_IN", "BOUN_OUT", "FenicsGraph", "GlobalFlux",
           "GlobalCrossectionFlux", "TangentFunction", "copy_from_nx_graph", "nxgraph_attribute_to_dolfin"]

class VaryingEdgeAttribute(Exception):
    """an edge attribute that is a function of position (and parameters) rather than one number per edge; carries the
    per-edge objects, in device edge order, for whoever can sample them"""

    def __init__(self, key, per_edge):
        super().__init__(f"edge attribute '{key}' varies along the edge")
        self.key, self.per_edge = key, per_edge


# Marker tags (graphnics/fenics_graph.py:22-25)
BIF_IN = 1
BIF_OUT = 2
BOUN_IN = 3
BOUN_OUT = 4


class _TangentField:
    """`G.tangent`: DG0 vector field (tangent of the edge a cell belongs to) -- kept implicit."""

    def __init__(self, G):
        self.G = G

    def vector(self):
        from .function import Vector
        dn = self.G._dn
        t = dn.tangents.repeat_interleave(dn.m, dim=0)
        return Vector(t.reshape(-1))

    def __call__(self, *x):
        from .function import Function, FunctionSpace
        f = Function(FunctionSpace(self.G.mesh, "DG", 0), net=self.G._dn, role="cell")
        out = []
        for k in range(self.G.geom_dim):
            f.vector().tensor.copy_(self.G._dn.tangents[:, k].repeat_interleave(self.G._dn.m))
            f.vector().touch()
            out.append(f(*x))
        return np.asarray(out)


class _StampedAttrs(dict):
    """Edge attribute dictionary that counts its mutations in a stamp shared by all edges of one graph (`_stamp`: the
    graph's one-element list, handed over by the factory FenicsGraph.__init__ installs): edge_attribute_array() walks
    the E dictionaries again only when some attribute was written since its last walk.  The stamp is held per instance
    (not per class), so copy.deepcopy / pickle of a graph give the copy its own stamp, shared by its own dictionaries."""
    __slots__ = ("_stamp",)

    def __init__(self, stamp=None):
        dict.__init__(self)
        self._stamp = stamp

    def _bump(self):
        st = getattr(self, "_stamp", None)
        if st is not None:
            st[0] += 1

    def __setitem__(self, k, v):
        self._bump()
        dict.__setitem__(self, k, v)

    def __delitem__(self, k):
        self._bump()
        dict.__delitem__(self, k)

    def __ior__(self, other):
        self._bump()
        dict.update(self, other)
        return self

    def update(self, *a, **kw):
        self._bump()
        dict.update(self, *a, **kw)

    def setdefault(self, k, default=None):
        self._bump()
        return dict.setdefault(self, k, default)

    def pop(self, *a):
        self._bump()
        return dict.pop(self, *a)

    def popitem(self):
        self._bump()
        return dict.popitem(self)

    def clear(self):
        self._bump()
        dict.clear(self)


class FenicsGraph(nx.DiGraph):
    """Class for constructing meshes and associated functions from networkx directed graphs.
    Node attribute "pos" (2 or 3 coordinates) is required; node labels must be 0..V-1 in
    insertion order (true for every reference generator; the reference silently mis-indexes
    otherwise, fenics_graph.py:71-72 -- we raise)."""

    def __init__(self):
        stamp = [0]
        # (networkx calls self.edge_attr_dict_factory() for every new edge: an instance attribute overrides the class's)
        self.edge_attr_dict_factory = partial(_StampedAttrs, stamp)
        nx.DiGraph.__init__(self)
        self._attr_stamp = stamp
        self._dn = None
        self._systems = {}

    # ---- arrays the kernels consume ---------------------------------------------------------
    def _graph_arrays(self):
        nodes = list(self.nodes())
        if nodes != list(range(len(nodes))):
            raise ValueError("FenicsGraph needs node labels 0..V-1 in insertion order "
                             "(use nx.convert_node_labels_to_integers)")
        pos = np.asarray([self.nodes[v]["pos"] for v in nodes], dtype=np.float64)
        edges = np.asarray([[u, v] for u, v in self.edges()], dtype=np.int64).reshape(-1, 2)
        if len(edges) and (edges[:, 0] == edges[:, 1]).any():
            raise ValueError("self-loops are not supported")
        return pos, edges

    def get_mesh(self, n=1):
        """(mesh, mf) with 2^n cells on each edge (fenics_graph.py:58-99)."""
        pos, edges = self._graph_arrays()
        dn = DeviceNetwork(pos, edges, n)
        return Mesh1D(dn), CellMarker(dn)

    # edge-structure mutators drop the cached list of edge attribute dictionaries (edge_attribute_array)
    def _drop_edge_cache(self):
        self.__dict__["_edge_dicts"] = None
        self.__dict__["_edge_attr_cache"] = {}

    def add_edge(self, u, v, **attr):
        self._drop_edge_cache()
        return super().add_edge(u, v, **attr)

    def add_edges_from(self, ebunch, **attr):
        self._drop_edge_cache()
        return super().add_edges_from(ebunch, **attr)

    def add_weighted_edges_from(self, ebunch, weight="weight", **attr):
        self._drop_edge_cache()
        return super().add_weighted_edges_from(ebunch, weight=weight, **attr)

    def remove_edge(self, u, v):
        self._drop_edge_cache()
        return super().remove_edge(u, v)

    def remove_edges_from(self, ebunch):
        self._drop_edge_cache()
        return super().remove_edges_from(ebunch)

    def remove_node(self, n):
        self._drop_edge_cache()
        return super().remove_node(n)

    def remove_nodes_from(self, nodes):
        self._drop_edge_cache()
        return super().remove_nodes_from(nodes)

    def clear(self):
        self._drop_edge_cache()
        return super().clear()

    def clear_edges(self):
        self._drop_edge_cache()
        return super().clear_edges()

    def update(self, edges=None, nodes=None):
        self._drop_edge_cache()
        return super().update(edges=edges, nodes=nodes)

    def make_mesh(self, n=1, comm=None):
        """Generates and stores the mesh with 2^n cells per edge (fenics_graph.py:102-129).

        comm (beyond the reference, which is serial): True or a torch.distributed process group -- the network is
        EDGE-PARTITIONED over the group's ranks, one GPU each (partition.partition_edges: whole subtrees per rank);
        this rank meshes its own edges only, the models built on the graph solve the global problem together, and
        `qp.vector().get_local()` returns the process-local part like a PETSc vector under MPI does
        (`G.local_edges`: global indices of the edges this rank holds, ascending)."""
        pos, edges = self._graph_arrays()
        self.geom_dim = pos.shape[1]
        self.num_edges = len(edges)
        self._group = None
        self.local_edges = None
        if comm is not None and comm is not False:
            import torch.distributed as dist
            group = None if comm is True else comm
            world = dist.get_world_size(group)
            if world > 1:
                from .partition import partition_edges
                epart, plan = partition_edges(edges, len(pos), world)
                self._dn = DeviceNetwork(pos, edges, n, part=(dist.get_rank(group), world), epart=epart)
                self._dn.host_plan = plan
                self._group = group if group is not None else dist.group.WORLD
                self.local_edges = self._dn.edge_ids
        if self.local_edges is None:
            self._dn = DeviceNetwork(pos, edges, n)
        self._systems = {}
        self.mesh = self._dn._host_mesh()
        self.mf = CellMarker(self._dn)
        self.record_bifurcation_and_boundary_nodes()
        self.assign_tangents()

    def _held_edges(self):
        """the (u, v) pairs whose mesh this process holds, in device order (all edges unless partitioned)"""
        ev = list(self.edges())
        if getattr(self, "local_edges", None) is not None:
            ev = [ev[i] for i in self.local_edges]
        return ev

    def make_submeshes(self):
        """Per-edge submeshes and vertex tag functions (fenics_graph.py:132-156)."""
        if self._dn is None:
            raise RuntimeError("call make_mesh() first")
        for i, (u, v) in enumerate(self._held_edges()):
            self.edges[u, v]["submesh"] = SubMesh(self._dn, i)
            self.edges[u, v]["vf"] = VertexMarker(self._dn, i)

    def record_bifurcation_and_boundary_nodes(self):
        """fenics_graph.py:161-184 (from the device vertex tables when a mesh exists)."""
        if self._dn is not None:
            self.bifurcation_ixs = self._dn.bifurcation_ixs()
            self.boundary_ixs = self._dn.boundary_ixs()
            if self._dn.n_lonely:
                deg = self._dn.deg.cpu().numpy()
                for v in np.nonzero(deg == 0)[0]:
                    print(f"Node {v} in G is lonely (i.e. unconnected)")
        else:
            bif, bnd = [], []
            for v in self.nodes():
                k = len(self.in_edges(v)) + len(self.out_edges(v))
                if k == 1:
                    bnd.append(v)
                elif k > 1:
                    bif.append(v)
                else:
                    print(f"Node {v} in G is lonely (i.e. unconnected)")
            self.bifurcation_ixs, self.boundary_ixs = bif, bnd
        self.num_bifurcations = len(self.bifurcation_ixs)

    def _edge_length_array(self):
        """[E] Euclidean edge lengths in G.edges() order (from the device when the whole graph is meshed here)"""
        if self._dn is not None and getattr(self, "local_edges", None) is None:
            return self._dn.elen.cpu().numpy()
        pos, edges = self._graph_arrays()
        return np.sqrt(((pos[edges[:, 1]] - pos[edges[:, 0]]) ** 2).sum(axis=1))

    def compute_edge_lengths(self):
        """edges[e]["length"] = |pos[v] - pos[u]| (fenics_graph.py:187-198)"""
        for e, length in zip(self.edges(), self._edge_length_array().tolist()):
            self.edges[e]["length"] = length

    def compute_vertex_degrees(self):
        """nodes[v]["degree"] = half the total length of the edges at v; degree_min / degree_max
        (fenics_graph.py:200-226)"""
        _, edges = self._graph_arrays()
        if any("length" not in d for _, _, d in self.edges(data=True)):
            self.compute_edge_lengths()
        length = np.asarray([d["length"] for _, _, d in self.edges(data=True)], dtype=np.float64)
        half = 0.5 * (np.bincount(edges[:, 0], weights=length, minlength=len(self)) +
                      np.bincount(edges[:, 1], weights=length, minlength=len(self)))
        for v, hv in zip(self.nodes(), half.tolist()):
            self.nodes[v]["degree"] = hv
        self.degree_min, self.degree_max = float(half.min()), float(half.max())

    def assign_tangents(self):
        """Unit tangent of every edge (fenics_graph.py:230-253): edges[e]['tangent'], G.tangents,
        G.tangent."""
        tang = self._dn.tangents.cpu().numpy()
        for i, (u, v) in enumerate(self._held_edges()):
            self.edges[u, v]["tangent"] = tang[i]
        self.tangents = list(nx.get_edge_attributes(self, "tangent").items())
        self.tangent = _TangentField(self)

    def dds(self, f):
        """derivative df/ds along the graph (fenics_graph.py:255-259)"""
        return dot(grad(f), tangent_components(3))

    def dds_i(self, f, i):
        """derivative df/ds along branch i (fenics_graph.py:261-266)"""
        return dot(grad(f), [float(t) for t in self.tangents[i][1]])

    def get_num_inlets_outlets(self):
        """fenics_graph.py:268-277"""
        tu, tv = self._dn.tag_u.cpu().numpy(), self._dn.tag_v.cpu().numpy()
        return int((tu == BOUN_IN).sum() + (tv == BOUN_IN).sum()), int((tu == BOUN_OUT).sum() + (tv == BOUN_OUT).sum())

    def mark_edge_vfs(self, node_ixs, tag_in, tag_out):
        """Tag, in the vertex markers of the per-edge submeshes, the end of every edge that goes INTO a node of
        `node_ixs` with tag_in and the start of every edge that goes OUT of one with tag_out
        (fenics_graph.py:280-311).  The reference finds the end by an exact coordinate search in the submesh; the
        SubMesh-local number of an edge's ends is known in closed form here (device.submesh_numbering).
        `make_submeshes()` produces the default tags on the GPU (gnx_vertex_tables); this method lets user code
        re-tag on the host view the way it can with the reference."""
        if self._dn is None:
            raise RuntimeError("call make_mesh() first")
        dn = self._dn
        vf, loc, m = dn.vf_host(), dn.loc_host, dn.m
        wanted = set(int(v) for v in node_ixs)
        for i, (u, v) in enumerate(self._held_edges()):
            first, last = (loc[m], loc[0]) if u > v else (loc[0], loc[m])      # SubMesh-local ids of the u / v end
            if v in wanted:
                vf[i, last] = tag_in
            if u in wanted:
                vf[i, first] = tag_out

    # ---- helpers used by the models -------------------------------------------------------------
    def edge_attribute_array(self, key, default=1.0):
        """[E] float array of a (constant-per-edge) attribute; Constants / floats / constant UFL.
        Objects shared by many edges (one `Constant` for the whole graph is the common case) are
        converted once."""
        # Nothing written to any edge attribute since the last walk for this key (the dictionaries count their
        # mutations, _StampedAttrs) and every mutable value found then (Constant) still holds the number it held:
        # the same array again.  2047 edges: 60-90 us of dictionary walk per solve otherwise.
        cache = self.__dict__.setdefault("_edge_attr_cache", {})
        hit = cache.get(key)
        if hit is not None and hit[0] == self._attr_stamp[0] and hit[1] == default and all(c._value == v for c, v in hit[2]):
            return hit[3]
        out = self._edge_attribute_walk(key, default)
        guards = self.__dict__.pop("_edge_attr_guards", None)
        if guards is not None:
            out.setflags(write=False)
            cache[key] = (self._attr_stamp[0], default, guards, out)
        return out

    def _edge_attribute_walk(self, key, default):
        # same order as self.edges(): by source node, then successor insertion order.  The attribute
        # dictionaries themselves are stable objects: the list of them is kept while the edge count stands.
        dicts = self.__dict__.get("_edge_dicts")
        if dicts is None:                                   # dropped by every method that adds / removes edges
            dicts = [d for nbrs in self._succ.values() for d in nbrs.values()]
            if getattr(self, "local_edges", None) is not None:          # partitioned: this rank's edges only
                dicts = [dicts[i] for i in self.local_edges]
            self.__dict__["_edge_dicts"] = dicts
        try:                                                  # (values are re-read on every solve: a user may change them)
            vals = list(map(itemgetter(key), dicts))
        except KeyError:                                      # some edge lacks the attribute: flow_models.py:228-231 default
            vals = [d.get(key, default) for d in dicts]
        if vals and vals.count(vals[0]) == len(vals):                     # one shared object / one value
            vals, rep = vals[:1], len(vals)
        else:
            rep = 1
        conv = {}
        guards = []                  # (Constant, value) pairs the cached array depends on; None: do not cache
        out = np.empty(len(vals))
        for i, v in enumerate(vals):
            k = id(v)
            f = conv.get(k)
            if f is None:
                if isinstance(v, Constant):
                    if guards is not None:
                        guards.append((v, v._value))
                elif not isinstance(v, (int, float, np.integer, np.floating)):
                    guards = None                          # (a UFL expression of parameters: re-evaluated every time)
                if isinstance(v, UFLExpr):
                    if v.expr.free_symbols - set(v.params):
                        raise VaryingEdgeAttribute(key, list(np.repeat(np.asarray(vals, dtype=object), rep)) if rep > 1 else vals)
                    f = float(v.expr.subs({s: float(c) for s, c in v.params.items()}))
                elif isinstance(v, Expression):
                    raise VaryingEdgeAttribute(key, list(np.repeat(np.asarray(vals, dtype=object), rep)) if rep > 1 else vals)
                else:
                    f = float(v)
                conv[k] = f
            out[i] = f
        self.__dict__["_edge_attr_guards"] = guards
        return np.repeat(out, rep) if rep > 1 else out


class _EdgewiseExpression(UserExpression):
    """expression on the global mesh defined edge by edge: a cell belongs to edge `mf[cell]`"""

    def __init__(self, G, **kwargs):
        self.G = G
        self._tangent = np.asarray([t for _, t in G.tangents], dtype=np.float64)      # [E, gdim]
        super().__init__(**kwargs)

    def _edge(self, cell):
        return self.G.mf[cell.index]


class GlobalFlux(_EdgewiseExpression):
    """vector flux = scalar flux of the edge times its unit tangent (fenics_graph.py:314-351);
    `qs`: one function per edge, or a single function on the global mesh"""

    def __init__(self, G, qs, **kwargs):
        self.qs = qs if isinstance(qs, list) else [qs]
        super().__init__(G, **kwargs)

    def eval_cell(self, values, x, cell):
        e = self._edge(cell)
        q = self.qs[e if len(self.qs) > 1 else 0]
        values[: self.G.geom_dim] = q(x) * self._tangent[e]

    def value_shape(self):
        return (self.G.geom_dim,)


class GlobalCrossectionFlux(_EdgewiseExpression):
    """the per-edge scalar fluxes as one scalar field on the global mesh (fenics_graph.py:354-375)"""

    def __init__(self, G, qs, **kwargs):
        self.qs = qs
        super().__init__(G, **kwargs)

    def eval_cell(self, values, x, cell):
        values[0] = self.qs[self._edge(cell)](x)

    def value_shape(self):
        return ()


class TangentFunction(_EdgewiseExpression):
    """unit tangent of the edge a cell lies on (fenics_graph.py:378-401); the kernels never call it"""

    def __init__(self, G, degree, **kwargs):
        self.degree = degree
        super().__init__(G, **kwargs)

    def eval_cell(self, values, x, cell):
        values[: self.G.geom_dim] = self._tangent[self._edge(cell)]

    def value_shape(self):
        return (self.G.geom_dim,)


def copy_from_nx_graph(G_nx):
    """FenicsGraph with the nodes, edges (in `G_nx.edges()` order -- that fixes the direction an undirected graph's
    edges get) and attributes of a networkx graph; attribute dictionaries are copied, values shared
    (fenics_graph.py:404-434)"""
    G = FenicsGraph()
    G.graph.update(G_nx.graph)
    G.add_nodes_from((v, dict(attrs)) for v, attrs in G_nx.nodes(data=True))
    G.add_edges_from((u, v, dict(attrs)) for u, v, attrs in G_nx.edges(data=True))
    return G


def nxgraph_attribute_to_dolfin(G, attr):
    """DG0 function of an edge attribute (fenics_graph.py:437-460; the reference interpolates into
    DG1, whose cellwise-constant values are the same)."""
    from .function import Function, FunctionSpace
    import torch
    vals = np.asarray(list(nx.get_edge_attributes(G, attr).values()), dtype=np.float64)
    f = Function(FunctionSpace(G.mesh, "DG", 0), net=G._dn, role="cell")
    f.vector().tensor.copy_(torch.from_numpy(np.repeat(vals, G._dn.m)).to(G._dn.device))
    f.vector().touch()
    return f
