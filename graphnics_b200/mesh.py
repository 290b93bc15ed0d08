"""Light host-side views of the device mesh with the dolfin method names graphnics users touch
(SURVEY.md 8b): `G.mesh.coordinates()`, `G.mf[cell]`, `edges[e]["submesh"]`, `edges[e]["vf"].array()`.
All data stay on the GPU; numpy copies are made lazily and cached.
"""
import numpy as np

__all__ = ["Mesh1D", "SubMesh", "CellMarker", "VertexMarker"]


class _Dim:
    def __init__(self, d):
        self._d = d

    def dim(self):
        return self._d


class Mesh1D:
    """Global 1-D mesh of the network (dolfin.Mesh stand-in), numbering of fenics_graph.py:58-99."""

    def __init__(self, dn):
        self._dn = dn
        self._coords = None
        self._cells = None

    def coordinates(self):
        if self._coords is None:
            self._coords = self._dn.coords.cpu().numpy()
        return self._coords

    def cells(self):
        if self._cells is None:
            self._cells = self._dn.cells.cpu().numpy()
        return self._cells

    def num_cells(self):
        return self._dn.num_cells

    def num_vertices(self):
        return self._dn.num_vertices

    def num_entities(self, d):
        return self.num_vertices() if d == 0 else self.num_cells()

    def geometry(self):
        return _Dim(self._dn.gdim)

    def topology(self):
        return _Dim(1)

    def geometric_dimension(self):
        return self._dn.gdim

    def _h(self):
        return self._dn.h.cpu().numpy()

    def hmin(self):
        return float(self._h().min())

    def hmax(self):
        return float(self._h().max())

    def cell_midpoints(self):
        c, x = self.cells(), self.coordinates()
        return 0.5 * (x[c[:, 0]] + x[c[:, 1]])


class CellMarker:
    """`G.mf`: cell -> edge index (MeshFunction('size_t', mesh, 1), fenics_graph.py:91-97)."""

    def __init__(self, dn):
        self._dn = dn
        self._arr = None

    def array(self):
        if self._arr is None:
            if self._dn.mf is not None:
                self._arr = self._dn.mf.cpu().numpy().astype(np.uint64)
            else:
                self._arr = (np.arange(self._dn.num_cells) >> self._dn.n).astype(np.uint64)
        return self._arr

    def __getitem__(self, c):
        idx = c.index() if callable(getattr(c, "index", None)) else getattr(c, "index", c)
        return int(idx) >> self._dn.n          # cells of edge e are [e*m, (e+1)*m)

    def __len__(self):
        return self._dn.num_cells


class SubMesh:
    """`edges[e]['submesh']`: xii.EmbeddedMesh(mf, i) stand-in (fenics_graph.py:139).  Vertex
    numbering by first encounter (same local pattern for every edge)."""

    def __init__(self, dn, i):
        self._dn, self.i = dn, int(i)
        self._coords = None

    @property
    def parent_vertex_indices(self):
        return self._dn.parent_vertex_host()[self.i]

    @property
    def parent_cell_indices(self):
        return self.i * self._dn.m + np.arange(self._dn.m)

    def coordinates(self):
        if self._coords is None:
            from . import fenics_graph  # noqa: F401
            gmesh = self._dn._host_mesh()
            self._coords = gmesh.coordinates()[self.parent_vertex_indices]
        return self._coords

    def cells(self):
        return self._dn.sub_cells_host()

    def num_cells(self):
        return self._dn.m

    def num_vertices(self):
        return self._dn.m + 1

    def geometry(self):
        return _Dim(self._dn.gdim)

    def topology(self):
        return _Dim(1)


class VertexMarker:
    """`edges[e]['vf']`: vertex tags of SubMesh e (fenics_graph.py:141-156)."""

    def __init__(self, dn, i):
        self._dn, self.i = dn, int(i)

    def array(self):
        return self._dn.vf_host()[self.i]
