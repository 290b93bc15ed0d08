"""`.vtp` / `.pvd` output for ParaView's TubeFilter, drop-in for graphnics/plot.py:19-170 (SURVEY.md 8f
rank 3).  The reference goes through the vtk package; the files are written here directly as VTK XML
PolyData (ascii), so nothing beyond numpy is needed."""
import os

import networkx as nx
import numpy as np

from .coefficients import UserExpression

__all__ = ["TubeFile", "TubeRadius"]


class TubeRadius(UserExpression):
    """Expression that evaluates the FenicsGraph edge attribute 'radius' at each point (plot.py:19-30)."""

    def __init__(self, G, **kwargs):
        self.G = G
        super().__init__(**kwargs)

    def eval_cell(self, value, x, cell):
        edge = list(self.G.edges())[self.G.mf[cell.index]]
        value[0] = self.G.edges()[edge]["radius"]


_PVD_HEAD = '<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n  <Collection>\n'
_PVD_FOOT = "</Collection>\n</VTKFile>"


def _values_at_vertices(func, mesh_view, dn):
    """func at every vertex of the global mesh"""
    coords = mesh_view.coordinates()
    role = getattr(func, "_role", None)
    if role == "vertex":
        return np.asarray(func.vector().get_local(), dtype=float)
    if role == "cell":                      # DG0: the value of one adjacent cell
        vals = np.asarray(func.vector().get_local(), dtype=float)
        cells = mesh_view.cells()
        out = np.empty(len(coords))
        out[cells[:, 1]] = vals
        out[cells[:, 0]] = vals
        return out
    return np.asarray([float(func(c)) for c in coords])


class TubeFile:
    """`.pvd` collection of `.vtp` files holding a network function and the vessel radius as point data
    (plot.py:33-161).  Usage:  TubeFile(G, "out.pvd") << f    or    << (f, step)"""

    def __init__(self, G, fname, **kwargs):
        f_name, f_ext = os.path.splitext(fname)
        assert f_ext == ".pvd", "TubeFile must have .pvd file ending"
        self.fname, self.G = f_name, G
        assert self.G.geom_dim == 3, f"Coordinates are {self.G.geom_dim}d, they need to be 3d"
        assert len(nx.get_edge_attributes(self.G, "radius")) > 0, "Graph must have radius attribute"
        with open(fname, "w") as fh:
            fh.write(_PVD_HEAD + _PVD_FOOT)

    def __lshift__(self, func):
        func, i = func if isinstance(func, tuple) else (func, 0)
        G = self.G
        dn = G._dn
        mesh = G.mesh
        coords, cells = mesh.coordinates(), mesh.cells()
        name = func.name() if callable(getattr(func, "name", None)) else "f"
        data = _values_at_vertices(func, mesh, dn)
        # radius of the edge a vertex lies on (graph nodes: one of their edges), like the reference's DG0
        # radius function on the unrefined mesh
        radius_e = np.asarray(list(nx.get_edge_attributes(G, "radius").values()), dtype=float)
        edge_of_cell = np.arange(len(cells)) >> dn.n
        rad = np.empty(len(coords))
        rad[cells[:, 1]] = radius_e[edge_of_cell]
        rad[cells[:, 0]] = radius_e[edge_of_cell]
        fmt = lambda a: " ".join(repr(float(v)) for v in np.asarray(a).ravel())
        nP, nL = len(coords), len(cells)
        xml = ['<?xml version="1.0"?>', '<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian">', "  <PolyData>",
               f'    <Piece NumberOfPoints="{nP}" NumberOfVerts="0" NumberOfLines="{nL}" NumberOfStrips="0" NumberOfPolys="0">',
               "      <PointData>",
               f'        <DataArray type="Float64" Name="{name}" NumberOfComponents="1" format="ascii">{fmt(data)}</DataArray>',
               f'        <DataArray type="Float64" Name="radius" NumberOfComponents="1" format="ascii">{fmt(rad)}</DataArray>',
               "      </PointData>", "      <Points>",
               f'        <DataArray type="Float64" NumberOfComponents="3" format="ascii">{fmt(coords)}</DataArray>',
               "      </Points>", "      <Lines>",
               '        <DataArray type="Int64" Name="connectivity" format="ascii">'
               + " ".join(str(int(v)) for v in cells.ravel()) + "</DataArray>",
               '        <DataArray type="Int64" Name="offsets" format="ascii">'
               + " ".join(str(2 * (k + 1)) for k in range(nL)) + "</DataArray>",
               "      </Lines>", "    </Piece>", "  </PolyData>", "</VTKFile>"]
        with open(f"{self.fname}{int(i):06d}.vtp", "w") as fh:
            fh.write("\n".join(xml) + "\n")
        with open(self.fname + ".pvd") as fh:
            content = fh.read().splitlines()
        short = self.fname.split("/")[-1]
        entry = f'<DataSet timestep="{i}" part="0" file="{short}{int(i):06d}.vtp" />'
        with open(self.fname + ".pvd", "w") as fh:
            fh.write("\n".join(content[:-2] + [entry] + content[-2:]))
        return self
