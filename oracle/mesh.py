"""Oracle, mesh level: DOLFIN-style interval mesh, uniform bisection, per-edge submeshes, tags.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates SURVEY.md Appendix A rules A1-A5
(DOLFIN 2019.1 / fenics_ii behaviour, recalled -- PARITY UNPINNED against real DOLFIN
output) at the reference call sites fenics_graph.py:58-99,132-156,280-311.

Everything here is done the *literal* way (level-by-level refinement, first-encounter
numbering, exact float coordinate search) so that it is algorithmically independent of the
closed forms the CUDA kernels use.
"""
import numpy as np

from .graph import BIF_IN, BIF_OUT, BOUN_IN, BOUN_OUT, node_tables

__all__ = ["init_mesh", "refine", "get_mesh", "submesh", "all_submeshes", "mark_vfs",
           "OracleMesh", "build_mesh"]


def init_mesh(pos, edges):
    """MeshEditor open/add_vertex/add_cell/close (fenics_graph.py:75-88).

    A1: close(order=True) stores each interval cell's two vertex ids ascending.
    """
    coords = np.array(pos, dtype=np.float64, copy=True)
    cells = np.sort(np.asarray(edges, dtype=np.int64).reshape(-1, 2), axis=1)
    return coords, cells


def refine(coords, cells):
    """One `refine(mesh)` (fenics_graph.py:96).  A2 + A3.

    * old vertices are copied, one midpoint per cell is appended IN CELL ORDER
      (new id = num_old_vertices + cell index);
    * midpoint = MeshEntity::midpoint(): x = 0; x += x_a; x += x_b; x /= 2;
    * children of parent c: 2c = (v0,new), 2c+1 = (new,v1) with (v0,v1) the parent's STORED
      order; the refined mesh is closed with ordering, so child 2c+1 is stored (v1,new).
    Returns (coords, cells, parent_cell).
    """
    nv, nc = len(coords), len(cells)
    mid = np.zeros((nc, coords.shape[1]), dtype=np.float64)
    mid += coords[cells[:, 0]]
    mid += coords[cells[:, 1]]
    mid /= 2.0
    new_ids = nv + np.arange(nc, dtype=np.int64)
    child = np.empty((2 * nc, 2), dtype=np.int64)
    child[0::2, 0] = cells[:, 0]
    child[0::2, 1] = new_ids
    child[1::2, 0] = new_ids
    child[1::2, 1] = cells[:, 1]
    child = np.sort(child, axis=1)
    parent = np.repeat(np.arange(nc, dtype=np.int64), 2)
    return np.vstack([coords, mid]), child, parent


def get_mesh(pos, edges, n=1):
    """FenicsGraph.get_mesh (fenics_graph.py:58-99): (coords, cells, mf).

    A4: adapt(mf, refined) copies mf[parent_cell[c]].
    """
    coords, cells = init_mesh(pos, edges)
    mf = np.arange(len(cells), dtype=np.int64)          # :91-92
    for _ in range(n):                                   # :95-97
        coords, cells, parent = refine(coords, cells)
        mf = mf[parent]
    return coords, cells, mf


def submesh(coords, cells, mf, marker):
    """xii.EmbeddedMesh(mf, marker) -> dolfin.SubMesh (fenics_graph.py:139).  A5.

    Marked parent cells ascending; submesh vertex ids by FIRST ENCOUNTER walking each cell's
    stored vertex pair; coordinates copied bit-for-bit; cells re-sorted in local numbering.
    Returns (sub_coords, sub_cells, parent_vertex, parent_cell).
    """
    parent_cell = np.nonzero(mf == marker)[0]
    local = {}
    parent_vertex = []
    sub_cells = np.empty((len(parent_cell), 2), dtype=np.int64)
    for k, c in enumerate(parent_cell):
        for j in range(2):
            v = int(cells[c, j])
            if v not in local:
                local[v] = len(parent_vertex)
                parent_vertex.append(v)
            sub_cells[k, j] = local[v]
    sub_cells = np.sort(sub_cells, axis=1)
    parent_vertex = np.asarray(parent_vertex, dtype=np.int64)
    return coords[parent_vertex].copy(), sub_cells, parent_vertex, parent_cell


def all_submeshes(coords, cells, mf, num_edges):
    """The E-fold loop of fenics_graph.py:138-139, vectorised (same semantics as `submesh`).

    Every edge has the same number m of cells, so results are stacked:
    sub_coords[E,m+1,gdim], sub_cells[E,m,2], parent_vertex[E,m+1], parent_cell[E,m].
    """
    order = np.argsort(mf, kind="stable")
    m = len(cells) // num_edges
    parent_cell = order.reshape(num_edges, m)
    assert (mf[parent_cell] == np.arange(num_edges)[:, None]).all()
    gd = coords.shape[1]
    sub_coords = np.empty((num_edges, m + 1, gd))
    sub_cells = np.empty((num_edges, m, 2), dtype=np.int64)
    parent_vertex = np.empty((num_edges, m + 1), dtype=np.int64)
    for i in range(num_edges):
        walk = cells[parent_cell[i]].ravel()
        uniq, first = np.unique(walk, return_index=True)
        seq = uniq[np.argsort(first, kind="stable")]           # first-encounter order
        assert len(seq) == m + 1
        lookup = np.searchsorted(uniq, walk)
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
        sub_cells[i] = np.sort(rank[lookup].reshape(m, 2), axis=1)
        parent_vertex[i] = seq
        sub_coords[i] = coords[seq]
    return sub_coords, sub_cells, parent_vertex, parent_cell


def mark_vfs(pos, edges, sub_coords, bif_ixs, bnd_ixs):
    """make_submeshes tagging (fenics_graph.py:141-156) via mark_edge_vfs (:280-311).

    vf[E,m+1] (size_t, init 0).  For node n in node_ixs: every in-edge gets `tag_in` at the
    FIRST submesh vertex whose coordinates equal pos[n] exactly, every out-edge `tag_out`.
    Bifurcations: (BIF_IN, BIF_OUT); boundary nodes: (BOUN_OUT, BOUN_IN)  <- swapped on purpose
    (:156-158).
    """
    E = len(edges)
    vf = np.zeros(sub_coords.shape[:2], dtype=np.int64)
    in_edges = {}
    out_edges = {}
    for i, (u, v) in enumerate(edges):
        out_edges.setdefault(int(u), []).append(i)
        in_edges.setdefault(int(v), []).append(i)

    def mark(node_ixs, tag_in, tag_out):
        for nd in node_ixs:
            p = np.asarray(pos[nd])
            for i in in_edges.get(nd, []):
                hit = np.where((sub_coords[i] == p).all(axis=1))[0]
                if len(hit) > 0:
                    vf[i, hit[0]] = tag_in
            for i in out_edges.get(nd, []):
                hit = np.where((sub_coords[i] == p).all(axis=1))[0]
                if len(hit) > 0:
                    vf[i, hit[0]] = tag_out

    mark(bif_ixs, BIF_IN, BIF_OUT)
    mark(bnd_ixs, BOUN_OUT, BOUN_IN)
    assert E == len(vf)
    return vf


class OracleMesh:
    """Everything `make_mesh(n)` + `make_submeshes()` leave on a FenicsGraph, as arrays."""

    def __init__(self, pos, edges, n):
        from .graph import edge_tangents
        self.pos = np.asarray(pos, dtype=np.float64)
        self.edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        self.n = n
        self.m = 1 << n
        self.V, self.gdim = self.pos.shape
        self.E = len(self.edges)
        self.coords, self.cells, self.mf = get_mesh(self.pos, self.edges, n)
        self.bif_ixs, self.bnd_ixs, self.degree = node_tables(self.V, self.edges)
        self.B = len(self.bif_ixs)
        self.tangents = edge_tangents(self.pos, self.edges)
        (self.sub_coords, self.sub_cells,
         self.parent_vertex, self.parent_cell) = all_submeshes(self.coords, self.cells, self.mf, self.E)
        self.vf = mark_vfs(self.pos, self.edges, self.sub_coords, self.bif_ixs, self.bnd_ixs)


def build_mesh(pos, edges, n):
    return OracleMesh(pos, edges, n)
