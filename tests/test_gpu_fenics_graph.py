"""/root/reference/tests/test_fenics_graph.py:14-56 through the drop-in API, plus the attributes
`make_mesh` / `make_submeshes` leave on the graph (fenics_graph.py:102-156) against the oracle and
the golden vectors produced by the reference's own Python code (tests/golden/graphs.json)."""
import networkx as nx
import numpy as np
import pytest

import oracle
from _util import golden_graphs, fromhex

pytestmark = pytest.mark.gpu


def test_fenics_graph():
    from graphnics_b200 import Y_bifurcation
    G = Y_bifurcation()
    G.make_mesh()
    mesh = G.mesh
    mesh_c = mesh.coordinates()
    for n in G.nodes:
        vertex_c = G.nodes[n]["pos"]
        vertex_ix = np.where((mesh_c == vertex_c).all(axis=1))[0]
        assert len(vertex_ix) == 1, "vertex coordinate is not a mesh coordinate"


def test_compute_vertex_degrees():
    from graphnics_b200 import Y_bifurcation, near
    G = Y_bifurcation()
    G.compute_vertex_degrees()
    degrees = nx.get_node_attributes(G, 'degree')
    assert near(degrees[0], 0.25)
    assert near(degrees[1], 0.9571, 0.01)


def test_tangent_vector():
    from graphnics_b200 import Y_bifurcation, honeycomb
    G = Y_bifurcation()
    G.make_mesh(n=3)
    G.make_submeshes()
    first_tangent = G.edges()[(0, 1)]['tangent']
    assert np.allclose(first_tangent, [0, 1])
    G = honeycomb(2, 2)
    v1, v2 = list(G.edges())[0]
    v1_pos = np.asarray(G.nodes()[v1]['pos'])
    v2_pos = np.asarray(G.nodes()[v2]['pos'])
    expected_tangent = np.asarray(v2_pos - v1_pos)
    expected_tangent *= 1 / np.linalg.norm(expected_tangent)
    computed_tangent = G.edges()[(v1, v2)]['tangent']
    assert np.allclose(expected_tangent, computed_tangent)


@pytest.mark.parametrize("gen,args,n", [("Y_bifurcation", (), 4), ("YY_bifurcation", (), 3), ("honeycomb", (4, 4), 2),
                                        ("line_graph", (3, 3), 5)])
def test_graph_attributes_after_make_mesh_and_submeshes(gen, args, n):
    import graphnics_b200 as gn
    G = getattr(gn, gen)(*args)
    G.make_mesh(n)
    G.make_submeshes()
    pos = np.asarray([G.nodes[v]["pos"] for v in G.nodes()], dtype=np.float64)
    edges = np.asarray(list(G.edges()), dtype=np.int64)
    om = oracle.build_mesh(pos, edges, n)
    assert G.geom_dim == pos.shape[1] and G.num_edges == len(edges)
    assert np.array_equal(G.mesh.coordinates().view(np.int64), om.coords.view(np.int64))
    assert np.array_equal(G.mesh.cells(), om.cells)
    assert np.array_equal(G.mf.array().astype(np.int64), om.mf)
    assert G.bifurcation_ixs == om.bif_ixs and G.boundary_ixs == om.bnd_ixs
    assert G.num_bifurcations == om.B
    assert [e for e, _ in G.tangents] == [tuple(e) for e in edges.tolist()]
    for i, (u, v) in enumerate(G.edges()):
        assert np.array_equal(np.asarray(G.edges[u, v]["tangent"]).view(np.int64), om.tangents[i].view(np.int64))
        sub, vf = G.edges[u, v]["submesh"], G.edges[u, v]["vf"]
        assert np.array_equal(sub.coordinates().view(np.int64), om.sub_coords[i].view(np.int64))
        assert np.array_equal(sub.cells(), om.sub_cells[i])
        assert np.array_equal(np.asarray(vf.array()), om.vf[i])
        assert np.array_equal(sub.parent_vertex_indices, om.parent_vertex[i])
        assert np.array_equal(sub.parent_cell_indices, om.parent_cell[i])
        assert G.mf[int(sub.parent_cell_indices[0])] == i
    ins, outs = G.get_num_inlets_outlets()
    assert ins == int((om.vf == 3).sum()) and outs == int((om.vf == 4).sum())
    G.compute_edge_lengths()
    L = oracle.edge_lengths(pos, edges)
    assert np.allclose([G.edges[e]["length"] for e in G.edges()], L, rtol=1e-15)


def test_golden_tables_from_the_reference_code():
    """bifurcation / boundary lists, tangents and inlet/outlet counts the REFERENCE's own Python produced
    (fenics stubbed; tests/golden/make_golden.py) for every golden graph, against the device tables"""
    from graphnics_b200.device import DeviceNetwork
    for name, c in golden_graphs().items():
        pos, edges = fromhex(c["pos"]), np.asarray(c["edges"], dtype=np.int64)
        dn = DeviceNetwork(pos, edges, 1)
        assert dn.bifurcation_ixs() == c["bifurcation_ixs"], name
        assert dn.boundary_ixs() == c["boundary_ixs"], name
        if "tangents" in c:
            assert np.array_equal(dn.tangents.cpu().numpy().view(np.int64), fromhex(c["tangents"]).view(np.int64)), name
