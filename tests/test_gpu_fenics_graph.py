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


def test_postprocessing_helpers():
    """GlobalFlux / GlobalCrossectionFlux (fenics_graph.py:314-375), nxgraph_attribute_to_dolfin (:437-460),
    point evaluation of the solution components, tabulate_dof_coordinates (demo/Graphnics demo.ipynb:282)"""
    import networkx as nx
    import graphnics_b200 as gn
    G = gn.Y_bifurcation()
    G.make_mesh(5)
    G.make_submeshes()
    for i, e in enumerate(G.edges()):
        G.edges[e]["Res"] = gn.Constant(1.0 + i)
        G.edges[e]["radius"] = 0.1 * (i + 1)
    model = gn.MixedHydraulicNetwork(G, p_bc=gn.Expression("2-x[1]", degree=2))
    qp = model.solve()
    E = G.num_edges
    # SURVEY.md Appendix C closed form: q constant per edge
    qs = [0.741549228561398, 0.444929537136839, 0.296619691424559]
    for i in range(E):
        assert abs(qp[i](G.nodes[1]["pos"]) - qs[i]) < 1e-11
        assert np.allclose(qp[i].vector().get_local(), qs[i], rtol=1e-11)
        assert qp[i].function_space().tabulate_dof_coordinates().shape == (G._dn.m + 1, 2)
    assert abs(qp[E + 1](0.0, 0.0) - 1.629225385719301) < 1e-11          # the multiplier = pressure at the bifurcation
    assert abs(qp[E]((0.0, 0.25)) - (2.0 - qs[0] * 1.0 * 0.25)) < 1e-2     # p drops linearly along edge 0 (DG0: cell mean, h/2 off)
    flux = gn.GlobalFlux(G, [qp[i] for i in range(E)])
    val = np.zeros(2)

    class _Cell:
        index = 0
    flux.eval_cell(val, np.array([0.0, 0.1]), _Cell())
    assert np.allclose(val, [0.0, qs[0]], atol=1e-11)                     # scalar flux times the tangent (0, 1)
    cs = gn.GlobalCrossectionFlux(G, [qp[i] for i in range(E)])
    v1 = np.zeros(1)
    cs.eval_cell(v1, np.array([0.0, 0.1]), _Cell())
    assert abs(v1[0] - qs[0]) < 1e-11
    r = gn.nxgraph_attribute_to_dolfin(G, "radius")
    assert abs(r(0.0, 0.2) - 0.1) < 1e-15 and abs(r(0.25, 0.75) - 0.3) < 1e-15


def test_network_poisson_and_stokes_api():
    """NetworkPoisson (flow_models.py:15-50) on a line: -u'' = 1, u(0)=u(1)=0 -> u = x(1-x)/2;
    NetworkStokes (:353-430; broken as shipped, here mu := nu) reduces to the mixed model for nu = 0"""
    import graphnics_b200 as gn
    G = gn.line_graph(2)
    G.make_mesh(7)
    u = gn.NetworkPoisson(G, f=gn.Constant(1.0), p_bc=gn.Constant(0.0)).solve()
    for x in (0.25, 0.5, 0.75):
        assert abs(u(x, 0.0) - 0.5 * x * (1 - x)) < 1e-4
    G = gn.Y_bifurcation()
    G.make_mesh(4)
    G.make_submeshes()
    a = gn.MixedHydraulicNetwork(G, p_bc=gn.Expression("x[1]", degree=2)).solve().vector().get_local()
    s0 = gn.NetworkStokes(G, p_bc=gn.Expression("x[1]", degree=2), nu=gn.Constant(0.0)).solve().vector().get_local()
    assert np.allclose(a, s0, rtol=1e-12, atol=1e-14)
    s1 = gn.NetworkStokes(G, p_bc=gn.Expression("x[1]", degree=2), nu=gn.Constant(0.5)).solve().vector().get_local()
    assert np.allclose(a[: 3 * 17], s1[: 3 * 17], rtol=1e-9)      # q is constant per edge: the viscous term vanishes
