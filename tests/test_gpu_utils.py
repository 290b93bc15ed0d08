"""/root/reference/tests/test_graph_utils.py:18-108 and tests/test_vtp_plot.py:20-50 through the drop-in
API (pre-/post-processing helpers around the hot path)."""
import glob
import os

import networkx as nx
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_graph_color():
    from graphnics_b200 import line_graph, Y_bifurcation, honeycomb, color_graph
    G = line_graph(3)
    color_graph(G)
    assert G.edges()[(0, 1)]["color"] == G.edges()[(1, 2)]["color"]
    G = Y_bifurcation()
    color_graph(G)
    assert len(set(nx.get_edge_attributes(G, "color").values())) == 3
    G = honeycomb(2, 3)
    color_graph(G)
    assert len(set(list(nx.get_edge_attributes(G, "color").values()))) == 18


def test_dist_from_source():
    from graphnics_b200 import line_graph, YY_bifurcation, DistFromSource, FunctionSpace, interpolate, near
    G = line_graph(10)
    dist_from_source = DistFromSource(G, 0, degree=2)
    V = FunctionSpace(G.mesh, "CG", 1)
    dist_from_source_i = interpolate(dist_from_source, V)
    coords = V.tabulate_dof_coordinates()
    lengths = np.linalg.norm(coords, axis=1)
    discrepancy = np.linalg.norm(np.asarray(dist_from_source_i.vector().get_local()) - np.asarray(lengths))
    assert near(discrepancy, 0, 1e-13), f"distance function discrepancy: {discrepancy}"
    G = YY_bifurcation()
    dist_from_source = DistFromSource(G, 0, degree=2)
    G.compute_edge_lengths()
    source = 0
    for node in [3, 4, 5]:
        path = nx.shortest_path(G, source, node)[1:]
        v1 = source
        dist = 0
        for v2 in path:
            dist += G.edges()[(v1, v2)]["length"]
            v1 = v2
        assert near(dist_from_source(G.nodes()[node]["pos"]), dist, 1e-14), f"Distance not computed correctly for node {node}"


def test_Murrays_law_on_double_bifurcation():
    from graphnics_b200 import YY_bifurcation, assign_radius_using_Murrays_law, near
    G = YY_bifurcation()
    G = assign_radius_using_Murrays_law(G, start_node=0, start_radius=1)
    for v in G.nodes():
        es_in = list(G.in_edges(v))
        es_out = list(G.out_edges(v))
        if len(es_out) == 0 or len(es_in) == 0:
            continue
        r_p_cubed = sum(G.edges()[e]["radius"] ** 3 for e in es_in)
        r_d_cubed = sum(G.edges()[e]["radius"] ** 3 for e in es_out)
        assert near(r_p_cubed, r_d_cubed, 1e-6), "Murrays law not satisfied"


def test_vtp_plot(tmp_path):
    from graphnics_b200 import YY_bifurcation, TubeFile, FunctionSpace, Function
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        G = YY_bifurcation(dim=3)
        for e in G.edges():
            G.edges()[e]['radius'] = 1
        ffile = TubeFile(G, 'test.pvd')
        V = FunctionSpace(G.mesh, 'CG', 1)
        v = Function(V)
        v.rename('u', '0.0')
        ffile << v
        for i in range(10):
            for j, e in enumerate(G.edges()):
                G.edges()[e]['radius'] = j + i
            v.rename('u', '0.0')
            v.name()
            ffile << (v, i)
        files = sorted(glob.glob('*.vtp'))
        assert len(files) == 10
        txt = open(files[-1]).read()
        nP = G.mesh.num_vertices()
        assert f'NumberOfPoints="{nP}"' in txt and 'Name="radius"' in txt and 'Name="u"' in txt
        import xml.etree.ElementTree as ET
        root = ET.parse(files[-1]).getroot()
        arrays = {a.get("Name"): a for a in root.iter("DataArray")}
        rad = np.array(arrays["radius"].text.split(), dtype=float)
        assert len(rad) == nP and set(rad) <= set(float(j + 9) for j in range(G.num_edges))
        conn = np.array(arrays["connectivity"].text.split(), dtype=int).reshape(-1, 2)
        assert np.array_equal(conn, G.mesh.cells())
        pvd = open('test.pvd').read()
        assert pvd.count("<DataSet") == 11 and pvd.strip().endswith("</VTKFile>")
    finally:
        os.chdir(cwd)


def test_secomb_reader(tmp_path):
    from graphnics_b200.read_network import get_graph
    txt = ("header1\nheader2\n  1  5  1  2 10.0  0.5  0.0 0.0 0.0  1.0 0.0 0.0\n  2  5  2  3  8.0  0.4  1.0 0.0 0.0  1.0 1.0 0.0\n"
           "  3  5  2  4  6.0  0.3  1.0 0.0 0.0  2.0 -1.0 0.5\ntrailer1\ntrailer2")
    p = tmp_path / "net.txt"
    p.write_text(txt)
    G = get_graph(str(p))
    assert list(G.edges()) == [(0, 1), (1, 2), (1, 3)]
    assert [G.edges[e]["radius"] for e in G.edges()] == [5.0, 4.0, 3.0]
    G.make_mesh(2)
    assert G.bifurcation_ixs == [1] and G.geom_dim == 3
