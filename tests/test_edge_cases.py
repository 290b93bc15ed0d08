"""Edge cases of the path (empty / ragged / extreme inputs), CPU-only parts: argument validation in the
C ABI happens before any CUDA call, host-side graph validation, size arithmetic at the int32 limit."""
import ctypes as C

import networkx as nx
import numpy as np
import pytest


def test_size_arithmetic_and_int32_guard():
    from graphnics_b200 import _lib
    lib = _lib.load()
    s = _lib.gnx_network_t()
    # C5: 19 generations, n = 7
    s.V, s.E, s.gdim, s.n, s.B, s.n_bnd, s.I = 524288, 524287, 2, 7, 262143, 262145, 786429
    assert lib.gnx_mixed_ndofs(C.byref(s)) == 524287 * (2 * 128 + 1) + 262143 == 135003902
    assert lib.gnx_mixed_nnz(C.byref(s)) == 524287 * (8 * 128 + 1) + 2 * 786429 + 262143
    assert lib.gnx_primal_ndofs(C.byref(s)) == 524287 * 128 + 524288 + 524287 * 127
    # one refinement too many for int32 indices: rejected with an error message, no CUDA call made
    s.n = 12
    dummy = (C.c_int32 * 4)()
    for name in ("edges", "bix", "inc_ptr", "inc_edge", "lam_ptr", "ebif_prefix", "loc", "tloc", "bif_nodes", "h"):
        setattr(s, name, C.addressof(dummy))
    assert lib.gnx_mixed_nnz(C.byref(s)) > 2 ** 31
    rc = lib.gnx_build_pattern_mixed(C.byref(s), C.addressof(dummy), C.addressof(dummy), None)
    assert rc == -1 and b"bad argument" in lib.gnx_last_error()


def test_argument_validation_without_gpu():
    from graphnics_b200 import _lib
    lib = _lib.load()
    assert lib.gnx_refine_1d(None, None, 4, 3, 2, 1, None, None, None, None, None) == -1
    assert b"bad argument" in lib.gnx_last_error()
    dummy = (C.c_double * 4)()
    p = C.addressof(dummy)
    assert lib.gnx_refine_1d(p, p, 4, 3, 4, 1, p, p, None, p, None) == -1          # gdim 4
    assert lib.gnx_refine_1d(p, p, 4, 3, 2, 25, p, p, None, p, None) == -1         # n too large
    assert lib.gnx_spmv_csr_f64(0, p, p, p, p, p, None) == -1                      # empty matrix
    assert lib.gnx_lincomb(10, 5, None, None, p, None) == -1                       # > 4 terms


def test_graph_validation():
    from graphnics_b200 import FenicsGraph
    G = FenicsGraph()
    G.add_nodes_from([1, 0])                   # labels not in insertion order
    G.nodes[0]["pos"] = [0, 0]; G.nodes[1]["pos"] = [1, 0]
    G.add_edge(0, 1)
    with pytest.raises(ValueError):
        G._graph_arrays()
    G = FenicsGraph()
    G.add_nodes_from([0, 1])
    G.nodes[0]["pos"] = [0, 0]; G.nodes[1]["pos"] = [1, 0]
    G.add_edge(0, 1); G.add_edge(1, 1)         # self loop
    with pytest.raises(ValueError):
        G._graph_arrays()
    G = FenicsGraph()
    assert isinstance(G, nx.DiGraph)


def test_generators_match_reference_golden_without_mesh():
    """node order, edge order and positions of the canned generators against what the reference's own
    generator code produced (tests/golden/graphs.json), without touching the GPU (mesh=False)"""
    from _util import golden_arrays
    from graphnics_b200 import line_graph, honeycomb, Y_bifurcation, YY_bifurcation
    for name, G in (("line_graph_3_dim3", line_graph(3, dim=3, mesh=False)), ("Y_bifurcation", Y_bifurcation(mesh=False)),
                    ("YY_bifurcation", YY_bifurcation(mesh=False)), ("honeycomb_4_4", honeycomb(4, 4, mesh=False)),
                    ("honeycomb_2_2", honeycomb(2, 2, mesh=False)), ("line_graph_2_dx2", line_graph(2, dx=2, mesh=False))):
        pos, edges = G._graph_arrays()
        gpos, gedges = golden_arrays(name)
        assert np.array_equal(edges, gedges), name
        assert np.array_equal(pos.view(np.int64), gpos.view(np.int64)), name


def test_make_arterial_tree_matches_reference_run():
    """data/generate_arterial_tree.py:34-188 (segment-intersection rejection included): edge lists equal
    the reference's own run (62 / 119 edges at N = 6 / 7, demo/Tree Profiling.ipynb:54-55), positions to
    rounding (the reference rotates with scipy's Rotation, we with the 2x2 matrix)."""
    from _util import golden_arrays
    from graphnics_b200 import make_arterial_tree
    counts = []
    for N in range(1, 8):
        signs = np.tile([-1, 1], 200).tolist()
        G = make_arterial_tree(N, directions=signs, gam=0.9, radius0=0.1, mesh=False)
        pos, edges = G._graph_arrays()
        gpos, gedges = golden_arrays(f"arterial_tree_N{N}")
        assert np.array_equal(edges, gedges)
        np.testing.assert_allclose(pos, gpos, rtol=0, atol=1e-12)
        counts.append(len(edges))
        assert len(signs) < 400 or N == 1          # `directions` is consumed, like the reference (:127)
    assert counts == [1, 3, 7, 15, 31, 62, 119]


def test_edge_attribute_cache_follows_every_way_of_writing_an_attribute():
    """edge_attribute_array() keeps its result while no edge attribute was written (the per-graph mutation stamp of the
    attribute dictionaries) and no Constant it found changed value; every networkx way of writing must invalidate it"""
    import networkx as nx
    import graphnics_b200 as gn
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": [float(i), 0.0]}) for i in range(5))
    G.add_edges_from([(0, 1), (1, 2), (1, 3), (3, 4)])
    c = gn.Constant(2.0)
    for e in G.edges():
        G.edges[e]["Res"] = c
    a = G.edge_attribute_array("Res", 1.0)
    assert a.tolist() == [2.0] * 4
    assert G.edge_attribute_array("Res", 1.0) is a                     # cached
    c.assign(3.0)                                                      # a Constant changed under the cache
    assert G.edge_attribute_array("Res", 1.0).tolist() == [3.0] * 4
    G.edges[1, 2]["Res"] = 5.0                                         # __setitem__
    assert G.edge_attribute_array("Res", 1.0).tolist() == [3.0, 5.0, 3.0, 3.0]
    nx.set_edge_attributes(G, {(0, 1): 7.0}, "Res")                    # networkx helper
    assert G.edge_attribute_array("Res", 1.0).tolist() == [7.0, 5.0, 3.0, 3.0]
    G[3][4].update(Res=9.0)                                            # dict.update through the adjacency view
    assert G.edge_attribute_array("Res", 1.0).tolist() == [7.0, 5.0, 3.0, 9.0]
    del G.edges[1, 3]["Res"]                                           # missing attribute -> default
    assert G.edge_attribute_array("Res", 1.0).tolist() == [7.0, 5.0, 1.0, 9.0]
    G.add_edge(2, 4, Res=11.0)                                         # structure change
    assert sorted(G.edge_attribute_array("Res", 1.0).tolist()) == [1.0, 5.0, 7.0, 9.0, 11.0]
    H = G.copy()                                                       # a copy has its own stamp
    assert isinstance(H, gn.FenicsGraph) and H._attr_stamp is not G._attr_stamp
    H.edges[0, 1]["Res"] = 13.0
    assert 13.0 in H.edge_attribute_array("Res", 1.0).tolist() and 13.0 not in G.edge_attribute_array("Res", 1.0).tolist()
    x = gn.Constant(1.0)                                               # a UFL expression of parameters is never cached
    u = 2.0 * x
    for e in G.edges():
        G.edges[e]["Ainv"] = u
    assert set(G.edge_attribute_array("Ainv", 1.0).tolist()) == {2.0}
    x.assign(4.0)
    assert set(G.edge_attribute_array("Ainv", 1.0).tolist()) == {8.0}


def test_edge_attribute_cache_survives_deepcopy_and_pickle():
    """a copied graph (copy.deepcopy, pickle, G.copy()) has its OWN mutation stamp, shared by its own attribute
    dictionaries: writing to the copy invalidates the copy's cache and only the copy's"""
    import copy
    import pickle
    import graphnics_b200 as gn
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": [float(i), 0.0]}) for i in range(3))
    G.add_edges_from([(0, 1), (1, 2)])
    G.edges[0, 1]["Res"] = 2.0
    assert G.edge_attribute_array("Res").tolist() == [2.0, 1.0]
    for H in (copy.deepcopy(G), pickle.loads(pickle.dumps(G)), G.copy()):
        assert isinstance(H, gn.FenicsGraph) and H._attr_stamp is not G._attr_stamp
        assert H.edges[0, 1]._stamp is H._attr_stamp
        assert H.edge_attribute_array("Res").tolist() == [2.0, 1.0]
        H.edges[0, 1]["Res"] = 5.0
        H.add_edge(2, 0, Res=7.0)
        assert H.edge_attribute_array("Res").tolist() == [5.0, 1.0, 7.0]
        assert G.edge_attribute_array("Res").tolist() == [2.0, 1.0]
