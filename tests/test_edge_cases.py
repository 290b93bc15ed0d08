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
