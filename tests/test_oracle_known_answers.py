"""CPU-only: the oracle against every known answer the reference (and the survey of it) holds for the
path -- SURVEY.md Appendix B (numbering worked example replayed from DOLFIN's rules A1-A5), Appendix C
(Ohm/Kirchhoff closed form; tree generator edge counts), the reference tests' physics properties
(tests/test_flow_models.py:17-90, tests/test_fenics_graph.py:29-35), and the oracle's two independent
solution routes against each other."""
import numpy as np
import pytest

import oracle
from _util import y_graph, golden_arrays, golden_graphs, fromhex


def test_appendix_b_numbering_worked_example():
    pos, edges = y_graph()
    om = oracle.build_mesh(pos, edges, 2)
    cells = [(0, 7), (4, 7), (1, 8), (4, 8), (1, 9), (5, 9), (2, 10), (5, 10), (1, 11), (6, 11), (3, 12), (6, 12)]
    assert [tuple(c) for c in om.cells.tolist()] == cells
    assert om.mf.tolist() == [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2]
    xy = {4: (0, .25), 7: (0, .125), 8: (0, .375), 9: (-.125, .625), 5: (-.25, .75), 10: (-.375, .875),
          11: (.125, .625), 6: (.25, .75), 12: (.375, .875)}
    for v, c in xy.items():
        assert tuple(om.coords[v]) == c
    assert om.parent_vertex.tolist() == [[0, 7, 4, 1, 8], [1, 9, 5, 2, 10], [1, 11, 6, 3, 12]]
    for i in range(3):
        assert om.sub_cells[i].tolist() == [[0, 1], [1, 2], [3, 4], [2, 4]]
    # vertex tags (fenics_graph.py:132-156): node 0 boundary (edge goes OUT: BOUN_IN=3), node 1 bifurcation
    # (edge 0 IN: BIF_IN=1; edges 1, 2 OUT: BIF_OUT=2), nodes 2, 3 boundary (edge goes IN: BOUN_OUT=4)
    assert om.vf.tolist() == [[3, 0, 0, 1, 0], [2, 0, 0, 4, 0], [2, 0, 0, 4, 0]]


def test_appendix_c_closed_form_and_symmetry():
    """Y, n=6, Res=[1,2,3], p_bc = 2 - x[1]: N=388, nnz(structural)=1353, q per edge, lambda, symmetry"""
    pos, edges = y_graph()
    om = oracle.build_mesh(pos, edges, 6)
    A, b, x = oracle.solve_mixed(om, Res=[1.0, 2.0, 3.0], p_bc=oracle.Coef(lambda x: 2.0 - x[..., 1], 2))
    lay = oracle.mixed_layout(om)
    assert A.shape[0] == 388
    assert (A.data != 0).sum() == 1353                     # E(7m+1) + 2I
    assert A.nnz == 1353 + 3 * 64 + 1                      # + explicit-zero p / lambda diagonals (init_a_form)
    q = x[:lay["p_off"]].reshape(3, 65)
    for i, val in enumerate((0.741549228561398, 0.444929537136839, 0.296619691424559)):
        assert np.allclose(q[i], val, rtol=1e-12)
    assert np.isclose(x[lay["lam_off"]], 1.629225385719301, rtol=1e-12)
    assert abs(q[0, 0] - q[1, 0] - q[2, 0]) < 1e-14       # Kirchhoff
    S = A.copy().tolil()
    S[lay["p_off"]:lay["lam_off"]] = -S[lay["p_off"]:lay["lam_off"]]
    S = S.tocsr()
    assert abs(S - S.T).max() < 1e-15                      # negating the p rows symmetrises the operator


@pytest.mark.parametrize("name", ["line_graph_3_dim3", "Y_bifurcation", "YY_bifurcation", "honeycomb_4_4"])
def test_reference_mass_conservation_property(name):
    """tests/test_flow_models.py:17-71 on the oracle: net flux at every bifurcation vanishes"""
    pos, edges = golden_arrays(name)
    om = oracle.build_mesh(pos, edges, 5)
    _, _, x = oracle.solve_mixed(om, Res=1.0, p_bc=oracle.Coef(lambda x: x[..., 0], 2))
    q, _, _ = oracle.mixed_components(om, x)
    for b in om.bif_ixs:
        tot = 0.0
        for i, (u, v) in enumerate(edges):
            if v == b:
                tot += oracle.eval_q_edge(om, q, i, pos[b])
            if u == b:
                tot -= oracle.eval_q_edge(om, q, i, pos[b])
        assert abs(tot) < 1e-10


def test_reference_hydraulic_network_property():
    """tests/test_flow_models.py:74-90 on the oracle: Dirichlet values exact, flux balance"""
    pos, edges = y_graph()
    om = oracle.build_mesh(pos, edges, 6)
    _, _, x = oracle.solve_primal(om, Res=1.0, p_bc=oracle.Coef(lambda x: -x[..., 1], 2))
    lay = oracle.primal_layout(om)
    p = x[lay["p_off"]:]
    assert p[0] == 0.0 and p[3] == -1.0 and p[2] == -1.0
    q = x[:lay["p_off"]].reshape(3, 64)
    assert abs(q[0].mean() - q[1].mean() - q[2].mean()) < 1e-12


def test_vertex_degrees_golden():
    """tests/test_fenics_graph.py:29-35: degree[0] = 0.25, degree[1] ~ 0.9571 (reference's own run)"""
    c = golden_graphs()["Y_bifurcation"]
    deg = [float.fromhex(h) for h in c["degree"]]
    assert deg[0] == 0.25 and abs(deg[1] - 0.9571) < 0.01
    pos, edges = fromhex(c["pos"]), np.asarray(c["edges"])
    mine = oracle.vertex_degrees(len(pos), pos, edges)
    assert np.array_equal(np.asarray(mine).view(np.int64), np.asarray(deg).view(np.int64))


@pytest.mark.parametrize("name,n", [("honeycomb_4_4", 3), ("YY_bifurcation", 5), ("arterial_tree_N6", 2),
                                    ("line_graph_2_dx2", 6)])
def test_lu_route_and_elimination_route_agree(name, n):
    pos, edges = golden_arrays(name)
    om = oracle.build_mesh(pos, edges, n)
    rng = np.random.default_rng(4)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    A = oracle.assemble_mixed(om, Res)
    b = oracle.assemble_mixed_rhs(om, f=oracle.Coef(lambda x: np.cos(x[..., 0]), 2), g=oracle.Coef(lambda x: x[..., 1], 2),
                                  p_bc=oracle.Coef(lambda x: 1 + x[..., 0], 2), project=True)
    x1 = oracle.lu_solve(A, b)
    x2 = oracle.solve_by_elimination(om, b, Res)
    assert np.linalg.norm(x1 - x2) <= 1e-11 * np.linalg.norm(x1)
