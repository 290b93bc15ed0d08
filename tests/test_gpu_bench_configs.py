"""GPU parity on the BASELINE.json configurations themselves ("next round" item 1): the workloads
bench.py times are checked against the CPU oracle at the sizes they are benchmarked at (C2, C3) or at the
largest refinement the oracle finishes in under a minute (C4: 20 of the 1000 steps; C5: the 19-generation
tree with 2 cells per edge).  Mesh / pattern bit-exact, matrix entries and right-hand side 1e-12, every
solution field 1e-8 relative L2 -- through `MixedSystem.assemble_rhs_solve`, i.e. the one-call
`gnx_step_mixed` path that `MixedHydraulicNetwork.solve()` and bench.py run.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def mixed_step_parity(pos, edges, n, pcol, route="lu", check_matrix=True):
    """one bench step on (pos, edges, n) with Res = 1, p_bc = x[pcol], f = g = 0, against the oracle;
    returns the dict bench.py prints as `parity`"""
    import torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    assert np.array_equal(dn.cells.cpu().numpy(), om.cells)
    assert np.array_equal(dn.mf.cpu().numpy(), om.mf)
    assert np.array_equal(dn.coords.cpu().numpy().view(np.int64), om.coords.view(np.int64))
    assert dn.bifurcation_ixs() == om.bif_ixs and dn.boundary_ixs() == om.bnd_ixs
    ms = MixedSystem(dn)
    lay = oracle.mixed_layout(om)
    assert ms.N == lay["N"]
    rp, ci = ms.build_pattern()
    co = ms.coeffs(np.ones(dn.E))
    vals = torch.full((ms.nnz,), np.nan, dtype=torch.float64, device="cuda")
    b = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
    x = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
    xq, _ = ms.dof_coordinates(False)
    rhs_kw = dict(pbc_out=ms.end_values(xq[:, pcol].contiguous()), pbc_node=dn.pos[:, pcol].contiguous())
    ms.assemble_rhs_solve(co, vals, b, x, rhs_kw, overlap=True, sync=False)      # what bench.py times
    torch.cuda.synchronize()
    pbc = oracle.Coef(lambda xx: xx[..., pcol], 1)
    b_ref = oracle.assemble_mixed_rhs(om, p_bc=pbc, project=False)
    np.testing.assert_allclose(b.cpu().numpy(), b_ref, rtol=1e-12, atol=1e-14)
    out = {}
    if check_matrix:
        A = oracle.assemble_mixed(om, 1.0)
        out["pattern_equal"] = bool(np.array_equal(rp.cpu().numpy(), A.indptr) and np.array_equal(ci.cpu().numpy(), A.indices))
        assert out["pattern_equal"]
        np.testing.assert_allclose(vals.cpu().numpy(), A.data, rtol=1e-12, atol=1e-300)
    x_ref = oracle.lu_solve(A, b_ref) if route == "lu" else oracle.solve_by_elimination(om, b_ref, 1.0)
    xg = x.cpu().numpy()
    out["rel_l2_q"] = rel(xg[:lay["p_off"]], x_ref[:lay["p_off"]])
    out["rel_l2_p"] = rel(xg[lay["p_off"]:lay["lam_off"]], x_ref[lay["p_off"]:lay["lam_off"]])
    out["rel_l2_lam"] = rel(xg[lay["lam_off"]:], x_ref[lay["lam_off"]:]) if om.B else 0.0
    assert max(out["rel_l2_q"], out["rel_l2_p"], out["rel_l2_lam"]) < 1e-8, out
    return out


def test_c2_benchmark_tree_exactly_as_benched():
    """BASELINE.json configs[1] as bench.py builds it: murray_tree(11), 2^9 cells per edge (N = 2 099 198),
    Res = 1, p_bc = x[1] (demo/Tree Profiling.ipynb:60-102), against the oracle's SuperLU solve"""
    import bench
    pos, edges = bench.tree_arrays(bench.TREE_LEVELS)
    assert len(edges) == 2047
    out = mixed_step_parity(pos, edges, bench.TREE_N, 1)
    assert out["pattern_equal"]


def test_c5_shaped_tree_19_generations():
    """BASELINE.json configs[4]'s graph (19 generations: E = 524 287, B = 262 143 -- the blocked rake with 256
    subtree blocks and a 255-node top part) at 2 cells per edge, oracle by the elimination route"""
    from graphnics_b200.synthetic import murray_tree
    pos, edges, _, _ = murray_tree(19, radius0=0.1, gam=0.9, L0=10.0, gdim=2)
    mixed_step_parity(pos, edges, 1, 1, route="elimination")


@pytest.fixture(scope="module")
def honeycomb200():
    from graphnics_b200.generators import honeycomb
    G = honeycomb(200, 200)                    # make_mesh(1), like the reference generator (generators.py:84)
    return G


def test_c3_honeycomb_200_mixed(honeycomb200):
    """BASELINE.json configs[2], MixedHydraulicNetwork: honeycomb(200,200), n = 1 (N = 684 805), p_bc = x[0]
    (tests/test_flow_models.py:41): the 80 800-row cyclic core goes through the multilevel PCG"""
    pos, edges = honeycomb200._graph_arrays()
    assert len(edges) == 120801
    mixed_step_parity(pos, edges, 1, 0)


def test_c3_honeycomb_200_primal_through_the_api(honeycomb200):
    """BASELINE.json configs[2], HydraulicNetwork(G, p_bc=Expression('x[0]')).solve() through the drop-in API"""
    import graphnics_b200 as gn
    G = honeycomb200
    model = gn.HydraulicNetwork(G, p_bc=gn.Expression("x[0]", degree=2))
    q, p = model.solve()
    pos, edges = G._graph_arrays()
    om = oracle.build_mesh(pos, edges, 1)
    _, _, xo = oracle.solve_primal(om, Res=1.0, p_bc=oracle.Coef(lambda x: x[..., 0], 2))
    lay = oracle.primal_layout(om)
    assert rel(q.vector().get_local(), xo[:lay["p_off"]]) < 1e-8
    assert rel(p.vector().get_local(), xo[lay["p_off"]:]) < 1e-8


def test_c4_pulsatile_tumour_graph_20_steps():
    """BASELINE.json configs[3]: TimeDepHydraulicNetwork on the 19 874-edge tumour-like graph, implicit Euler
    with the time step of the 1000-step run (dt = 1e-3), operator reused; the first 20 states against the
    oracle's literal replay of time_stepping_stokes (graphnics/flowmodels/time_stepping.py:126-221)"""
    import graphnics_b200 as gn
    from graphnics_b200.synthetic import tumour_like_graph
    pos, edges, _ = tumour_like_graph()
    n, steps, T = 4, 20, 0.02
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
    G.add_edges_from(map(tuple, edges.tolist()))
    assert np.array_equal(G._graph_arrays()[1], edges)
    G.make_mesh(n)
    t = gn.Constant(0)
    f = gn.sin(2 * np.pi * 50 * t) * (1.0 + 0.0 * gn.SpatialCoordinate(G.mesh)[0])       # pulsatile source
    model = gn.TimeDepHydraulicNetwork(G, f=f, p_bc=gn.Constant(0), Res=gn.Constant(1.0), Ainv=gn.Constant(1.0))
    sols = gn.time_stepping_stokes(model, t=t, t_steps=steps, T=T, reassemble_lhs=False)
    assert len(sols) == steps
    om = oracle.build_mesh(pos, edges, n)
    fo = oracle.TimeCoef(lambda x, tt: np.sin(2 * np.pi * 50 * tt) * (1.0 + 0.0 * x[..., 0]), degree=2, clock="const")
    xs = oracle.time_stepping("primal", om, f=fo, g=0.0, p_bc=0.0, Res=1.0, Ainv=1.0, t_steps=steps, T=T, scheme="IE")
    assert len(xs) == steps
    lay = oracle.primal_layout(om)
    for k in range(steps):
        xg = sols[k].vector().get_local()
        for lo, hi in ((0, lay["p_off"]), (lay["p_off"], lay["N"])):
            ref = xs[k][lo:hi]
            assert np.linalg.norm(xg[lo:hi] - ref) <= 1e-8 * max(np.linalg.norm(ref), 1e-30), (k, lo)
