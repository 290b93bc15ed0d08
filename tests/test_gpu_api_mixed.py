"""The reference's own tests for the mixed model, run UNCHANGED IN SHAPE through the drop-in API
(`from graphnics_b200 import *` instead of `from graphnics import *`), plus comparison of what the
API produces with the CPU oracle.

Mirrors /root/reference/tests/test_flow_models.py:17-71 (mass conservation), :122-161
(manufactured solution, mixed) and tests/test_time_stepping.py:47-60,98-136 (time stepping, mixed).
"""
import networkx as nx
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def test_mass_conservation():
    """tests/test_flow_models.py:17-71"""
    from graphnics_b200 import (line_graph, Y_bifurcation, YY_bifurcation, honeycomb, Constant, Expression,
                                MixedHydraulicNetwork, ii_assemble, ii_convert, ii_Function, LUSolver,
                                BIF_IN, BIF_OUT, near)
    tests = {
        "Graph line": line_graph(3, dim=3),
        "Y bifurcation": Y_bifurcation(),
        "YY bifurcation": YY_bifurcation(),
        "honeycomb": honeycomb(4, 4),
    }
    for test_name in tests:
        G = tests[test_name]
        G.make_mesh(5)
        G.make_submeshes()
        prop_dict = {key: {"Res": Constant(1), "Ainv": Constant(1)} for key in list(G.edges.keys())}
        nx.set_edge_attributes(G, prop_dict)
        model = MixedHydraulicNetwork(G, p_bc=Expression("x[0]", degree=2))
        W = model.W
        a = model.a_form()
        L = model.L_form()
        A, b = map(ii_assemble, (a, L))
        A, b = map(ii_convert, (A, b))
        qp = ii_Function(W)
        solver = LUSolver(A, "mumps")
        solver.solve(qp.vector(), b)
        edge_list = list(G.edges.keys())
        for bif in G.bifurcation_ixs:
            total_flux = 0
            conn_edges = {
                **{e: (1, BIF_IN, edge_list.index(e)) for e in G.in_edges(bif)},
                **{e: (-1, BIF_OUT, edge_list.index(e)) for e in G.out_edges(bif)},
            }
            for e in conn_edges:
                sign, tag, e_ix = conn_edges[e]
                total_flux += sign * qp[e_ix](G.nodes[bif]["pos"])
            assert near(total_flux, 1e-3, 2e-3), f"Mass is not conserved at bifurcation {bif} for {test_name}"
            assert abs(total_flux) < 1e-10

        # the same system from the oracle (matrix 1e-12, rhs 1e-12, solution 1e-8)
        pos = np.asarray([G.nodes[v]["pos"] for v in G.nodes()], dtype=np.float64)
        edges = np.asarray(list(G.edges()), dtype=np.int64)
        om = oracle.build_mesh(pos, edges, 5)
        Ao = oracle.assemble_mixed(om, 1.0)
        bo = oracle.assemble_mixed_rhs(om, p_bc=oracle.Coef(lambda x: x[..., 0], 2), project=True)
        Ag = A.tocsr()
        assert np.array_equal(Ag.indptr, Ao.indptr) and np.array_equal(Ag.indices, Ao.indices)
        np.testing.assert_allclose(Ag.data, Ao.data, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(b.get_local(), bo, rtol=1e-11, atol=1e-13)
        xo = oracle.lu_solve(Ao, bo)
        xg = qp.vector().get_local()
        assert np.linalg.norm(xg - xo) <= 1e-8 * np.linalg.norm(xo)


def hydraulic_manufactured_solution(G, Ainv, Res):
    """tests/test_flow_models.py:94-119"""
    from graphnics_b200 import SpatialCoordinate, sin, cos
    xx = SpatialCoordinate(G.mesh)
    q = sin(2 * 3.14 * xx[0])
    p = cos(2 * 3.14 * xx[0])
    g = Res * q + G.dds(p)
    f = G.dds(q)
    return f, q, p, g


def test_mixed_hydraulic():
    """tests/test_flow_models.py:122-161"""
    from graphnics_b200 import (line_graph, Constant, MixedHydraulicNetwork, FunctionSpace, project, errornorm)
    Ainv = 0.0
    Res = 1
    G = line_graph(2, dx=2)
    G.make_mesh(8)
    G.make_submeshes()
    f, q, p, g = hydraulic_manufactured_solution(G, Ainv, Res)
    p = project(p, FunctionSpace(G.mesh, "CG", 2))
    prop_dict = {key: {"Res": Constant(Res), "Ainv": Constant(Ainv)} for key in list(G.edges.keys())}
    nx.set_edge_attributes(G, prop_dict)
    model = MixedHydraulicNetwork(G, f=f, g=g, p_bc=p)
    sol = model.solve()
    qh, ph = sol
    pa = project(p, FunctionSpace(G.mesh, "CG", 2))
    p_error = errornorm(ph, pa)
    qa = project(q, FunctionSpace(G.mesh, "CG", 3))
    q_error = errornorm(qh, qa)
    assert q_error < 1e-2, f"Mixed hydraulic model not giving correct flux, q_error = {q_error}"
    assert p_error < 1e-1, f"Mixed hydraulic model not giving correct pressure, p_error = {p_error}"

    # against the oracle with the same (projected, degree 2) data
    pos = np.array([[0.0, 0.0], [2.0, 0.0]])
    edges = np.array([[0, 1]])
    om = oracle.build_mesh(pos, edges, 8)
    w = 2 * 3.14
    co = lambda fn: oracle.Coef(fn, 2)
    _, _, xo = oracle.solve_mixed(om, Res=1.0, f=co(lambda x: w * np.cos(w * x[..., 0])),
                            g=co(lambda x: np.sin(w * x[..., 0]) - w * np.sin(w * x[..., 0])),
                            p_bc=co(lambda x: np.cos(w * x[..., 0])), project=True)
    xg = sol.vector().get_local()
    assert np.linalg.norm(xg - xo) <= 1e-8 * np.linalg.norm(xo), np.linalg.norm(xg - xo) / np.linalg.norm(xo)


def sympy_hydraulic_manufactured_solution(G, Ainv, Res):
    """tests/test_time_stepping.py:47-60: q = t sin(2*3.14 x), p = cos(2*3.14 x) (p does not depend on t),
    g = Ainv q_t + Res q + p_x, f = q_x, all as `Expression(ccode, degree=2, t=0)`"""
    import sympy as sym
    from graphnics_b200 import Expression
    x, t = sym.symbols("x[0] t")
    q = t * sym.sin(2 * 3.14 * x)
    p = sym.cos(2 * 3.14 * x)
    g = Ainv * sym.diff(q, t) + Res * q + sym.diff(p, x)
    f = sym.diff(q, x)
    q, p, f, g = [Expression(sym.printing.ccode(expr), degree=2, t=0) for expr in (q, p, f, g)]
    return q, p, f, g


@pytest.mark.parametrize("scheme", ["IE", "CN"])
def test_time_stepping_mixed_hydraulic(scheme):
    """tests/test_time_stepping.py:98-136 with the reference's own data and bounds (error_q < 1e-1,
    error_p < 1e-1; the reference only runs IE: "TODO: Test with CN"), then every stored state against the
    oracle's literal replay of the loop"""
    from graphnics_b200 import (line_graph, Constant, TimeDepMixedHydraulicNetwork, time_stepping_stokes,
                                FunctionSpace, project, errornorm)
    G = line_graph(2)
    G.make_mesh(5)
    G.make_submeshes()
    Ainv = 2
    Res = 10
    for e in G.edges():
        G.edges()[e]["Ainv"] = Ainv
        G.edges()[e]["Res"] = Res
    T = 1
    t_steps = 10
    q, p, f, g = sympy_hydraulic_manufactured_solution(G, Ainv, Res)
    t = Constant(0)
    t.assign(0)
    model = TimeDepMixedHydraulicNetwork(G, f=f, g=g, p_bc=p, Ainv=Ainv)
    qps = time_stepping_stokes(model, t, t_steps=t_steps, T=T, t_step_scheme=scheme)
    assert len(qps) == t_steps
    qh, ph = qps[-1]
    q.t = T
    p.t = T
    qa = project(q, FunctionSpace(G.mesh, "CG", 3))
    pa = project(p, FunctionSpace(G.mesh, "CG", 2))
    error_q = errornorm(qh, qa)
    error_p = errornorm(ph, pa)
    if scheme == "IE":                 # the reference's assertions, unchanged
        assert error_q < 1e-1, f"Large error in time stepping mixed hydraulic model, error_q={error_q:1.1e}"
        assert error_p < 1e-1, f"Large error in time stepping mixed hydraulic model, error_p={error_p:1.1e}"
    else:                              # CN: the flux is as good; the pressure (a multiplier) oscillates
        assert error_q < 1e-1

    # every stored state against the oracle's literal restatement of the loop (quirks included)
    w = 2 * 3.14
    om = oracle.build_mesh(np.array([[0.0, 0.0], [1.0, 0.0]]), np.array([[0, 1]]), 5)
    TC = lambda fn: oracle.TimeCoef(fn, degree=2, clock="attr")
    fo = TC(lambda x, tt: tt * w * np.cos(w * x[..., 0]))
    go = TC(lambda x, tt: Ainv * np.sin(w * x[..., 0]) + Res * tt * np.sin(w * x[..., 0]) - w * np.sin(w * x[..., 0]))
    po = TC(lambda x, tt: np.cos(w * x[..., 0]))
    xs = oracle.time_stepping("mixed", om, f=fo, g=go, p_bc=po, Res=float(Res), Ainv=float(Ainv),
                              t_steps=t_steps, T=T, scheme=scheme, project=True)
    assert len(xs) == len(qps)
    for k, (xo, qpk) in enumerate(zip(xs, qps)):
        xg = qpk.vector().get_local()
        assert np.linalg.norm(xg - xo) <= 1e-8 * max(np.linalg.norm(xo), 1e-30), (scheme, k)


def test_time_stepping_initial_state_given_as_a_list_of_components():
    """qp0 as the reference takes it (time_stepping.py:162-165): one entry per space of W, projected component-wise"""
    from graphnics_b200 import (line_graph, Constant, Expression, TimeDepMixedHydraulicNetwork, TimeDepHydraulicNetwork,
                                time_stepping_stokes)
    G = line_graph(3)
    G.make_mesh(4)
    G.make_submeshes()
    Ainv, Res, T, t_steps = 2.0, 3.0, 0.5, 5
    for e in G.edges():
        G.edges()[e]["Res"] = Res
    q0 = Expression("sin(x[0])", degree=2)
    p0 = Expression("1 + x[0]", degree=1)
    model = TimeDepMixedHydraulicNetwork(G, p_bc=Expression("x[0]", degree=2), Ainv=Ainv)
    E, B = 2, 1
    qps = time_stepping_stokes(model, Constant(0), t_steps=t_steps, T=T, qp0=[q0] * E + [p0] + [Constant(0.25)] * B)
    pos = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    om = oracle.build_mesh(pos, np.array([[0, 1], [1, 2]]), 4)
    lay = oracle.mixed_layout(om)
    x0 = np.zeros(lay["N"])
    x0[:lay["p_off"]] = oracle.project_p1_edges(om, oracle.Coef(lambda x: np.sin(x[..., 0]), 2)).ravel()
    mid = om.coords[om.cells].mean(axis=1)
    x0[lay["p_off"]:lay["lam_off"]] = 1.0 + mid[:, 0]                 # cell means of a linear function
    x0[lay["lam_off"]:] = 0.25
    np.testing.assert_allclose(qps[0].vector().get_local(), x0, rtol=1e-12, atol=1e-14)
    xs = oracle.time_stepping("mixed", om, p_bc=oracle.Coef(lambda x: x[..., 0], 2), Res=Res, Ainv=Ainv, t_steps=t_steps,
                              T=T, x0=x0, scheme="IE")
    for k in range(t_steps):
        xg = qps[k].vector().get_local()
        assert np.linalg.norm(xg - xs[k]) <= 1e-8 * np.linalg.norm(xs[k]), k
    # primal model: [DG0 flux | CG1 pressure]
    G.make_mesh(4)
    modelp = TimeDepHydraulicNetwork(G, p_bc=Expression("x[0]", degree=2), Res=Constant(Res), Ainv=Constant(Ainv))
    qps = time_stepping_stokes(modelp, Constant(0), t_steps=t_steps, T=T, qp0=[q0, p0], t_step_scheme="CN")
    layp = oracle.primal_layout(om)
    x0 = np.zeros(layp["N"])
    xa, xb = om.coords[om.cells[:, 0]][:, 0], om.coords[om.cells[:, 1]][:, 0]
    x0[:layp["p_off"]] = (np.sin(xa) + 4 * np.sin(0.5 * (xa + xb)) + np.sin(xb)) / 6.0
    x0[layp["p_off"]:] = 1.0 + om.coords[:, 0]
    np.testing.assert_allclose(qps[0].vector().get_local(), x0, rtol=1e-12, atol=1e-14)
    xs = oracle.time_stepping("primal", om, p_bc=oracle.Coef(lambda x: x[..., 0], 2), Res=Res, Ainv=Ainv, t_steps=t_steps,
                              T=T, x0=x0, scheme="CN")
    for k in range(t_steps):
        xg = qps[k].vector().get_local()
        assert np.linalg.norm(xg - xs[k]) <= 1e-8 * np.linalg.norm(xs[k]), k


def test_zero_shape_form_and_real_projections():
    """init_a_form() assembles to the pattern with zero values (flow_models.py:266-280); project() onto the
    lowest-order spaces really projects (DG0: cell means, per-edge CG1: consistent-mass L2 projection)"""
    from graphnics_b200 import (Y_bifurcation, MixedHydraulicNetwork, Expression, FunctionSpace, project, ii_assemble)
    G = Y_bifurcation()
    G.make_mesh(4)
    G.make_submeshes()
    model = MixedHydraulicNetwork(G, p_bc=Expression("x[1]", degree=2))
    Z = ii_assemble(model.init_a_form()).tocsr()
    A = ii_assemble(model.a_form(model.init_a_form())).tocsr()
    assert np.array_equal(Z.indptr, A.indptr) and np.array_equal(Z.indices, A.indices) and not Z.data.any() and A.data.any()
    with pytest.raises(NotImplementedError):
        model.a_form(ii_assemble(model.a_form()))
    pos = np.array([[0, 0], [0, .5], [-.5, 1], [.5, 1]], dtype=np.float64)
    om = oracle.build_mesh(pos, np.array([[0, 1], [1, 2], [1, 3]]), 4)
    f = Expression("sin(3*x[0]) + x[1]*x[1]", degree=2)
    fo = oracle.Coef(lambda x: np.sin(3 * x[..., 0]) + x[..., 1] ** 2, 2)
    p0 = project(f, FunctionSpace(G.mesh, "DG", 0)).vector().get_local()
    xa, xb = om.coords[om.cells[:, 0]], om.coords[om.cells[:, 1]]
    np.testing.assert_allclose(p0, (fo(xa) + 4 * fo(0.5 * (xa + xb)) + fo(xb)) / 6.0, rtol=1e-13)
    ref = oracle.project_p1_edges(om, fo)
    for i in range(3):
        qi = project(f, model.W[i]).vector().get_local()
        np.testing.assert_allclose(qi, ref[i], rtol=1e-11, atol=1e-13)


def test_resistance_given_as_a_function_of_position():
    """edge attribute "Res" as an Expression / UFL of x (flow_models.py:228-233, demo/Vasomotion in arterial
    trees.ipynb:222-289): sampled at the cell midpoints; matrix entries 1e-12 and solution 1e-8 against the oracle with
    the same cell-wise resistances, through model.solve() and through the explicit assemble / LUSolver idiom"""
    from graphnics_b200 import (Y_bifurcation, MixedHydraulicNetwork, Expression, Constant, SpatialCoordinate, ii_assemble,
                                ii_convert, ii_Function, LUSolver)
    n = 5
    G = Y_bifurcation()
    G.make_mesh(n)
    G.make_submeshes()
    r0 = Expression("1 + x[0]*x[0] + 0.5*x[1]", degree=2)
    xx = SpatialCoordinate(G.mesh)
    G.edges[0, 1]["Res"] = r0
    G.edges[1, 2]["Res"] = 2.0 + xx[1] * xx[1]          # UFL
    G.edges[1, 3]["Res"] = Constant(3.0)
    model = MixedHydraulicNetwork(G, f=Constant(0.3), g=Constant(-0.2), p_bc=Expression("x[1]", degree=2))
    qp = model.solve()
    pos = np.array([[0, 0], [0, .5], [-.5, 1], [.5, 1]], dtype=np.float64)
    om = oracle.build_mesh(pos, np.array([[0, 1], [1, 2], [1, 3]]), n)
    mid = 0.5 * (om.sub_coords[np.arange(3)[:, None], om.sub_cells[:, :, 0]] + om.sub_coords[np.arange(3)[:, None], om.sub_cells[:, :, 1]])
    R = np.stack([1 + mid[0, :, 0] ** 2 + 0.5 * mid[0, :, 1], 2.0 + mid[1, :, 1] ** 2, np.full(om.m, 3.0)])
    A = oracle.assemble_mixed(om, R)
    b = oracle.assemble_mixed_rhs(om, f=0.3, g=-0.2, p_bc=oracle.Coef(lambda x: x[..., 1], 2), project=True)
    x_ref = oracle.lu_solve(A, b)
    Ag = model.A.tocsr()
    assert np.array_equal(Ag.indptr, A.indptr) and np.array_equal(Ag.indices, A.indices)
    np.testing.assert_allclose(Ag.data, A.data, rtol=1e-12, atol=1e-300)
    xg = qp.vector().get_local()
    assert np.linalg.norm(xg - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    Am, bm = map(ii_assemble, (model.a_form(), model.L_form()))
    Am, bm = map(ii_convert, (Am, bm))
    qp2 = ii_Function(model.W)
    LUSolver(Am, "mumps").solve(qp2.vector(), bm)
    assert np.linalg.norm(qp2.vector().get_local() - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    # a time-dependent resistance: the operator follows `t` when it is re-assembled
    r0.t = 0.0
    rt = Expression("1 + t + x[0]*x[0]", degree=2, t=0.0)
    for e in G.edges():
        G.edges[e]["Res"] = rt
    x0 = MixedHydraulicNetwork(G, p_bc=Expression("x[1]", degree=2)).solve().vector().get_local()
    rt.t = 2.0
    x1 = MixedHydraulicNetwork(G, p_bc=Expression("x[1]", degree=2)).solve().vector().get_local()
    Rt = lambda t: 1 + t + mid[:, :, 0] ** 2
    b0 = oracle.assemble_mixed_rhs(om, p_bc=oracle.Coef(lambda x: x[..., 1], 2), project=True)
    for xg, t in ((x0, 0.0), (x1, 2.0)):
        xr = oracle.lu_solve(oracle.assemble_mixed(om, Rt(t)), b0)
        assert np.linalg.norm(xg - xr) <= 1e-8 * np.linalg.norm(xr)
