"""GPU parity tests of the primal model (HydraulicNetwork / TimeDepHydraulicNetwork) against the CPU
oracle, and the reference's own tests for it run through the drop-in API
(/root/reference/tests/test_flow_models.py:74-90,165-196, tests/test_time_stepping.py:17-44,63-95)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from _util import y_graph, random_graph, golden_arrays

pytestmark = pytest.mark.gpu

CASES = [("Y", 0), ("Y", 1), ("Y", 6), ("Y3", 4), ("rand0", 0), ("rand0", 3), ("rand1", 5), ("rand2", 7),
         ("honeycomb_2_2", 2), ("honeycomb_4_4", 5), ("YY_bifurcation", 5), ("arterial_tree_N7", 6),
         ("line_graph_3_dim3", 5), ("line_graph_2_dx2", 10), ("line_graph_2_dx2", 0)]


def graph(name):
    if name == "Y":
        return y_graph()
    if name == "Y3":
        return y_graph(3)
    if name.startswith("rand"):
        return random_graph(int(name[4:]), V=14 + 3 * int(name[4:]), extra=5)
    return golden_arrays(name)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def eng():
    import torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine_primal import PrimalSystem
    assert torch.cuda.is_available()
    return DeviceNetwork, PrimalSystem


class _C:
    """minimal CompiledCoefficient stand-in evaluated on the device through torch (test input only)"""

    def __init__(self, fn, degree):
        self.fn, self.degree, self.user, self.is_constant = fn, degree, None, False

    def eval_device(self, x, **kw):
        import torch
        return torch.from_numpy(self.fn(x.cpu().numpy())).to(x.device)


@pytest.mark.parametrize("name,n", CASES)
def test_primal_pattern_values_rhs_bc_solve(eng, name, n):
    import torch
    DeviceNetwork, PrimalSystem = eng
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    ps = PrimalSystem(dn)
    lay = oracle.primal_layout(om)
    assert ps.N == lay["N"]
    A = oracle.assemble_primal(om, 1.7)
    rp, ci = ps.build_pattern()
    assert np.array_equal(rp.cpu().numpy(), A.indptr)
    assert np.array_equal(ci.cpu().numpy(), A.indices)
    co = ps.coeffs(1.7)
    vals = ps.assemble(co)
    np.testing.assert_allclose(vals.cpu().numpy(), A.data, rtol=1e-12, atol=1e-300)
    # right-hand side (degree-2 data), Dirichlet elimination, solve against LU
    gf = lambda x: np.sin(x[..., 0]) + x[..., 1] ** 2
    ff = lambda x: np.cos(x[..., 1]) - x[..., 0]
    pf = lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1]
    b = oracle.assemble_primal_rhs(om, f=oracle.Coef(ff, 2), g=oracle.Coef(gf, 2))
    bg = ps.rhs(_C(ff, 2), _C(gf, 2))
    np.testing.assert_allclose(bg.cpu().numpy(), b, rtol=1e-12, atol=1e-14)
    dofs, bvals = oracle.primal_bc(om, oracle.Coef(pf, 1))
    Abc, bbc = oracle.apply_bc(A, b, dofs, bvals)
    g_node = torch.from_numpy(pf(pos)).cuda()
    vals_bc, b_bc = vals.clone(), bg.clone()
    from graphnics_b200._lib import check, ptr
    from graphnics_b200.device import current_stream
    check(ps.lib.gnx_apply_bc_primal(dn.net_ref(), 1.0, ptr(g_node), ptr(vals_bc), ptr(b_bc), current_stream()))
    np.testing.assert_allclose(vals_bc.cpu().numpy(), Abc.data, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(b_bc.cpu().numpy(), bbc, rtol=1e-12, atol=1e-13)
    x_ref = oracle.lu_solve(Abc, bbc)
    x, info = ps.solve_checked(co, vals_bc, b_bc)
    x = x.cpu().numpy()
    assert info.cg_converged
    assert rel(x[:lay["p_off"]], x_ref[:lay["p_off"]]) < 1e-8, info
    assert rel(x[lay["p_off"]:], x_ref[lay["p_off"]:]) < 1e-8, info
    assert info.rel_resid < 1e-10, info
    # Dirichlet values are reproduced exactly
    assert np.array_equal(x[dofs], bvals)


def test_primal_variable_resistance_and_theta_operator(eng):
    """per-cell Res (vasomotion-type coefficient) and D + theta dt A with a per-cell mass coefficient"""
    import torch
    DeviceNetwork, PrimalSystem = eng
    pos, edges = graph("rand1")
    n = 4
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    ps = PrimalSystem(dn)
    rng = np.random.default_rng(4)
    Rc = rng.uniform(0.5, 2.0, size=len(om.cells))           # global cell order
    Mc = rng.uniform(1.0, 3.0, size=len(om.cells))
    slot = ps.cell_slot().cpu().numpy()
    def local(a):
        out = np.empty_like(a)
        out[slot] = a
        return torch.from_numpy(out).cuda()
    s = 0.05
    h = np.linalg.norm(om.coords[om.cells[:, 1]] - om.coords[om.cells[:, 0]], axis=1)
    Dm = sp.diags(np.concatenate([Mc * h, np.zeros(len(om.coords))]))
    A0 = oracle.assemble_primal(om, Rc)
    Aeff = sp.csr_matrix((s * A0.data, A0.indices, A0.indptr), shape=A0.shape)   # keep the explicit p diagonal
    Aeff.data[3 * np.arange(len(om.cells))] += Mc * h                             # q rows: [diag, p, p]
    co = ps.coeffs(local(Mc + s * Rc), s)
    vals = ps.assemble(co)
    rp, ci = ps.build_pattern()
    Ag = sp.csr_matrix((vals.cpu().numpy(), ci.cpu().numpy(), rp.cpu().numpy()), shape=Aeff.shape)
    assert abs(Ag - Aeff).max() <= 1e-12 * abs(Aeff).max()
    b = rng.normal(size=ps.N)
    pf = lambda x: 0.3 * x[..., 0] - x[..., 2]
    dofs, bvals = oracle.primal_bc(om, oracle.Coef(pf, 1))
    Abc, bbc = oracle.apply_bc(Aeff, b, dofs, bvals)
    x_ref = oracle.lu_solve(Abc, bbc)
    bt = torch.from_numpy(b).cuda()
    ps.apply_bc_rhs(s, torch.from_numpy(pf(pos)).cuda(), bt)
    np.testing.assert_allclose(bt.cpu().numpy(), bbc, rtol=1e-12, atol=1e-13)
    x, info = ps.solve(co, bt)
    assert rel(x.cpu().numpy(), x_ref) < 1e-8, info
    # mass apply
    xx = rng.normal(size=ps.N)
    y = ps.mass_apply(ps.coeffs(local(Mc), 1.0), torch.from_numpy(xx).cuda()).cpu().numpy()
    np.testing.assert_allclose(y, Dm @ xx, rtol=1e-13, atol=1e-15)


def test_hydraulic_network():
    """tests/test_flow_models.py:74-90"""
    from graphnics_b200 import Y_bifurcation, HydraulicNetwork, Expression, near
    G = Y_bifurcation()
    G.make_mesh(6)
    model = HydraulicNetwork(G, p_bc=Expression('-x[1]', degree=2))
    q, p = model.solve()
    assert near(q(0, 0), q(0.5, 1) + q(-0.5, 1), 1e-3)
    assert near(p(0, 0), 0)
    assert near(p(0.5, 1), -1)
    # against the oracle
    pos, edges = y_graph()
    om = oracle.build_mesh(pos, edges, 6)
    _, _, xo = oracle.solve_primal(om, Res=1.0, p_bc=oracle.Coef(lambda x: -x[..., 1], 2))
    xg = np.concatenate([q.vector().get_local(), p.vector().get_local()])
    assert rel(xg, xo) < 1e-8


def hydraulic_manufactured_solution(G, Ainv, Res):
    """tests/test_flow_models.py:94-119"""
    from graphnics_b200 import SpatialCoordinate, sin, cos
    xx = SpatialCoordinate(G.mesh)
    q = sin(2 * 3.14 * xx[0])
    p = cos(2 * 3.14 * xx[0])
    g = Res * q + G.dds(p)
    f = G.dds(q)
    return f, q, p, g


def test_hydraulic():
    """tests/test_flow_models.py:165-196"""
    from graphnics_b200 import line_graph, HydraulicNetwork, FunctionSpace, project, errornorm
    Ainv = 0.0
    Res = 1
    G = line_graph(2, dx=2)
    G.make_mesh(10)
    f, q, p, g = hydraulic_manufactured_solution(G, Ainv, Res)
    p = project(p, FunctionSpace(G.mesh, "CG", 2))
    model = HydraulicNetwork(G, f=f, g=g, p_bc=p)
    sol = model.solve()
    qh, ph = sol
    pa = project(p, FunctionSpace(G.mesh, "CG", 2))
    p_error = errornorm(ph, pa)
    qa = project(q, FunctionSpace(G.mesh, "CG", 3))
    q_error = errornorm(qh, qa)
    assert q_error < 1e-2, f"Hydraulic model not giving correct flux, q_error = {q_error}"
    assert p_error < 1e-4, f"Hydraulic model not giving correct pressure, p_error = {p_error}"
    w = 2 * 3.14
    om = oracle.build_mesh(np.array([[0.0, 0.0], [2.0, 0.0]]), np.array([[0, 1]]), 10)
    co = lambda fn: oracle.Coef(fn, 2)
    _, _, xo = oracle.solve_primal(om, Res=1.0, f=co(lambda x: w * np.cos(w * x[..., 0])),
                                   g=co(lambda x: np.sin(w * x[..., 0]) - w * np.sin(w * x[..., 0])),
                                   p_bc=co(lambda x: np.cos(w * x[..., 0])))
    assert rel(sol.vector().get_local(), xo) < 1e-8


def time_manufactured_solution(G, Ainv, Res):
    """tests/test_time_stepping.py:17-44 (`ufl.diff` is exported as `diff`)"""
    from graphnics_b200 import Constant, SpatialCoordinate, sin, cos, diff
    t = Constant(0)
    xx = SpatialCoordinate(G.mesh)
    q = t * sin(2 * 3.14 * xx[0])
    p = t * cos(2 * 3.14 * xx[0])
    g = Ainv * diff(q, t) + Res * q + G.dds(p)
    f = G.dds(q)
    return q, p, f, g, t


def test_time_stepping_hydraulic():
    """tests/test_time_stepping.py:63-95"""
    from graphnics_b200 import (line_graph, TimeDepHydraulicNetwork, time_stepping_stokes, FunctionSpace, project,
                                errornorm)
    G = line_graph(2)
    G.make_mesh(10)
    Ainv = 3
    Res = 10
    T = 1
    t_steps = 100
    G.make_mesh(8)
    q, p, f, g, t = time_manufactured_solution(G, Ainv, Res)
    w = 2 * 3.14
    om = oracle.build_mesh(np.array([[0.0, 0.0], [1.0, 0.0]]), np.array([[0, 1]]), 8)
    TC = lambda fn: oracle.TimeCoef(fn, degree=2, clock="const")
    fo = TC(lambda x, tt: tt * w * np.cos(w * x[..., 0]))
    go = TC(lambda x, tt: Ainv * np.sin(w * x[..., 0]) + Res * tt * np.sin(w * x[..., 0]) - tt * w * np.sin(w * x[..., 0]))
    po = TC(lambda x, tt: tt * np.cos(w * x[..., 0]))
    for time_step_scheme in ['IE', 'CN']:
        t.assign(0)
        model = TimeDepHydraulicNetwork(G, f=f, g=g, p_bc=p, Res=Res, Ainv=Ainv)
        qps = time_stepping_stokes(model, t, t_steps=t_steps, T=T, t_step_scheme=time_step_scheme)
        qh, ph = qps[-1]
        t.assign(T)
        qa = project(q, FunctionSpace(G.mesh, "CG", 3))
        pa = project(p, FunctionSpace(G.mesh, "CG", 2))
        error_q = errornorm(qh, qa)
        error_p = errornorm(ph, pa)
        assert error_q < 1e-2, f'Large error in time stepping hydraulic model with {time_step_scheme},  error_q={error_q:1.1e}'
        assert error_p < 1.5e-2, f'Large error in time stepping hydraulic model with {time_step_scheme}, error_q={error_p:1.1e}'
        xs = oracle.time_stepping("primal", om, f=fo, g=go, p_bc=po, Res=float(Res), Ainv=float(Ainv),
                                  t_steps=t_steps, T=T, scheme=time_step_scheme)
        assert len(xs) == len(qps) == t_steps
        for k in (1, 2, 50, 99):
            assert rel(qps[k].vector().get_local(), xs[k]) < 1e-8, (time_step_scheme, k)
