"""CPU-only: the closed-form numbering / row math that the CUDA kernels execute
(graphnics_b200/csrc/gnx_index.h, compiled for the host by tests/hostcheck) against the oracle's
literal level-by-level restatement.  Bit-exact for integers and coordinates."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from _util import HostNet, y_graph, random_graph, golden_arrays

CASES = [("Y", 0), ("Y", 1), ("Y", 2), ("Y", 5), ("Y3", 3), ("rand0", 0), ("rand0", 3),
         ("rand1", 4), ("honeycomb_2_2", 2), ("arterial_tree_N5", 3), ("line_graph_3_dim3", 4)]


def graph(name):
    if name == "Y":
        return y_graph()
    if name == "Y3":
        return y_graph(3)
    if name.startswith("rand"):
        return random_graph(int(name[4:]))
    return golden_arrays(name)


@pytest.mark.parametrize("name,n", CASES)
def test_mesh_bit_exact(name, n):
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    hn = HostNet(pos, edges, n)
    assert np.array_equal(hn.cells, om.cells)
    assert np.array_equal(hn.mf, om.mf)
    assert np.array_equal(hn.coords.view(np.int64), om.coords.view(np.int64)), "coordinates not bit-exact"
    assert np.array_equal(hn.tangents.view(np.int64), om.tangents.view(np.int64)), "tangents not bit-exact"


@pytest.mark.parametrize("name,n", CASES)
def test_submesh_numbering(name, n):
    from graphnics_b200.device import submesh_numbering
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    loc, tloc = submesh_numbering(n)
    m = 1 << n
    # parent vertex of local vertex l on edge e == vertex id of position tloc[l]
    from _util import hostcheck_lib  # noqa: F401  (ensures the shim is built)
    V, E = len(pos), len(edges)
    for e in range(E):
        lo, hi = min(edges[e]), max(edges[e])
        for l in range(m + 1):
            t = int(tloc[l])
            if t == 0:
                vid = lo
            elif t == m:
                vid = hi
            else:
                k = n - (t & -t).bit_length() + 1
                half = 1 << (k - 1)
                sprime = t >> (n - k + 1)
                vid = V + E * (half - 1) + e * half + (sprime ^ (sprime >> 1))
            assert vid == om.parent_vertex[e, l]


@pytest.mark.parametrize("name,n", CASES)
def test_pattern_and_values(name, n):
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    hn = HostNet(pos, edges, n)
    rng = np.random.default_rng(1)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    A = oracle.assemble_mixed(om, Res)
    rowptr, colidx = hn.pattern()
    assert hn.N == A.shape[0] and hn.nnz == A.nnz
    assert np.array_equal(rowptr, A.indptr)
    assert np.array_equal(colidx, A.indices)
    vals = hn.values(a_edge=Res)
    np.testing.assert_allclose(vals, A.data, rtol=1e-12, atol=1e-300)
    # NetworkStokes viscous term and a sigma-scaled operator
    visc = rng.uniform(0.1, 1.0, size=len(edges))
    A2 = oracle.assemble_mixed(om, Res, visc=visc)
    np.testing.assert_allclose(hn.values(a_edge=Res, k_edge=visc), A2.data, rtol=1e-11, atol=1e-300)
    # mass matrix on the same pattern
    D = oracle.assemble_mass_q_mixed(om, 2.5)
    Dg = sp.csr_matrix((hn.values(a_const=2.5, sigma=0.0), colidx, rowptr), shape=A.shape)
    assert abs(Dg - D).max() <= 1e-12 * abs(D).max()


@pytest.mark.parametrize("name,n", CASES)
def test_rhs(name, n):
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    hn = HostNet(pos, edges, n)
    lin = oracle.Coef(lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1], 1)
    g = oracle.Coef(lambda x: np.sin(x[..., 0]) + x[..., 1] ** 2, 1)
    f = oracle.Coef(lambda x: np.cos(x[..., 1]) - x[..., 0], 1)
    b = oracle.assemble_mixed_rhs(om, f=f, g=g, p_bc=lin, project=False)
    nod = lambda c: c(om.sub_coords).ravel()
    bh = hn.rhs(f=nod(f), g=nod(g), pbc_out=hn.end_values(nod(lin)), pbc_node=lin(pos))
    np.testing.assert_allclose(bh, b, rtol=1e-12, atol=1e-14)
    # Constant data path
    b2 = oracle.assemble_mixed_rhs(om, f=0.75, g=-1.25, p_bc=2.0)
    bh2 = hn.rhs(f_const=0.75, g_const=-1.25, pbc_out=np.full(len(edges), 2.0), pbc_node=np.full(len(pos), 2.0))
    np.testing.assert_allclose(bh2, b2, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name,n", CASES)
def test_primal_pattern_values_rhs(name, n):
    """HydraulicNetwork closed-form rows (gnx_primal.h) against the oracle's element-by-element assembly"""
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    hn = HostNet(pos, edges, n)
    rng = np.random.default_rng(2)
    A = oracle.assemble_primal(om, 1.7)
    rowptr, colidx = hn.primal_pattern()
    assert np.array_equal(rowptr, A.indptr)
    assert np.array_equal(colidx, A.indices)
    np.testing.assert_allclose(hn.primal_values(a_const=1.7), A.data, rtol=1e-12, atol=1e-300)
    # per-cell resistance (global cell order in the oracle, edge-local order in the kernels)
    Rc = rng.uniform(0.5, 2.0, size=len(om.cells))
    A2 = oracle.assemble_primal(om, Rc)
    a_cell = np.empty_like(Rc)
    a_cell[hn.cell_slot()] = Rc
    np.testing.assert_allclose(hn.primal_values(a_cell=a_cell), A2.data, rtol=1e-12, atol=1e-300)
    # mass matrix and a theta-scheme operator
    D = oracle.assemble_mass_q_primal(om, 2.5)
    Dg = sp.csr_matrix((hn.primal_values(a_const=2.5, sigma=0.0), colidx, rowptr), shape=A.shape)
    assert abs(Dg - D).max() <= 1e-12 * abs(D).max()
    # right-hand side, degree-2 and degree-1 data and constants
    xa, xb = om.coords[om.cells[:, 0]], om.coords[om.cells[:, 1]]
    pts = np.stack([xa, 0.5 * (xa + xb), xb], axis=1)
    for deg in (2, 1):
        g = oracle.Coef(lambda x: np.sin(x[..., 0]) + x[..., 1] ** 2, deg)
        f = oracle.Coef(lambda x: np.cos(x[..., 1]) - x[..., 0], deg)
        b = oracle.assemble_primal_rhs(om, f=f, g=g)
        bh = hn.primal_rhs(f_cell=f(pts).ravel(), f_lin=deg <= 1, g_cell=g(pts).ravel(), g_lin=deg <= 1)
        np.testing.assert_allclose(bh, b, rtol=1e-12, atol=1e-14)
    b2 = oracle.assemble_primal_rhs(om, f=0.75, g=-1.25)
    np.testing.assert_allclose(hn.primal_rhs(fc=0.75, gc=-1.25), b2, rtol=1e-12, atol=1e-14)


def test_vf_tags_match_oracle():
    from _util import host_tables
    from graphnics_b200.device import submesh_numbering
    for name, n in CASES:
        pos, edges = graph(name)
        om = oracle.build_mesh(pos, edges, n)
        t = host_tables(len(pos), edges)
        loc, _ = submesh_numbering(n)
        m = 1 << n
        vf = np.zeros_like(om.vf)
        flip = edges[:, 0] > edges[:, 1]
        ar = np.arange(len(edges))
        vf[ar, np.where(flip, loc[m], loc[0])] = t["tag_u"]
        vf[ar, np.where(flip, loc[0], loc[m])] = t["tag_v"]
        assert np.array_equal(vf, om.vf), name
        assert t["bif_nodes"].tolist() == om.bif_ixs and t["bnd_nodes"].tolist() == om.bnd_ixs


def test_rpn_interpreter_matches_sympy():
    """graphnics_b200/csrc/gnx_rpn.h (host build) against sympy on the expressions the reference's
    tests use (tests/test_flow_models.py:110-117, tests/test_time_stepping.py:36-56)."""
    import ctypes as C
    import sympy as sp
    from _util import hostcheck_lib
    from graphnics_b200.coefficients import (Expression, Constant, SpatialCoordinate, sin, cos, diff, dot, grad,
                                            as_coefficient, conditional, lt)
    hc = hostcheck_lib()
    t = Constant(0.3)
    xx = SpatialCoordinate(3)
    q = t * sin(2 * 3.14 * xx[0])
    p = t * cos(2 * 3.14 * xx[0])
    cases = [Expression("x[0]", degree=2), Expression("2-x[1]", degree=2), Expression("-x[1]", degree=2),
             Expression("t*sin(2*3.14*x[0]) + pow(x[1], 3) / (1 + x[2]*x[2])", degree=2, t=0.7),
             3.0 * diff(q, t) + 10 * q + dot(grad(p), None), dot(grad(q), None),
             conditional(lt(xx[0], 0.5), xx[1] ** 2, sp.sqrt(2) * xx[0])]
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(20, 3))
    tau = rng.normal(size=(20, 3))
    for c in cases:
        cc = as_coefficient(c)
        prog = cc.program()
        ref = cc.eval_host(pts, tangents=tau)
        for i in range(len(pts)):
            x = np.ascontiguousarray(pts[i]); ta = np.ascontiguousarray(tau[i])
            v = hc.hc_rpn(C.byref(prog), x.ctypes.data_as(C.c_void_p), 3, ta.ctypes.data_as(C.c_void_p), C.c_double(0.0))
            assert np.isclose(v, ref[i], rtol=1e-13, atol=1e-13), (c, v, ref[i])
    # mutable parameters
    e = Expression("t*x[0]", degree=1, t=0.0)
    cc = as_coefficient(e)
    e.t = 2.5
    assert np.isclose(cc.eval_host(np.array([[2.0, 0, 0]]))[0], 5.0)
    assert cc.program().params[0] == 2.5
