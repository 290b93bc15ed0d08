"""GPU parity tests (through the C ABI, via graphnics_b200.engine) against the CPU oracle.

Bit-exact: mesh coordinates, cells, markers, tangents, tags, node tables, dof numbering, CSR
pattern.  1e-12 relative: matrix entries, right-hand side.  1e-8 relative L2: solution fields.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from _util import y_graph, random_graph, golden_arrays, golden_graphs

pytestmark = pytest.mark.gpu

CASES = [("Y", 0), ("Y", 1), ("Y", 3), ("Y", 6), ("Y", 10), ("Y3", 4), ("rand0", 0), ("rand0", 3),
         ("rand1", 5), ("rand2", 7), ("honeycomb_2_2", 2), ("honeycomb_4_4", 5), ("YY_bifurcation", 5),
         ("arterial_tree_N5", 3), ("arterial_tree_N7", 6), ("line_graph_3_dim3", 5), ("line_graph_2_dx2", 8)]


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
    from graphnics_b200.engine import MixedSystem
    assert torch.cuda.is_available()
    return DeviceNetwork, MixedSystem


@pytest.mark.parametrize("name,n", CASES)
def test_mesh_and_tables_bit_exact(eng, name, n):
    DeviceNetwork, _ = eng
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    assert np.array_equal(dn.cells.cpu().numpy(), om.cells)
    assert np.array_equal(dn.mf.cpu().numpy(), om.mf)
    assert np.array_equal(dn.coords.cpu().numpy().view(np.int64), om.coords.view(np.int64))
    assert np.array_equal(dn.tangents.cpu().numpy().view(np.int64), om.tangents.view(np.int64))
    assert dn.bifurcation_ixs() == om.bif_ixs
    assert dn.boundary_ixs() == om.bnd_ixs
    assert np.array_equal(dn.vf_host(), om.vf)
    assert np.array_equal(dn.parent_vertex_host(), om.parent_vertex)
    # cell lengths (edge-local order) against the oracle geometry
    h = dn.h.cpu().numpy().reshape(dn.E, dn.m)
    assert np.allclose(h.sum(axis=1), oracle.edge_lengths(pos, edges), rtol=1e-13)
    assert dn.B == om.B and dn.mixed_ndofs == oracle.mixed_layout(om)["N"]


@pytest.mark.parametrize("name,n", CASES)
def test_pattern_values_rhs(eng, name, n):
    DeviceNetwork, MixedSystem = eng
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    ms = MixedSystem(dn)
    rng = np.random.default_rng(3)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    A = oracle.assemble_mixed(om, Res)
    rowptr, colidx = ms.build_pattern()
    assert np.array_equal(rowptr.cpu().numpy(), A.indptr)
    assert np.array_equal(colidx.cpu().numpy(), A.indices)
    vals = ms.assemble(ms.coeffs(Res)).cpu().numpy()
    np.testing.assert_allclose(vals, A.data, rtol=1e-12, atol=1e-300)
    visc = rng.uniform(0.1, 1.0, size=len(edges))
    A2 = oracle.assemble_mixed(om, Res, visc=visc)
    np.testing.assert_allclose(ms.assemble(ms.coeffs(Res, visc)).cpu().numpy(), A2.data, rtol=1e-11, atol=1e-300)
    D = oracle.assemble_mass_q_mixed(om, 2.5)
    Dg = sp.csr_matrix((ms.assemble_mass_q(2.5).cpu().numpy(), A.indices, A.indptr), shape=A.shape)
    assert abs(Dg - D).max() <= 1e-12 * abs(D).max()
    # rhs with nodal (un-projected) data
    import torch
    lin = oracle.Coef(lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1], 1)
    g = oracle.Coef(lambda x: np.sin(x[..., 0]) + x[..., 1] ** 2, 1)
    f = oracle.Coef(lambda x: np.cos(x[..., 1]) - x[..., 0], 1)
    b = oracle.assemble_mixed_rhs(om, f=f, g=g, p_bc=lin, project=False)
    xq, xm = ms.dof_coordinates()
    assert np.array_equal(xq.cpu().numpy().reshape(om.sub_coords.shape).view(np.int64), om.sub_coords.view(np.int64))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    nod = lambda c: dev(c(om.sub_coords).ravel())
    bg = ms.rhs(nod(f), nod(g), ms.end_values(nod(lin)), dev(lin(pos))).cpu().numpy()
    np.testing.assert_allclose(bg, b, rtol=1e-12, atol=1e-14)
    # Constant data (edge-tile kernel), incl. a constant p_bc on both kinds of boundary ends
    b2 = oracle.assemble_mixed_rhs(om, f=0.75, g=-1.25, p_bc=2.0)
    bg2 = ms.rhs(None, None, dev(np.full(len(edges), 2.0)), dev(np.full(len(pos), 2.0)), f_const=0.75, g_const=-1.25)
    np.testing.assert_allclose(bg2.cpu().numpy(), b2, rtol=1e-12, atol=1e-14)
    b3 = oracle.assemble_mixed_rhs(om, p_bc=lin, project=False)
    bg3 = ms.rhs(None, None, ms.end_values(nod(lin)), dev(lin(pos)))
    np.testing.assert_allclose(bg3.cpu().numpy(), b3, rtol=1e-12, atol=1e-14)
    # SpMV against scipy
    x = rng.normal(size=A.shape[0])
    y = ms.spmv(ms.assemble(ms.coeffs(Res)), dev(x)).cpu().numpy()
    np.testing.assert_allclose(y, A @ x, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name,n", CASES)
def test_solve_matches_lu(eng, name, n):
    import torch
    DeviceNetwork, MixedSystem = eng
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    ms = MixedSystem(dn)
    rng = np.random.default_rng(5)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    lin = oracle.Coef(lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1], 1)
    g = oracle.Coef(lambda x: np.sin(3 * x[..., 0]) + x[..., 1], 1)
    f = oracle.Coef(lambda x: np.cos(2 * x[..., 1]) - x[..., 0], 1)
    A = oracle.assemble_mixed(om, Res)
    b = oracle.assemble_mixed_rhs(om, f=f, g=g, p_bc=lin, project=False)
    x_ref = oracle.lu_solve(A, b)
    co = ms.coeffs(Res)
    vals = ms.assemble(co)
    bd = torch.from_numpy(b).cuda()
    x, info = ms.solve_checked(co, vals, bd)
    x = x.cpu().numpy()
    lay = oracle.mixed_layout(om)
    q, qr = x[:lay["p_off"]], x_ref[:lay["p_off"]]
    p, pr = x[lay["p_off"]:lay["lam_off"]], x_ref[lay["p_off"]:lay["lam_off"]]
    assert info.cg_converged
    assert rel(q, qr) < 1e-8 and rel(p, pr) < 1e-8, (rel(q, qr), rel(p, pr), info)
    if om.B:
        assert rel(x[lay["lam_off"]:], x_ref[lay["lam_off"]:]) < 1e-8
    assert info.rel_resid < 1e-10, info


def test_stokes_and_sigma_scaled_solve(eng):
    """NetworkStokes-type operator (a M + k K) and a theta-scheme operator (sigma != 1)."""
    import torch
    DeviceNetwork, MixedSystem = eng
    pos, edges = graph("rand1")
    n = 4
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    ms = MixedSystem(dn)
    rng = np.random.default_rng(7)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    visc = rng.uniform(0.1, 1.0, size=len(edges))
    b = rng.normal(size=ms.N)
    b[oracle.mixed_layout(om)["lam_off"]:] = 0.0
    A = oracle.assemble_mixed(om, Res, visc=visc)
    x_ref = oracle.lu_solve(A, b)
    co = ms.coeffs(Res, visc)
    x, info = ms.solve_checked(co, ms.assemble(co), torch.from_numpy(b).cuda())
    assert rel(x.cpu().numpy(), x_ref) < 1e-8, info
    # D + theta*dt*A with D = Ainv*M:  a = Ainv + s*Res, sigma = s
    s, Ainv = 0.05, 2.0
    Aeff = (oracle.assemble_mass_q_mixed(om, Ainv) + s * oracle.assemble_mixed(om, Res)).tocsr()
    x_ref = oracle.lu_solve(Aeff, b)
    co = ms.coeffs(Ainv + s * Res, None, s)
    vals = ms.assemble(co)
    rp, ci = ms.build_pattern()
    Ag = sp.csr_matrix((vals.cpu().numpy(), ci.cpu().numpy(), rp.cpu().numpy()), shape=Aeff.shape)
    assert abs(Ag - Aeff).max() <= 1e-12 * abs(Aeff).max()
    x, info = ms.solve_checked(co, vals, torch.from_numpy(b).cuda())
    assert rel(x.cpu().numpy(), x_ref) < 1e-8, info


def test_reference_known_answers(eng):
    """SURVEY.md Appendix C (Ohm/Kirchhoff closed form) and tests/test_flow_models.py:17-71
    (mass conservation at every bifurcation)."""
    import torch
    DeviceNetwork, MixedSystem = eng
    pos, edges = y_graph()
    dn = DeviceNetwork(pos, edges, 6)
    ms = MixedSystem(dn)
    co = ms.coeffs([1.0, 2.0, 3.0])
    pbc = torch.from_numpy(2.0 - pos[:, 1]).cuda()
    xq, _ = ms.dof_coordinates(False)
    b = ms.rhs(None, None, ms.end_values((2.0 - xq[:, 1]).contiguous()), pbc)
    x, info = ms.solve(co, b)
    x = x.cpu().numpy()
    m = dn.m
    q = x[:dn.p_off].reshape(3, m + 1)
    assert np.allclose(q[0], 0.741549228561398, rtol=1e-12)
    assert np.allclose(q[1], 0.444929537136839, rtol=1e-12)
    assert np.allclose(q[2], 0.296619691424559, rtol=1e-12)
    assert np.isclose(x[dn.lam_off], 1.629225385719301, rtol=1e-12)
    # mass conservation on the reference's own graphs (n=5, Res=1, p_bc = x[0])
    for name in ("line_graph_3_dim3", "Y_bifurcation", "YY_bifurcation", "honeycomb_4_4"):
        pos, edges = golden_arrays(name)
        dn = DeviceNetwork(pos, edges, 5)
        ms = MixedSystem(dn)
        co = ms.coeffs(1.0)
        xq, _ = ms.dof_coordinates(False)
        b = ms.rhs(None, None, ms.end_values(xq[:, 0].contiguous()), torch.from_numpy(np.ascontiguousarray(pos[:, 0])).cuda())
        x, info = ms.solve(co, b)
        x = x.cpu().numpy()
        q = x[:dn.p_off].reshape(dn.E, dn.m + 1)
        loc = dn.loc_host
        for bnode in dn.bifurcation_ixs():
            tot = 0.0
            for i, (u, v) in enumerate(edges):
                flip = u > v
                if v == bnode:
                    tot += q[i, loc[0] if flip else loc[dn.m]]
                if u == bnode:
                    tot -= q[i, loc[dn.m] if flip else loc[0]]
            assert abs(tot) < 1e-10, (name, bnode, tot)


def _big_graph(name):
    from graphnics_b200.synthetic import murray_tree
    if name == "tree14":                      # B = 8191: multi-CTA rake with a single-CTA top
        pos, edges, _, _ = murray_tree(14, gdim=2)
        return pos, edges
    if name == "honeycomb40":                 # B ~ 3.4k, no rakeable node: cooperative PCG on the core
        from graphnics_b200.generators import honeycomb
        G = honeycomb(40, 40, mesh=False)
        return G._graph_arrays()
    if name == "rtree":                       # unbalanced random tree, B > 2048: blocked rake with uneven blocks
        rng = np.random.default_rng(4)
        nn = 12000
        par = np.array([rng.integers(0, k) for k in range(1, nn)])
        edges = np.stack([par, np.arange(1, nn)], 1)
        return rng.normal(size=(nn, 2)) * 30.0, edges
    if name == "honeycomb70":                 # B ~ 10k > 8192: two-level PCG, one CTA per aggregate
        from graphnics_b200.generators import honeycomb
        G = honeycomb(70, 70, mesh=False)
        return G._graph_arrays()
    if name in ("lollipop", "lollipop9k"):    # cycle (core) with trees hanging off it (raked)
        rng = np.random.default_rng(11)
        nc = 3000 if name == "lollipop" else 9000
        th = 2 * np.pi * np.arange(nc) / nc
        pos = [np.stack([np.cos(th), np.sin(th)], 1) * 50]
        edges = [np.stack([np.arange(nc), (np.arange(nc) + 1) % nc], 1)]
        nxt = nc
        for k in range(0, nc, 3):             # a 3-level binary tree on every third cycle node
            par = [k]
            for lev in range(3):
                new = []
                for p in par:
                    for _ in range(2):
                        pos.append(pos[0][k][None] * (1.1 + 0.1 * lev) + rng.normal(size=(1, 2)))
                        edges.append(np.array([[p, nxt]]))
                        new.append(nxt); nxt += 1
                par = new
        pos, edges = np.vstack(pos), np.vstack(edges)
        order = np.lexsort((edges[:, 1], edges[:, 0]))
        return pos, edges[order]
    raise KeyError(name)


@pytest.mark.parametrize("name,n", [("tree14", 2), ("honeycomb40", 1), ("lollipop", 1), ("honeycomb70", 1),
                                    ("lollipop9k", 1), ("rtree", 1)])
def test_solve_large_graphs(eng, name, n):
    """systems whose Schur solve needs the cooperative multi-CTA kernel (B > 2048)"""
    import torch
    DeviceNetwork, MixedSystem = eng
    pos, edges = _big_graph(name)
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    assert dn.B > 2048
    ms = MixedSystem(dn)
    rng = np.random.default_rng(9)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    lin = oracle.Coef(lambda x: 1.0 + 0.2 * x[..., 0] - 0.05 * x[..., 1], 1)
    f = oracle.Coef(lambda x: np.cos(0.2 * x[..., 1]), 1)
    A = oracle.assemble_mixed(om, Res)
    b = oracle.assemble_mixed_rhs(om, f=f, g=0.0, p_bc=lin, project=False)
    x_ref = oracle.lu_solve(A, b)
    co = ms.coeffs(Res)
    vals = ms.assemble(co)
    rp, ci = ms.build_pattern()
    assert np.array_equal(rp.cpu().numpy(), A.indptr) and np.array_equal(ci.cpu().numpy(), A.indices)
    np.testing.assert_allclose(vals.cpu().numpy(), A.data, rtol=1e-12, atol=1e-300)
    x, info = ms.solve_checked(co, vals, torch.from_numpy(b).cuda())
    plan = ms.schur_plan()[3]
    if name in ("tree14", "rtree"):
        assert plan.nc == 0 and info.cg_iterations == 0
        assert plan.blocks is not None and plan.blocks["nb"] > 1          # blocked rake
    if name == "honeycomb40":
        assert plan.nc == dn.B and info.cg_iterations > 10
    if name == "lollipop":
        assert 0 < plan.nc < dn.B
    if name in ("honeycomb70", "lollipop9k"):  # core > 8192 rows: two-level PCG on 128 CTAs
        assert plan.nc > 8192 and plan.agg["na"] == 128 and info.cg_batches == 128
        if name == "lollipop9k":
            assert plan.nc < dn.B
    assert info.cg_converged and abs(info.cg_batches) > 1      # > 1 CTA (negative: one cluster)
    lay = oracle.mixed_layout(om)
    x = x.cpu().numpy()
    for lo, hi in ((0, lay["p_off"]), (lay["p_off"], lay["lam_off"]), (lay["lam_off"], lay["N"])):
        assert rel(x[lo:hi], x_ref[lo:hi]) < 1e-8, (name, lo, rel(x[lo:hi], x_ref[lo:hi]), info)
    # warm start from the converged multipliers: no more than a couple of iterations
    info2 = ms.solve(co, torch.from_numpy(b).cuda(), x=torch.empty_like(torch.from_numpy(b).cuda()), warm_start=True)[1]
    assert info2.cg_iterations <= max(2, info.cg_iterations // 4)


@pytest.mark.parametrize("name,n", [("Y", 0), ("Y", 1), ("Y", 6), ("rand1", 5), ("honeycomb_4_4", 5), ("arterial_tree_N7", 6),
                                    ("line_graph_2_dx2", 8)])
def test_fused_rhs_elimination_and_overlapped_step(eng, name, n):
    """gnx_rhs_eliminate_mixed (Constant f, g: b is produced by the elimination kernel) gives the same b
    bit for bit as gnx_assemble_rhs_mixed, and the fused / overlapped step the same solution as the
    plain sequence assemble -> rhs -> solve"""
    import torch
    DeviceNetwork, MixedSystem = eng
    pos, edges = graph(name)
    dn = DeviceNetwork(pos, edges, n)
    ms = MixedSystem(dn)
    rng = np.random.default_rng(8)
    co = ms.coeffs(rng.uniform(0.5, 2.0, size=len(edges)))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    kw = dict(pbc_out=dev(rng.normal(size=len(edges))), pbc_node=dev(rng.normal(size=len(pos))), f_const=0.75, g_const=-1.25)
    b_ref = ms.rhs(**kw)
    vals_ref = ms.assemble(co)
    x_ref, _ = ms.solve(co, b_ref)
    for overlap in (False, True):
        vals = torch.full((ms.nnz,), np.nan, dtype=torch.float64, device="cuda")
        b = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
        x = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
        ms.assemble_rhs_solve(co, vals, b, x, kw, overlap=overlap)
        torch.cuda.synchronize()
        assert np.array_equal(b.cpu().numpy().view(np.int64), b_ref.cpu().numpy().view(np.int64))
        assert np.array_equal(vals.cpu().numpy().view(np.int64), vals_ref.cpu().numpy().view(np.int64))
        assert np.array_equal(x.cpu().numpy().view(np.int64), x_ref.cpu().numpy().view(np.int64))


@pytest.mark.parametrize("name,n", [("Y", 6), ("honeycomb_4_4", 5), ("arterial_tree_N7", 6), ("arterial_tree_N7", 9)])
def test_edge_records_change_no_bit(eng, name, n):
    """gnx_network_t.erec (one 16-byte record per edge instead of the chain edges -> bix -> ebif_prefix) and the
    copy-engine staging of the cell lengths in the streaming assembly are a faster way to the SAME numbers: values, b
    and x with and without the records are identical bit for bit (persistent kernel and kernel chain alike)"""
    import torch
    DeviceNetwork, MixedSystem = eng
    pos, edges = graph(name)
    rng = np.random.default_rng(11)
    res = rng.uniform(0.5, 2.0, size=len(edges))
    pbc_o, pbc_n = rng.normal(size=len(edges)), rng.normal(size=len(pos))
    out = []
    for with_records in (True, False):
        dn = DeviceNetwork(pos, edges, n)
        if not with_records:
            dn.erec = None
            dn._net = None
            assert not dn.net.erec
        else:
            assert dn.net.erec
        ms = MixedSystem(dn)
        co = ms.coeffs(res)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        kw = dict(pbc_out=dev(pbc_o), pbc_node=dev(pbc_n))
        vals_a = ms.assemble(co)
        vals = torch.full((ms.nnz,), np.nan, dtype=torch.float64, device="cuda")
        b = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
        x = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
        ms.assemble_rhs_solve(co, vals, b, x, kw)
        torch.cuda.synchronize()
        out.append([t.cpu().numpy().view(np.int64) for t in (vals_a, vals, b, x)])
    for a, c in zip(*out):
        assert np.array_equal(a, c)
    assert np.array_equal(out[0][0], out[0][1])            # stand-alone assembly == the step's


def test_edge_flux_path_of_large_networks_changes_no_bit(eng):
    """networks too large for the persistent kernel: the one-call step gathers the edges' end flux / end pressure once
    (gnx_edge_flux_mixed) and back-substitutes from them (gnx_edge_backsub_mixed_flux); the separate entry points
    gather per block (gnx_edge_backsub_mixed_const) or read b (gnx_edge_backsub_mixed) -- same x bit for bit"""
    import ctypes as C
    import torch
    from graphnics_b200.synthetic import murray_tree
    DeviceNetwork, MixedSystem = eng
    pos, edges, _, _ = murray_tree(15, gdim=2)
    dn = DeviceNetwork(pos, edges, 6)
    ms = MixedSystem(dn)
    rng = np.random.default_rng(5)
    co = ms.coeffs(rng.uniform(0.5, 2.0, size=len(edges)))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    kw = dict(pbc_out=dev(rng.normal(size=len(edges))), pbc_node=dev(rng.normal(size=len(pos))), f_const=0.5, g_const=-0.25)
    b_ref = ms.rhs(**kw)
    x_ref, _ = ms.solve(co, b_ref)                       # elimination of a given b, generic back-substitution
    vals = torch.empty(ms.nnz, dtype=torch.float64, device="cuda")
    b = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
    x = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
    ms.assemble_rhs_solve(co, vals, b, x, kw)            # one-call step: kernel chain with the flux path
    torch.cuda.synchronize()
    assert not int(ms.lib.gnx_step_is_resident())
    assert np.array_equal(b.cpu().numpy().view(np.int64), b_ref.cpu().numpy().view(np.int64))
    assert np.array_equal(x.cpu().numpy().view(np.int64), x_ref.cpu().numpy().view(np.int64))
    # the two entry points on their own, from the step's scratch arrays (cond is intact; rsrc / ftot now hold Pu / q0)
    d = ms._bufs()
    x2 = torch.full((ms.N,), np.nan, dtype=torch.float64, device="cuda")
    ptr = lambda t: C.c_void_p(t.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    rc = ms.lib.gnx_edge_backsub_mixed_flux(dn.net_ref(), co.ref(), 0.5, -0.25, ptr(kw["pbc_out"]), ptr(kw["pbc_node"]),
                                            ptr(d["ftot"]), ptr(d["rsrc"]), ptr(x2), st)
    assert rc == 0
    torch.cuda.synchronize()
    q_p = slice(0, dn.lam_off)
    assert np.array_equal(x2[q_p].cpu().numpy().view(np.int64), x_ref[q_p].cpu().numpy().view(np.int64))
