"""GPU parity on awkward inputs: refinements beyond the edge-tile kernels (n = 11, 12: the
thread-per-cell fallback kernels with several chunks per edge), a lonely node, antiparallel edges
(a 2-cycle), two disconnected components, a single edge without refinement, u > v everywhere."""
import numpy as np
import pytest

import oracle
from _util import y_graph

pytestmark = pytest.mark.gpu


def _cases():
    pos, edges = y_graph()
    yield "Y n=11", pos, edges, 11
    yield "Y n=12", pos, edges, 12
    yield "Y reversed n=11", pos, edges[:, ::-1].copy()[np.lexsort((edges[:, 0], edges[:, 1]))], 11
    # lonely node 4
    yield "lonely node", np.vstack([pos, [[3.0, 3.0]]]), edges, 3
    # antiparallel edges between two bifurcations + tails
    p = np.array([[0, 0], [1, 0], [2, 0.5], [3, 0], [1, 1.0]], dtype=np.float64)
    e = np.array([[0, 1], [1, 2], [1, 4], [2, 1], [2, 3]])
    yield "antiparallel", p, e, 4
    # two components
    p = np.array([[0, 0], [1, 0], [2, 0], [0, 5], [1, 5], [2, 6], [2, 4]], dtype=np.float64)
    e = np.array([[0, 1], [1, 2], [3, 4], [4, 5], [4, 6]])
    yield "two components", p, e, 2
    yield "single edge n=0", np.array([[0.0, 0.0, 0.0], [0.3, 0.4, 1.2]]), np.array([[0, 1]]), 0
    yield "single reversed edge n=5", np.array([[0.0, 0.0], [2.0, 1.0]]), np.array([[1, 0]]), 5


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_mixed_and_primal_on_awkward_inputs(case):
    import torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    from graphnics_b200.engine_primal import PrimalSystem
    from graphnics_b200._lib import check, ptr
    from graphnics_b200.device import current_stream
    name, pos, edges, n = case
    om = oracle.build_mesh(pos, edges, n)
    dn = DeviceNetwork(pos, edges, n)
    assert np.array_equal(dn.cells.cpu().numpy(), om.cells)
    assert np.array_equal(dn.coords.cpu().numpy().view(np.int64), om.coords.view(np.int64))
    assert dn.bifurcation_ixs() == om.bif_ixs and dn.boundary_ixs() == om.bnd_ixs
    assert np.array_equal(dn.vf_host(), om.vf)
    rng = np.random.default_rng(2)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    gd = pos.shape[1]
    lin = oracle.Coef(lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., gd - 1], 1)
    f = oracle.Coef(lambda x: np.cos(2 * x[..., 1]) - x[..., 0], 1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    # ---- mixed
    ms = MixedSystem(dn)
    A = oracle.assemble_mixed(om, Res)
    b = oracle.assemble_mixed_rhs(om, f=f, p_bc=lin, project=False)
    rp, ci = ms.build_pattern()
    assert np.array_equal(rp.cpu().numpy(), A.indptr) and np.array_equal(ci.cpu().numpy(), A.indices)
    co = ms.coeffs(Res)
    vals = ms.assemble(co)
    np.testing.assert_allclose(vals.cpu().numpy(), A.data, rtol=1e-12, atol=1e-300)
    xq, _ = ms.dof_coordinates(False)
    xqh = xq.cpu().numpy()
    bg = ms.rhs(dev(f(xqh)), None, ms.end_values(dev(lin(xqh))), dev(lin(pos)))
    np.testing.assert_allclose(bg.cpu().numpy(), b, rtol=1e-12, atol=1e-13)
    x, info = ms.solve_checked(co, vals, bg)
    x_ref = oracle.lu_solve(A, b)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-8 * np.linalg.norm(x_ref), (name, info)
    # ---- primal (needs every component to touch a Dirichlet node: true for all cases here)
    ps = PrimalSystem(dn)
    Ap = oracle.assemble_primal(om, 1.3)
    rp, ci = ps.build_pattern()
    assert np.array_equal(rp.cpu().numpy(), Ap.indptr) and np.array_equal(ci.cpu().numpy(), Ap.indices)
    cop = ps.coeffs(1.3)
    pv = ps.assemble(cop)
    np.testing.assert_allclose(pv.cpu().numpy(), Ap.data, rtol=1e-12, atol=1e-300)
    if name == "lonely node":
        return                                   # its p row is empty: the system is singular in the reference too
    bp = oracle.assemble_primal_rhs(om, f=0.7, g=-0.4)
    dofs, bvals = oracle.primal_bc(om, lin)
    Abc, bbc = oracle.apply_bc(Ap, bp, dofs, bvals)
    bt = ps.rhs(None, None)
    from graphnics_b200.coefficients import as_coefficient
    bt = ps.rhs(as_coefficient(0.7), as_coefficient(-0.4))
    check(ps.lib.gnx_apply_bc_primal(dn.net_ref(), 1.0, ptr(dev(lin(pos))), ptr(pv), ptr(bt), current_stream()))
    np.testing.assert_allclose(bt.cpu().numpy(), bbc, rtol=1e-12, atol=1e-13)
    xp, infop = ps.solve_checked(cop, pv, bt)
    xp_ref = oracle.lu_solve(Abc, bbc)
    assert np.linalg.norm(xp.cpu().numpy() - xp_ref) <= 1e-8 * np.linalg.norm(xp_ref), (name, infop)
