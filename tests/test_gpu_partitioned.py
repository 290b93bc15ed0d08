"""Edge-partitioned execution on ONE GPU: the ranks of a 2-, 3- and 4-way partition
are played one after the other in this process (same kernels, same buffers; the all-gather is a
concatenation), and every rank's local solution is compared with the oracle's monolithic LU solve.
The NCCL path proper is exercised by tools/check_partitioned.py under torchrun (gpurun --gpus N)."""
import numpy as np
import pytest

import oracle
from _util import random_graph, golden_arrays

pytestmark = pytest.mark.gpu


def _tree(levels):
    from graphnics_b200.synthetic import murray_tree
    pos, edges, _, _ = murray_tree(levels, gdim=2)
    return pos, edges


@pytest.mark.parametrize("name,n,world", [("tree7", 3, 2), ("tree7", 3, 4), ("honeycomb_4_4", 2, 3),
                                          ("rand2", 4, 2), ("arterial_tree_N7", 1, 4), ("tree9", 0, 4)])
def test_partitioned_solve_matches_monolithic_lu(name, n, world):
    import torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    if name.startswith("tree"):
        pos, edges = _tree(int(name[4:]))
    elif name.startswith("rand"):
        pos, edges = random_graph(int(name[4:]), V=20, extra=5)
    else:
        pos, edges = golden_arrays(name)
    om = oracle.build_mesh(pos, edges, n)
    lay = oracle.mixed_layout(om)
    E, m = len(edges), 1 << n
    rng = np.random.default_rng(3)
    Res = rng.uniform(0.5, 2.0, size=E)
    lin = oracle.Coef(lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1], 1)
    g = oracle.Coef(lambda x: np.sin(3 * x[..., 0]) + x[..., 1], 1)
    f = oracle.Coef(lambda x: np.cos(2 * x[..., 1]) - x[..., 0], 1)
    A = oracle.assemble_mixed(om, Res)
    b = oracle.assemble_mixed_rhs(om, f=f, g=g, p_bc=lin, project=False)
    x_ref = oracle.lu_solve(A, b)

    ranks = []
    for r in range(world):
        dn = DeviceNetwork(pos, edges, n, part=(r, world))
        ms = MixedSystem(dn)
        # local rhs from the same nodal data, through the kernels
        xq, _ = ms.dof_coordinates(False)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        xqh = xq.cpu().numpy()
        bl = ms.rhs(dev(f(xqh)), dev(g(xqh)), ms.end_values(dev(lin(xqh))), dev(lin(pos)))
        co = ms.coeffs(Res[dn.e0:dn.e1])
        # this rank's share of the global vectors
        idx = np.concatenate([np.arange(dn.e0 * (m + 1), dn.e1 * (m + 1)),
                              lay["p_off"] + np.arange(dn.e0 * m, dn.e1 * m),
                              lay["lam_off"] + np.arange(dn.B)])
        assert dn.B == om.B and ms.N == len(idx)
        np.testing.assert_allclose(bl.cpu().numpy()[:dn.lam_off], b[idx][:dn.lam_off], rtol=1e-12, atol=1e-14)
        ms.eliminate(co, bl)
        ranks.append((dn, ms, co, bl, idx))
    # the all-gather, played by hand
    gathered = torch.cat([ms._bufs()["pack"] for _, ms, _, _, _ in ranks])

    def exchange(d):
        d["gathered"].copy_(gathered)
    P = []
    for dn, ms, co, bl, idx in ranks:
        x, info = ms.solve(co, bl, exchange=exchange, eliminated=True)
        assert info.cg_converged
        x = x.cpu().numpy()
        for lo, hi in ((0, dn.p_off), (dn.p_off, dn.lam_off), (dn.lam_off, ms.N)):
            ref = x_ref[idx][lo:hi]
            assert np.linalg.norm(x[lo:hi] - ref) <= 1e-8 * max(np.linalg.norm(ref), 1e-30), (name, world, lo)
        P.append(ms._bufs()["P"].cpu().numpy().copy())
        # subdomain matrix: q and p rows equal the global rows restricted to the local columns
        vals = ms.assemble(co).cpu().numpy()
        rp, ci = (t.cpu().numpy() for t in ms.build_pattern())
        import scipy.sparse as sp
        Al = sp.csr_matrix((vals, ci, rp), shape=(ms.N, ms.N))
        Ag = A[idx][:, idx]
        assert abs(Al[:dn.lam_off] - Ag[:dn.lam_off]).max() <= 1e-12 * abs(Ag).max()
    # replicated multipliers are bit-identical on every rank
    for r in range(1, world):
        assert np.array_equal(P[0].view(np.int64), P[r].view(np.int64))
    # lambda rows of the subdomain matrices sum to the global lambda rows
    S = None
    for dn, ms, co, bl, idx in ranks:
        vals = ms.assemble(co).cpu().numpy()
        rp, ci = (t.cpu().numpy() for t in ms.build_pattern())
        import scipy.sparse as sp
        Al = sp.csr_matrix((vals, ci, rp), shape=(ms.N, ms.N))[dn.lam_off:]
        # scatter local columns to global columns
        G = sp.csr_matrix((np.ones(len(idx)), (np.arange(len(idx)), idx)), shape=(len(idx), lay["N"]))
        part = (Al @ G)
        S = part if S is None else S + part
    # the explicit-zero lambda diagonal is present on every rank: it sums to zero as well
    assert abs(S - A[lay["lam_off"]:]).max() <= 1e-12
