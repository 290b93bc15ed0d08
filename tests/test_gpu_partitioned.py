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
                                          ("rand2", 4, 2), ("arterial_tree_N7", 1, 4), ("tree9", 0, 4),
                                          ("honeycomb48", 0, 2)])     # core > 2048 rows: replicated two-level PCG
def test_partitioned_solve_matches_monolithic_lu(name, n, world):
    import torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    if name.startswith("tree"):
        pos, edges = _tree(int(name[4:]))
    elif name.startswith("rand"):
        pos, edges = random_graph(int(name[4:]), V=20, extra=5)
    elif name == "honeycomb48":
        from graphnics_b200.generators import honeycomb
        pos, edges = honeycomb(48, 48, mesh=False)._graph_arrays()
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
        if name == "honeycomb48":
            assert ms.schur_plan()[3].agg is not None and info.cg_batches == 128
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


@pytest.mark.parametrize("name,n,world", [("honeycomb_4_4", 2, 3), ("tree7", 3, 4), ("rand2", 3, 2)])
def test_partitioned_primal_solve_matches_monolithic_lu(name, n, world):
    """HydraulicNetwork edge-partitioned (BASELINE config 3): q and interior p per rank, node p replicated"""
    import torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine_primal import PrimalSystem
    if name.startswith("tree"):
        pos, edges = _tree(int(name[4:]))
    elif name.startswith("rand"):
        pos, edges = random_graph(int(name[4:]), V=20, extra=5)
    else:
        pos, edges = golden_arrays(name)
    om = oracle.build_mesh(pos, edges, n)
    lay = oracle.primal_layout(om)
    E, m, V = len(edges), 1 << n, len(pos)
    pf = lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1]
    A = oracle.assemble_primal(om, 1.3)
    b = oracle.assemble_primal_rhs(om, f=0.0, g=-0.4)
    dofs, bvals = oracle.primal_bc(om, oracle.Coef(pf, 1))
    Abc, bbc = oracle.apply_bc(A, b, dofs, bvals)
    x_ref = oracle.lu_solve(Abc, bbc)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    from graphnics_b200.coefficients import as_coefficient
    ranks = []
    for r in range(world):
        dn = DeviceNetwork(pos, edges, n, part=(r, world))
        ps = PrimalSystem(dn)
        co = ps.coeffs(1.3)
        bl = ps.rhs(None, as_coefficient(-0.4))
        ps.apply_bc_rhs(1.0, dev(pf(pos)), bl)
        # Dirichlet values are replicated data: every rank holds them at every boundary node
        bn = bl[ps.p_off: ps.p_off + V]
        bnd = torch.from_numpy(np.asarray(om.bnd_ixs)).cuda()
        bn[bnd] = dev(pf(pos))[bnd]
        ps.eliminate(co, bl)
        ranks.append((dn, ps, co, bl))
    gathered = torch.cat([ps._bufs()["pack"] for _, ps, _, _ in ranks])

    def exchange(d):
        d["gathered"].copy_(gathered)
    for dn, ps, co, bl in ranks:
        x, info = ps.solve(co, bl, exchange=exchange, eliminated=True)
        assert info.cg_converged
        x = x.cpu().numpy()
        # q of the local cells = global cells [e0*m, e1*m); p at the graph nodes; p at local interior vertices
        q_ref = x_ref[dn.e0 * m: dn.e1 * m]
        assert np.linalg.norm(x[:ps.NC] - q_ref) <= 1e-8 * max(np.linalg.norm(q_ref), 1e-30), (name, world)
        touched = np.unique(edges[dn.e0:dn.e1])
        p_nodes = x[ps.p_off: ps.p_off + V]
        p_ref = x_ref[lay["p_off"]: lay["p_off"] + V]
        assert np.linalg.norm(p_nodes[touched] - p_ref[touched]) <= 1e-8 * np.linalg.norm(p_ref[touched])
        # interior vertices: local id V + (k-level block) maps to the global id of the same (edge, position)
        pv_loc = dn.parent_vertex_host()                      # [E_loc, m+1] local vertex ids
        pv_glob = om.parent_vertex[dn.e0:dn.e1]               # global vertex ids, same local numbering pattern
        interior = pv_loc >= V
        xi = x[ps.p_off + pv_loc[interior]]
        ri = x_ref[lay["p_off"] + pv_glob[interior]]
        assert np.linalg.norm(xi - ri) <= 1e-8 * max(np.linalg.norm(ri), 1e-30)


def _case(name):
    if name.startswith("tree"):
        return _tree(int(name[4:]))
    if name.startswith("rand"):
        return random_graph(int(name[4:]), V=20, extra=5)
    if name == "lollipop":
        from test_partition import _lollipop
        return _lollipop()
    if name == "honeycomb48":
        from graphnics_b200.generators import honeycomb
        return honeycomb(48, 48, mesh=False)._graph_arrays()
    return golden_arrays(name)


@pytest.mark.parametrize("name,n,world", [("tree7", 3, 2), ("tree7", 6, 4), ("tree12", 1, 8), ("tree13", 0, 4), ("tree14", 0, 2),
                                          ("arterial_tree_N7", 2, 3), ("rand2", 4, 2), ("lollipop", 2, 4),
                                          ("honeycomb_4_4", 2, 3), ("honeycomb48", 0, 2)])
@pytest.mark.parametrize("fused", [False, True])
def test_subtree_partitioned_solve_matches_monolithic_lu(name, n, world, fused):
    """gnx_exchange mode 1 with the ranks played in one process: every rank eliminates and rakes ITS subtrees, stores
    what it exports straight into the other ranks' buffers (phase 1 of all ranks), then every rank processes the
    replicated bifurcations and sweeps down (phase 2).  Each rank's share of the solution against the oracle's LU."""
    import torch
    from graphnics_b200.parallel import PartitionedMixed
    pos, edges = _case(name)
    om = oracle.build_mesh(pos, edges, n)
    lay = oracle.mixed_layout(om)
    E = len(edges)
    rng = np.random.default_rng(3)
    Res = rng.uniform(0.5, 2.0, size=E)
    lin = oracle.Coef(lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1], 1)
    A = oracle.assemble_mixed(om, Res)
    if fused:                                 # Constant data: the elimination kernel produces b
        b = oracle.assemble_mixed_rhs(om, f=0.75, g=-1.25, p_bc=lin, project=False)
    else:
        g = oracle.Coef(lambda x: np.sin(3 * x[..., 0]) + x[..., 1], 1)
        f = oracle.Coef(lambda x: np.cos(2 * x[..., 1]) - x[..., 0], 1)
        b = oracle.assemble_mixed_rhs(om, f=f, g=g, p_bc=lin, project=False)
    x_ref = oracle.lu_solve(A, b)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    pms = [PartitionedMixed(pos, edges, n, rank=r, world=world) for r in range(world)]
    assert all(pm.ms.subtree for pm in pms)
    nw = pms[0].ms.xch_words()
    slots = [[torch.full((nw,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(world)] for _ in range(2)]
    ranks = []
    for pm in pms:
        ms, dn = pm.ms, pm.dn
        ms.xch_attach(slots)
        idx = pm.local_dofs_of_global()
        co = ms.coeffs(Res[pm.edge_ids()])
        xq, _ = ms.dof_coordinates(False)
        xqh = xq.cpu().numpy()
        if fused:
            kw = dict(const_rhs=dict(f_const=0.75, g_const=-1.25, pbc_out=ms.end_values(dev(lin(xqh))), pbc_node=dev(lin(pos))))
            bl = torch.full((ms.N,), float("nan"), dtype=torch.float64, device="cuda")
        else:
            kw = {}
            bl = ms.rhs(dev(f(xqh)), dev(g(xqh)), ms.end_values(dev(lin(xqh))), dev(lin(pos)))
        x = torch.full((ms.N,), float("nan"), dtype=torch.float64, device="cuda")
        ranks.append((pm, co, bl, x, kw, idx))
    for rep in range(3):                      # repeated solves alternate the two buffers
        for pm, co, bl, x, kw, idx in ranks:
            pm.ms.solve(co, bl, x=x, xphase=1, sync=False, **kw)
        infos = [pm.ms.solve(co, bl, x=x, xphase=2, **kw)[1] for pm, co, bl, x, kw, idx in ranks]
    lam_all = []
    for (pm, co, bl, x, kw, idx), info in zip(ranks, infos):
        dn = pm.dn
        assert info.cg_converged
        xh = x.cpu().numpy()
        np.testing.assert_allclose(bl.cpu().numpy()[:dn.lam_off], b[idx][:dn.lam_off], rtol=1e-12, atol=1e-14)
        for lo, hi in ((0, dn.p_off), (dn.p_off, dn.lam_off)):
            ref = x_ref[idx][lo:hi]
            assert np.linalg.norm(xh[lo:hi] - ref) <= 1e-8 * max(np.linalg.norm(ref), 1e-30), (name, world, pm.rank, lo)
        v = pm.valid_multipliers()
        lam, lam_ref = xh[dn.lam_off:][v], x_ref[lay["lam_off"]:][v]
        assert np.linalg.norm(lam - lam_ref) <= 1e-8 * max(np.linalg.norm(lam_ref), 1e-30)
        lam_all.append((v, xh[dn.lam_off:]))
    # every multiplier is computed somewhere; replicated ones carry identical bits on every rank
    seen = np.zeros(om.B, dtype=int)
    for v, _ in lam_all:
        seen[v] += 1
    assert (seen >= 1).all()
    repl = np.nonzero(seen == world)[0]
    for v, lam in lam_all[1:]:
        assert np.array_equal(lam[repl].view(np.int64), lam_all[0][1][repl].view(np.int64))
    plan0 = pms[0].ms.schur_plan()[3]
    if name.startswith("tree"):
        assert int((plan0.owner < 0).sum()) <= 2 * world        # a handful of replicated bifurcations
