"""CPU, world_size 2 (gloo): the host-side logic of the edge-partitioned solve --
contiguous equal edge ranges, ONE all-gather of the packed per-rank [cond | rsrc | ftot] blocks,
the chunk/stride addressing the Schur kernel uses on the gathered buffer, replicated Schur solve
through the rake plan, local back-substitution -- with the arithmetic played by the numpy oracle.
Every rank's share of the solution is compared with the monolithic LU solve.  Also the max-over-ranks
timing reduction bench.py uses."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _rank_main(rank, world, port, name, n, out):
    import torch
    import torch.distributed as dist
    import oracle
    from _util import golden_arrays, random_graph, host_tables
    from graphnics_b200.device import partition_range
    from graphnics_b200.schur_plan import SchurPlan
    from schur_host import solve_host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, edges = random_graph(3, V=24, extra=6) if name == "rand" else golden_arrays(name)
        om = oracle.build_mesh(pos, edges, n)
        lay = oracle.mixed_layout(om)
        E, m = om.E, om.m
        rng = np.random.default_rng(5)
        Res = rng.uniform(0.5, 2.0, size=E)
        A = oracle.assemble_mixed(om, Res)
        b = oracle.assemble_mixed_rhs(om, f=oracle.Coef(lambda x: np.cos(x[..., 0]), 1),
                                      g=oracle.Coef(lambda x: x[..., 1], 1),
                                      p_bc=oracle.Coef(lambda x: 1 + x[..., 0], 1), project=False)
        x_ref = oracle.lu_solve(A, b)
        # ---- this rank's edges
        e0, e1, chunk = partition_range(E, rank, world)
        mine = np.arange(e0, e1)
        order = oracle.edge_order(om)
        cond, rsrc, ftot = oracle.eliminate_edges(om, b, Res, edges=mine, order=order)
        pack = torch.zeros(3 * chunk, dtype=torch.float64)
        for i, arr in enumerate((cond, rsrc, ftot)):
            pack[i * chunk: i * chunk + len(mine)] = torch.from_numpy(arr)
        gathered = torch.empty(world * 3 * chunk, dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, pack)
        g = gathered.numpy()
        # the kernel's addressing: arr[(e / chunk) * rank_stride + e % chunk]
        e = np.arange(E)
        idx = (e // chunk) * (3 * chunk) + e % chunk
        cond_g, rsrc_g, ftot_g = g[idx], g[chunk:][idx], g[2 * chunk:][idx]
        c_all, r_all, f_all = oracle.eliminate_edges(om, b, Res, order=order)
        assert np.array_equal(cond_g, c_all) and np.array_equal(rsrc_g, r_all) and np.array_equal(ftot_g, f_all)
        # ---- replicated Schur solve (rake plan + core), identical on every rank
        S, rhs, bix = oracle.schur_system(om, cond_g, rsrc_g, ftot_g, b[lay["lam_off"]:])
        t = host_tables(len(pos), edges)
        plan = SchurPlan(edges, t["bix"], t["B"])
        P = solve_host(plan, cond_g, S.diagonal(), rhs)
        Pt = torch.from_numpy(P.copy())
        lo, hi = Pt.clone(), Pt.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "replicated multipliers differ between ranks"
        # ---- local back-substitution and comparison of this rank's share
        x = oracle.backsub_edges(om, b, Res, cond_g, rsrc_g, P, bix, edges=mine, order=order)
        own = np.concatenate([np.arange(e0 * (m + 1), e1 * (m + 1)), lay["p_off"] + np.arange(e0 * m, e1 * m)])
        err = np.linalg.norm(x[own] - x_ref[own]) / np.linalg.norm(x_ref[own])
        errl = np.linalg.norm(P - x_ref[lay["lam_off"]:]) / max(np.linalg.norm(x_ref[lay["lam_off"]:]), 1e-30)
        # ---- timing reduction used by bench.py: max over ranks
        tt = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out.put((rank, float(err), float(errl), float(tt[0]), len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,n,port", [("honeycomb_4_4", 2, 29621), ("arterial_tree_N7", 3, 29622), ("rand", 1, 29623)])
def test_partitioned_algorithm_world2_gloo(name, n, port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, name, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, err, errl, tmax, ne in res:
        assert err < 1e-10 and errl < 1e-10, (rank, err, errl)
        assert tmax == 2.0 and ne > 0


def test_partition_ranges_cover_all_edges():
    from graphnics_b200.device import partition_range
    for E in (1, 7, 2047, 4095, 524287):
        for world in (1, 2, 3, 4, 8):
            if world > E:
                continue
            rs = [partition_range(E, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == E
            for a, b in zip(rs[:-1], rs[1:]):
                assert a[1] == b[0] and a[2] == b[2]
            assert all(e1 - e0 <= c for e0, e1, c in rs)
