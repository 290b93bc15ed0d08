"""CPU: the subtree-aligned edge partition (graphnics_b200/partition.py), the per-rank views of the Schur plan
(schur_plan.RankPlan) and the numpy model of the partitioned Schur solve -- local sweeps, ONE exchange of the
edge scalars at replicated nodes and of the local subtree roots' (d, r) pairs, replicated top -- against a
sparse direct solve of the full Schur matrix."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from _util import host_tables, random_graph, golden_arrays
from graphnics_b200.partition import partition_edges, contiguous_partition, bifurcation_index
from graphnics_b200.schur_plan import SchurPlan, RankPlan, node_owners
from graphnics_b200.synthetic import murray_tree
from schur_host import solve_host_ranks
from test_schur_plan import _schur


def _lollipop():
    rng = np.random.default_rng(11)
    nc = 60
    th = 2 * np.pi * np.arange(nc) / nc
    pos = [np.stack([np.cos(th), np.sin(th)], 1) * 50]
    edges = [np.stack([np.arange(nc), (np.arange(nc) + 1) % nc], 1)]
    nxt = nc
    for k in range(0, nc, 3):
        par = [k]
        for lev in range(4):
            new = []
            for p in par:
                for _ in range(2):
                    pos.append(pos[0][k][None] * (1.1 + 0.1 * lev) + rng.normal(size=(1, 2)))
                    edges.append(np.array([[p, nxt]]))
                    new.append(nxt); nxt += 1
            par = new
    pos, edges = np.vstack(pos), np.vstack(edges)
    return pos, edges[np.lexsort((edges[:, 1], edges[:, 0]))]


CASES = {
    "tree8": lambda: murray_tree(8)[:2],
    "tree12": lambda: murray_tree(12)[:2],
    "arterial_tree_N7": lambda: golden_arrays("arterial_tree_N7"),
    "rand1": lambda: random_graph(1, V=40, extra=2),
    "rand2": lambda: random_graph(2, V=25, extra=9),
    "honeycomb_4_4": lambda: golden_arrays("honeycomb_4_4"),
    "lollipop": _lollipop,
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("world", [2, 3, 8])
def test_partitioned_schur_model_matches_direct_solve(name, world):
    pos, edges = CASES[name]()
    edges = np.asarray(edges)
    E, V = len(edges), len(pos)
    if world > E:
        pytest.skip("more ranks than edges")
    t = host_tables(V, edges)
    bix, B = t["bix"], t["B"]
    epart, plan = partition_edges(edges, V, world)
    assert epart.min() >= 0 and epart.max() < world and len(epart) == E
    bix2, B2 = bifurcation_index(V, edges)
    assert B2 == B and np.array_equal(bix2, bix)
    owner = node_owners(plan, edges, bix, epart)
    views = [RankPlan(plan, owner, r) for r in range(world)]
    # every raked node is swept by exactly one rank or replicated on all; core nodes are replicated
    raked = np.concatenate(plan.levels) if plan.levels else np.zeros(0, dtype=np.int64)
    for i in raked:
        mine = [r for r, v in enumerate(views) if any(i in lv for lv in v.levels[:v.l_sync])]
        repl = [r for r, v in enumerate(views) if any(i in lv for lv in v.levels[v.l_sync:])]
        assert (len(mine) == 1 and not repl and owner[i] == mine[0]) or (not mine and len(repl) == world and owner[i] < 0)
    assert (owner[plan.core_nodes] < 0).all()
    # a local node's edges are all on its rank
    bu, bv = bix[edges[:, 0]], bix[edges[:, 1]]
    for e in range(E):
        for b in (bu[e], bv[e]):
            if b >= 0 and owner[b] >= 0:
                assert epart[e] == owner[b]
    # numbers: random conductances over 6 decades, Schur rows from the conductances a rank can see
    rng = np.random.default_rng(1)
    cond = rng.uniform(0.5, 2.0, size=E) * 10.0 ** rng.uniform(-3, 3, size=E)
    S, diag = _schur(edges, bix, B, cond)
    rhs = rng.normal(size=B)
    inc = [[] for _ in range(B)]
    for e in range(E):
        for b in (bu[e], bv[e]):
            if b >= 0:
                inc[b].append(e)

    def diag_rows(c, i):
        v = sum(c[e] for e in inc[i])
        assert np.isfinite(v), "a rank built a row from a scalar it was never sent"
        return v
    cond_of_rank = [np.where(epart == r, cond, np.nan) for r in range(world)]
    at_repl = (np.where(bu >= 0, owner[np.maximum(bu, 0)] < 0, False) | np.where(bv >= 0, owner[np.maximum(bv, 0)] < 0, False))
    exported = [np.nonzero(at_repl & (epart == r))[0] for r in range(world)]
    xs, nsent = solve_host_ranks(views, cond_of_rank, diag_rows, lambda c, i: rhs[i], exported)
    xr = spla.spsolve(S.tocsc(), rhs)
    for r, (x, v) in enumerate(zip(xs, views)):
        mine = np.concatenate(v.levels + [v.core_nodes]) if (v.levels or v.nc) else np.zeros(0, dtype=np.int64)
        assert np.abs(x[mine] - xr[mine]).max() <= 1e-9 * np.abs(xr).max()
    # replicated values are identical bits on every rank
    repl = np.nonzero(owner < 0)[0]
    for x in xs[1:]:
        assert np.array_equal(x[repl].view(np.int64), xs[0][repl].view(np.int64))


@pytest.mark.parametrize("levels,world,nrepl", [(11, 2, 1), (13, 4, 3), (14, 8, 7), (19, 8, 7)])
def test_binary_tree_is_cut_into_equal_subtrees(levels, world, nrepl):
    """C2 per rank (bench.py --gpus N) and C5 on 8 GPUs: 2^k ranks get one subtree each, 2^k - 1 bifurcations are
    replicated, (edges at them + subtree roots) is all that crosses NVLink"""
    pos, edges, _, _ = murray_tree(levels, gdim=2)
    E, V = len(edges), len(pos)
    epart, plan = partition_edges(edges, V, world)
    bix, B = bifurcation_index(V, edges)
    owner = node_owners(plan, edges, bix, epart)
    assert int((owner < 0).sum()) == nrepl
    counts = np.bincount(epart, minlength=world)
    assert counts.max() - counts.min() <= world and counts.max() <= -(-E // world) + 1
    views = [RankPlan(plan, owner, r) for r in range(world)]
    for v in views:
        assert int((v.ncls & 4 > 0).sum()) == 1          # one exported subtree root per rank
        assert v.n_mine == (B - nrepl) // world + nrepl
    bu, bv = bix[edges[:, 0]], bix[edges[:, 1]]
    at_repl = (np.where(bu >= 0, owner[np.maximum(bu, 0)] < 0, False) | np.where(bv >= 0, owner[np.maximum(bv, 0)] < 0, False))
    assert int(at_repl.sum()) == 2 * nrepl + 1          # 3 edges per replicated node, shared ones counted once


def test_cyclic_network_falls_back_to_index_ranges():
    pos, edges = golden_arrays("honeycomb_4_4")
    epart, plan = partition_edges(edges, len(pos), 3)
    assert plan.nc == plan.B
    c = np.bincount(epart, minlength=3)
    assert c.max() - c.min() <= 2 and (np.diff(epart) >= 0).all()
    assert np.array_equal(contiguous_partition(10, 3), [0, 0, 0, 0, 1, 1, 1, 1, 2, 2])
