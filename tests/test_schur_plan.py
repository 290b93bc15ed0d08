"""CPU-only: the symbolic rake/core plan of the Schur solve (graphnics_b200/schur_plan.py) and the
numpy model of the device algorithm, against a sparse direct solve of the full Schur matrix."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from _util import host_tables, random_graph, golden_arrays
from graphnics_b200.schur_plan import SchurPlan
from graphnics_b200.synthetic import murray_tree


def _schur(edges, bix, B, cond):
    bu, bv = bix[edges[:, 0]], bix[edges[:, 1]]
    diag = np.zeros(B)
    np.add.at(diag, bu[bu >= 0], cond[bu >= 0])
    np.add.at(diag, bv[bv >= 0], cond[bv >= 0])
    inner = (bu >= 0) & (bv >= 0)
    S = sp.coo_matrix((np.concatenate([-cond[inner], -cond[inner]]),
                       (np.concatenate([bu[inner], bv[inner]]), np.concatenate([bv[inner], bu[inner]]))),
                      shape=(B, B)).tocsr() + sp.diags(diag)
    return S, diag


def _line(nn):
    pos = np.zeros((nn, 2)); pos[:, 0] = np.arange(nn)
    return pos, np.stack([np.arange(nn - 1), np.arange(1, nn)], 1)


CASES = {
    "tree8": lambda: murray_tree(8)[:2],
    "rand0": lambda: random_graph(0, V=30, extra=4),
    "rand1": lambda: random_graph(1, V=40, extra=2),
    "rand2": lambda: random_graph(2, V=25, extra=9),
    "honeycomb_4_4": lambda: golden_arrays("honeycomb_4_4"),
    "YY_bifurcation": lambda: golden_arrays("YY_bifurcation"),
    "line3": lambda: golden_arrays("line_graph_3_dim3"),
    "arterial_tree_N7": lambda: golden_arrays("arterial_tree_N7"),
    "line40": lambda: _line(40),
    "line400": lambda: _line(400),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_plan_solves_schur_system(name):
    pos, edges = CASES[name]()
    edges = np.asarray(edges)
    t = host_tables(len(pos), edges)
    bix, B = t["bix"], t["B"]
    rng = np.random.default_rng(1)
    cond = rng.uniform(0.5, 2.0, size=len(edges)) * 10.0 ** rng.uniform(-3, 3, size=len(edges))
    S, diag = _schur(edges, bix, B, cond)
    rhs = rng.normal(size=B)
    plan = SchurPlan(edges, bix, B)
    # structure: every node is raked exactly once or is in the core
    lv = np.concatenate(plan.levels) if plan.levels else np.zeros(0, dtype=np.int64)
    assert len(np.unique(np.concatenate([lv, plan.core_nodes]))) == B == len(lv) + plan.nc
    # a raked node's parent is raked later or sits in the core
    level_of = np.full(B, 10 ** 9)
    for l, nodes in enumerate(plan.levels):
        level_of[nodes] = l
    for i in lv:
        p = plan.parent[i]
        assert p < 0 or level_of[p] > level_of[i]
    x = plan.solve_host(cond, diag, rhs)
    xr = spla.spsolve(S.tocsc(), rhs)
    assert np.abs(x - xr).max() <= 1e-10 * np.abs(xr).max()
    if name.startswith("tree") or name.startswith("arterial"):
        assert plan.nc == 0
    if name == "honeycomb_4_4":
        assert plan.nc == B
