"""CPU-only: the symbolic rake/core plan of the Schur solve (graphnics_b200/schur_plan.py) and the
numpy model of the device algorithm, against a sparse direct solve of the full Schur matrix."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from _util import host_tables, random_graph, golden_arrays
from graphnics_b200.schur_plan import SchurPlan
from schur_host import solve_host
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
    x = solve_host(plan, cond, diag, rhs)
    xr = spla.spsolve(S.tocsc(), rhs)
    assert np.abs(x - xr).max() <= 1e-10 * np.abs(xr).max()
    if name.startswith("tree") or name.startswith("arterial"):
        assert plan.nc == 0
    if name == "honeycomb_4_4":
        assert plan.nc == B


def _two_level_pcg(plan, agg, cond, diag, rhs, rtol=1e-13, max_iter=5000):
    """numpy model of core_pcg_twolevel (gnx_solve.cu): the coarse matrix comes from the plan's cut lists
    and row sums, the fine-level diagonal from row_sub / col_agg, exactly as the kernel sums them;
    z = D^-1 r + W1 dE^-1 W1^T r + W2 E^-1 W2^T r is never stored whole."""
    nc, na, ns = plan.nc, agg["na"], agg["ns"]
    nf = na * ns
    rows = np.repeat(np.arange(nc), np.diff(plan.crp))
    cval = -np.asarray(cond)[plan.cedge]
    dg = np.asarray(diag)[plan.core_nodes]
    S = sp.csr_matrix((cval, (rows, plan.ccol)), shape=(nc, nc)) + sp.diags(dg)
    rowsum = dg + np.bincount(rows, weights=cval, minlength=nc)
    E = np.zeros((na, na))
    for I in range(na):
        for J in range(na):
            if I != J:
                ks = agg["cut_k"][agg["cut_ptr"][I * na + J]:agg["cut_ptr"][I * na + J + 1]]
                E[I, J] = cval[ks].sum() if len(ks) else 0.0
        E[I, I] = rowsum[agg["agg_ptr"][I]:agg["agg_ptr"][I + 1]].sum() - E[I].sum()
    row_sub, col_sub = agg["row_sub"], agg["col_agg"]
    assert np.array_equal(row_sub, np.repeat(np.arange(nf), np.diff(agg["sub_ptr"])))
    assert np.array_equal(agg["agg_ptr"], agg["sub_ptr"][::ns]) and np.array_equal(col_sub, row_sub[plan.ccol])
    same = col_sub == row_sub[rows]
    dE = np.bincount(row_sub, weights=dg, minlength=nf) + np.bincount(row_sub[rows][same], weights=cval[same], minlength=nf)
    agg_of = row_sub // ns
    W2 = sp.csr_matrix((np.ones(nc), (np.arange(nc), agg_of)), shape=(nc, na))
    W1 = sp.csr_matrix((np.ones(nc), (np.arange(nc), row_sub)), shape=(nc, nf))
    assert np.allclose(E, (W2.T @ S @ W2).toarray(), rtol=1e-12, atol=1e-12 * abs(E).max())
    assert np.allclose(dE, (W1.T @ S @ W1).diagonal(), rtol=1e-12)
    Ei = np.linalg.inv(E)
    b = np.asarray(rhs)[plan.core_nodes]

    def prec(r):
        s1 = np.bincount(row_sub, weights=r, minlength=nf)
        s2 = s1.reshape(na, ns).sum(axis=1)
        y1 = (s1 / dE) if ns > 1 else np.zeros(nf)
        return r / dg + y1[row_sub] + (Ei @ s2)[agg_of]
    x = np.zeros(nc); r = b.copy()
    z = prec(r)
    p = z.copy(); rz = r @ z; bb = b @ b
    for it in range(1, max_iter + 1):
        w = S @ p
        al = rz / (p @ w)
        x += al * p; r -= al * w
        if r @ r <= rtol * rtol * bb:
            return x, it, S
        z = prec(r)
        rzn = r @ z
        p = z + (rzn / rz) * p
        rz = rzn
    return x, max_iter, S


def _jacobi_pcg_iters(S, b, rtol=1e-13, max_iter=20000):
    d = S.diagonal()
    x = np.zeros_like(b); r = b.copy(); z = r / d; p = z.copy(); rz = r @ z; bb = b @ b
    for it in range(1, max_iter + 1):
        w = S @ p
        al = rz / (p @ w)
        x += al * p; r -= al * w
        if r @ r <= rtol * rtol * bb:
            return it
        z = r / d; rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return max_iter


@pytest.mark.parametrize("na,ns", [(2, 1), (8, 1), (32, 1), (8, 4), (32, 8)])
def test_two_level_aggregates(na, ns):
    """aggregate_core: balanced contiguous aggregates, renumbered core CSR = permuted original, cut lists
    give W^T S W, and the two-level PCG reaches the direct solution in fewer iterations than Jacobi"""
    from graphnics_b200.generators import honeycomb
    G = honeycomb(24, 24, mesh=False)
    nodes = list(G.nodes())
    idx = {v: i for i, v in enumerate(nodes)}
    pos = np.array([G.nodes[v]["pos"] for v in nodes], dtype=np.float64)
    edges = np.array([(idx[u], idx[v]) for u, v in G.edges()], dtype=np.int64)
    t = host_tables(len(pos), edges)
    bix, B = t["bix"], t["B"]
    rng = np.random.default_rng(3)
    cond = rng.uniform(0.5, 2.0, size=len(edges))
    S_full, diag = _schur(edges, bix, B, cond)
    rhs = rng.normal(size=B)
    plan = SchurPlan(edges, bix, B)
    assert plan.nc == B                                     # a lattice has nothing to rake
    before = sp.csr_matrix((-cond[plan.cedge], (np.repeat(np.arange(plan.nc), np.diff(plan.crp)), plan.ccol)),
                           shape=(plan.nc, plan.nc))
    old_nodes = plan.core_nodes.copy()
    node_of = np.empty(B, dtype=np.int64)
    node_of[bix[bix >= 0]] = np.nonzero(bix >= 0)[0]
    agg = plan.aggregate_core(pos[node_of], na=na, ns=ns)
    assert agg["na"] == na and agg["ns"] == ns and agg["agg_ptr"][0] == 0 and agg["agg_ptr"][-1] == plan.nc
    sizes = np.diff(agg["agg_ptr"])
    assert sizes.min() >= 1 and sizes.max() - sizes.min() <= 1 + int(np.log2(na))
    assert sorted(plan.core_nodes.tolist()) == sorted(old_nodes.tolist())
    perm = plan.core_of[old_nodes]                          # old core row -> new core row
    after = sp.csr_matrix((-cond[plan.cedge], (np.repeat(np.arange(plan.nc), np.diff(plan.crp)), plan.ccol)),
                          shape=(plan.nc, plan.nc))
    P = sp.csr_matrix((np.ones(plan.nc), (perm, np.arange(plan.nc))), shape=(plan.nc, plan.nc))
    assert abs(P @ before @ P.T - after).max() == 0.0
    agg_of = np.repeat(np.arange(na), sizes)
    assert np.array_equal(agg["col_agg"] // ns, agg_of[plan.ccol])
    x, its, S = _two_level_pcg(plan, agg, cond, diag, rhs)
    ref = spla.spsolve(S_full.tocsc(), rhs)
    assert np.allclose(x, ref[plan.core_nodes], rtol=1e-9, atol=1e-11)
    jac = _jacobi_pcg_iters(S, rhs[plan.core_nodes])
    assert its < jac, (its, jac)
    # the plan's own model (direct core solve) still works on the renumbered core
    assert np.allclose(solve_host(plan, cond, diag, rhs), ref, rtol=1e-9, atol=1e-11)


def _blocked_rake_model(plan, ba, cond, diag, rhs):
    """numpy model of the blocked rake (gnx_solve.cu: blk_up -> top part -> blk_down)"""
    d = np.array(diag, dtype=np.float64); r = np.array(rhs, dtype=np.float64)
    w = np.where(plan.pedge >= 0, -np.asarray(cond)[np.maximum(plan.pedge, 0)], 0.0)
    lev = ba["level_of"]
    kids = lambda i: plan.ch_idx[plan.ch_ptr[i]:plan.ch_ptr[i + 1]]
    # blocks, independently of each other: only in-block children are touched
    for b in range(ba["nb"]):
        nodes = ba["blk_nodes"][ba["blk_ptr"][b]:ba["blk_ptr"][b + 1]]
        assert list(lev[nodes]) == sorted(lev[nodes])
        inside = set(nodes.tolist())
        for i in nodes:
            for c in kids(i):
                assert c in inside
                d[i] -= w[c] * w[c] / d[c]; r[i] -= w[c] * r[c] / d[c]
    # upper part, level by level
    x = np.zeros(plan.B)
    top = ba["top_nodes"]
    for i in top:                                           # sorted by level
        for c in kids(i):
            d[i] -= w[c] * w[c] / d[c]; r[i] -= w[c] * r[c] / d[c]
    for i in top[::-1]:
        p = plan.parent[i]
        x[i] = (r[i] - (w[i] * x[p] if p >= 0 else 0.0)) / d[i]
    for b in range(ba["nb"]):
        nodes = ba["blk_nodes"][ba["blk_ptr"][b]:ba["blk_ptr"][b + 1]]
        for i in nodes[::-1]:
            p = plan.parent[i]
            x[i] = (r[i] - (w[i] * x[p] if p >= 0 else 0.0)) / d[i]
    return x


def _random_tree(nn, seed, chain=0.0):
    rng = np.random.default_rng(seed)
    par = np.array([rng.integers(max(0, k - 1 if rng.random() < chain else 0), k) for k in range(1, nn)])
    edges = np.stack([par, np.arange(1, nn)], 1)
    pos = rng.normal(size=(nn, 2))
    return pos, edges


@pytest.mark.parametrize("name,cap,top_max", [("tree13", 1024, 2048), ("tree10", 100, 64), ("rtree", 300, 2048),
                                              ("rchain", 128, 4096), ("forest", 50, 500)])
def test_blocked_rake_plan(name, cap, top_max):
    if name.startswith("tree"):
        pos, edges = murray_tree(int(name[4:]), gdim=2)[:2]
    elif name == "rtree":
        pos, edges = _random_tree(6000, 1)
    elif name == "rchain":
        pos, edges = _random_tree(5000, 2, chain=0.9)
    else:                                                   # several trees side by side
        parts, off = [], 0
        for k in range(5):
            p, e = _random_tree(400 + 100 * k, 10 + k)
            parts.append((p, e + off)); off += len(p)
        pos = np.vstack([p for p, _ in parts]); edges = np.vstack([e for _, e in parts])
    edges = np.asarray(edges)
    t = host_tables(len(pos), edges)
    bix, B = t["bix"], t["B"]
    rng = np.random.default_rng(5)
    cond = rng.uniform(0.5, 2.0, size=len(edges))
    S, diag = _schur(edges, bix, B, cond)
    rhs = rng.normal(size=B)
    plan = SchurPlan(edges, bix, B)
    ba = plan.block_arrays(cap=cap, top_max=top_max)
    if plan.nc != 0:
        assert ba is None
        return
    assert ba is not None, "upper part larger than expected"
    sizes = np.diff(ba["blk_ptr"])
    assert ba["nb"] == len(sizes) and sizes.min() >= 1 and sizes.max() <= cap
    assert len(ba["top_nodes"]) <= top_max and sizes.sum() + len(ba["top_nodes"]) == B
    in_blk = ba["blk_pos"] >= 0
    assert in_blk.sum() == sizes.sum() and not in_blk[ba["top_nodes"]].any()
    assert np.array_equal(np.sort(np.concatenate([ba["blk_nodes"], ba["top_nodes"]])), np.arange(B))
    # the root of a block is the only node whose parent lies outside it (or is missing)
    for b in range(ba["nb"]):
        nodes = ba["blk_nodes"][ba["blk_ptr"][b]:ba["blk_ptr"][b + 1]]
        inside = set(nodes.tolist())
        outside = [i for i in nodes if plan.parent[i] < 0 or plan.parent[i] not in inside]
        assert len(outside) == 1 and outside[0] == nodes[-1]
        assert ba["blk_lv"][2 * b] == ba["level_of"][nodes[0]] and ba["blk_lv"][2 * b + 1] == ba["level_of"][nodes[-1]]
        assert np.array_equal(ba["blk_pos"][nodes], np.arange(len(nodes)))
    if len(ba["top_nodes"]):
        assert ba["l_top"] == ba["level_of"][ba["top_nodes"]].min()
    x = _blocked_rake_model(plan, ba, cond, diag, rhs)
    ref = spla.spsolve(S.tocsc(), rhs)
    assert np.allclose(x, ref, rtol=1e-9, atol=1e-11)
