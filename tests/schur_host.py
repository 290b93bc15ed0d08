"""numpy model of the device Schur solve (rake up-sweep, core solve, down-sweep) -- TEST INFRASTRUCTURE: pins the
algebra of graphnics_b200/schur_plan.py's plans on the CPU, incl. the rank views of a partitioned network."""
import numpy as np


def solve_host(plan, cond, diag, rhs, core_solver=None):
    d = np.array(diag, dtype=np.float64)
    r = np.array(rhs, dtype=np.float64)
    x = np.zeros(plan.B)
    w = np.where(plan.pedge >= 0, -np.asarray(cond)[np.maximum(plan.pedge, 0)], 0.0)

    def gather(nodes):
        for i in nodes:
            for c in plan.ch_idx[plan.ch_ptr[i]:plan.ch_ptr[i + 1]]:
                d[i] -= w[c] * w[c] / d[c]
                r[i] -= w[c] * r[c] / d[c]
    for lv in plan.levels:
        gather(lv)
    if plan.nc:
        gather(plan.core_nodes)
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        rows = np.repeat(np.arange(plan.nc), np.diff(plan.crp))
        S = sp.csr_matrix((-np.asarray(cond)[plan.cedge], (rows, plan.ccol)), shape=(plan.nc, plan.nc))
        S = S + sp.diags(d[plan.core_nodes])
        x[plan.core_nodes] = spla.spsolve(S.tocsc(), r[plan.core_nodes]) if core_solver is None \
            else core_solver(S, r[plan.core_nodes])
    for lv in reversed(plan.levels):
        for i in lv:
            p = plan.parent[i]
            x[i] = (r[i] - (w[i] * x[p] if p >= 0 else 0.0)) / d[i]
    return x


def solve_host_ranks(views, cond_of_rank, diag_rows, rhs_rows, edge_exported):
    """Play the ranks of a partitioned solve: views[r] = RankPlan of rank r; cond_of_rank[r] = the edge scalars
    rank r can see BEFORE the exchange (NaN elsewhere); diag_rows / rhs_rows(cond) build the Schur rows of given
    nodes from a full conductance array.  Returns (x per rank, number of exported (d, r) pairs per rank)."""
    world = len(views)
    B = views[0].B
    # ---- before the exchange: every rank sweeps its local nodes with what it owns
    d = [np.full(B, np.nan) for _ in range(world)]
    r = [np.full(B, np.nan) for _ in range(world)]
    sent = []
    for k, v in enumerate(views):
        c = cond_of_rank[k]
        w = np.where(v.pedge >= 0, -c[np.maximum(v.pedge, 0)], 0.0)
        for lv in v.levels[:v.l_sync]:
            for i in lv:
                d[k][i], r[k][i] = diag_rows(c, i), rhs_rows(c, i)
                for ch in v.ch_idx[v.ch_ptr[i]:v.ch_ptr[i + 1]]:
                    d[k][i] -= w[ch] * w[ch] / d[k][ch]
                    r[k][i] -= w[ch] * r[k][ch] / d[k][ch]
        sent.append(np.nonzero(v.ncls & 4)[0])
    # ---- the exchange: exported edge scalars and subtree-root pairs reach everybody
    cond_x = []
    for k in range(world):
        c = cond_of_rank[k].copy()
        for q in range(world):
            ex = edge_exported[q]
            c[ex] = cond_of_rank[q][ex]
        cond_x.append(c)
        for q in range(world):
            d[k][sent[q]] = d[q][sent[q]]
            r[k][sent[q]] = r[q][sent[q]]
    # ---- after it: replicated nodes (identical on every rank), then the local down-sweeps
    xs = []
    for k, v in enumerate(views):
        c = cond_x[k]
        w = np.where(v.pedge >= 0, -c[np.maximum(v.pedge, 0)], 0.0)

        def fold(i):
            for ch in v.ch_idx[v.ch_ptr[i]:v.ch_ptr[i + 1]]:
                d[k][i] -= w[ch] * w[ch] / d[k][ch]
                r[k][i] -= w[ch] * r[k][ch] / d[k][ch]
        for lv in v.levels[v.l_sync:]:
            for i in lv:
                d[k][i], r[k][i] = diag_rows(c, i), rhs_rows(c, i)
                fold(i)
        x = np.full(B, np.nan)
        if v.nc:
            import scipy.sparse as sp
            import scipy.sparse.linalg as spla
            for i in v.core_nodes:
                d[k][i], r[k][i] = diag_rows(c, i), rhs_rows(c, i)
                fold(i)
            rows = np.repeat(np.arange(v.nc), np.diff(v.crp))
            S = sp.csr_matrix((-c[v.cedge], (rows, v.ccol)), shape=(v.nc, v.nc)) + sp.diags(d[k][v.core_nodes])
            x[v.core_nodes] = spla.spsolve(S.tocsc(), r[k][v.core_nodes])
        for lv in reversed(v.levels):
            for i in lv:
                p = v.parent[i]
                x[i] = (r[k][i] - (w[i] * x[p] if p >= 0 else 0.0)) / d[k][i]
        xs.append(x)
    return xs, [len(s_) for s_ in sent]
